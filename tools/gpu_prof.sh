#!/bin/bash
# launch list (all kernels, device time) + one full ncu capture of the fused CVF kernel
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cvf_stream -s 1 -c 1 -f -o gpurun_out/cvf_prof \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1
ls -la gpurun_out
