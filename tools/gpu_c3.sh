#!/bin/bash
# BASELINE config C3: synthetic 1280x720 D=64 fp32, 1xB200, ncu HBM-GB/s capture on CVF
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:cvf_stream -s 2 -c 1 -f -o gpurun_out/cvf_prof_c3 \
    python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_under_ncu.log 2>&1
python bench.py --workload C3 --steps 20 --warmup 3 > gpurun_out/bench_c3.log 2>&1; echo "exit $?" >> gpurun_out/bench_c3.log
python bench.py --workload C5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_n1.log 2>&1; echo "exit $?" >> gpurun_out/bench_c5_n1.log
tail -2 gpurun_out/bench_c3.log | cut -c1-200; tail -2 gpurun_out/bench_c5_n1.log | cut -c1-200
