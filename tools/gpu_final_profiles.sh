#!/bin/bash
# evidence for profiles/: launch list of the bench command + one full ncu capture of the CVF kernel + a clean bench line
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cvf_stream -s 2 -c 1 -f -o gpurun_out/cvf_prof \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1
python bench.py > gpurun_out/bench_final.log 2>&1; echo "exit $?" >> gpurun_out/bench_final.log
tail -2 gpurun_out/bench_final.log | cut -c1-300
