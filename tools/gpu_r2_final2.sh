#!/bin/bash
# final single-GPU evidence for profiles/: launch list under ncu, clean bench lines (C4 exact default incl. tolerance-mode leg and
# CPU arm, C3, C5 on one GPU), the reference arm, PP and FGF timings, smoke.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_bench_c4.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_under_ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c4_exact.json 2> gpurun_out/bench_err.log; echo "exit $?" >> gpurun_out/bench_err.log
python bench.py --steps 20 --warmup 5 --cvf-mode 1 --no-cpu-baseline > gpurun_out/r2_bench_c4_mixed.json 2>> gpurun_out/bench_err.log
python bench.py --steps 20 --warmup 5 --workload C3 --no-cpu-baseline > gpurun_out/r2_bench_c3_exact.json 2>> gpurun_out/bench_err.log
python bench.py --steps 10 --warmup 3 --workload C5 --no-cpu-baseline > gpurun_out/r2_bench_c5_n1_exact.json 2>> gpurun_out/bench_err.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/bench_err.log
python tools/pp_time.py C4 > gpurun_out/r2_pp_time.txt 2>&1
python tools/fgf_time.py > gpurun_out/r2_fgf_time.txt 2>&1
python tools/write_peak.py > gpurun_out/r2_write_peak.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_final.log; tail -3 gpurun_out/pytest_final.log
tail -3 gpurun_out/bench_err.log; cut -c1-300 gpurun_out/r2_bench_c4_exact.json; cat gpurun_out/r2_pp_time.txt gpurun_out/r2_fgf_time.txt gpurun_out/r2_write_peak.txt; du -sh gpurun_out
