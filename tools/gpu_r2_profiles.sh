#!/bin/bash
# round-2 evidence for profiles/: launch list of the bench command, full ncu captures of the CVF kernel (exact, mixed; C4 and C3)
# exported to CSV on the box (the .ncu-rep files are too big to bring back), and clean bench lines.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_bench_c4.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_under_ncu.log 2>&1
for cfg in "C4 0" "C4 1" "C3 0"; do
  set -- $cfg
  ncu --set full --clock-control none --import-source on -k regex:cvf_stream -s 2 -c 1 -f -o /tmp/r2_cvf_$1_mode$2 \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --workload $1 --cvf-mode $2 > gpurun_out/bench_under_ncu_$1_$2.log 2>&1
  ncu -i /tmp/r2_cvf_$1_mode$2.ncu-rep --page raw --csv > gpurun_out/r2_cvf_$1_mode$2_raw.csv 2>/dev/null
  ncu -i /tmp/r2_cvf_$1_mode$2.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2_cvf_$1_mode$2_source.csv.gz
done
ncu --set full --clock-control none -k regex:"cvc_both|wta_kernel|pp_wmf|guide_kernel" -s 8 -c 5 -f -o /tmp/r2_other \
      python tools/pp_time.py C4 > gpurun_out/pp_under_ncu.log 2>&1
ncu -i /tmp/r2_other.ncu-rep --page raw --csv > gpurun_out/r2_other_kernels_raw.csv 2>/dev/null
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c4_exact.json 2> gpurun_out/bench_err.log; echo "exit $?" >> gpurun_out/bench_err.log
python bench.py --steps 20 --warmup 5 --cvf-mode 1 --no-cpu-baseline > gpurun_out/r2_bench_c4_mixed.json 2>> gpurun_out/bench_err.log
python bench.py --steps 20 --warmup 5 --workload C3 --no-cpu-baseline > gpurun_out/r2_bench_c3_exact.json 2>> gpurun_out/bench_err.log
python bench.py --steps 10 --warmup 3 --workload C5 --no-cpu-baseline > gpurun_out/r2_bench_c5_n1_exact.json 2>> gpurun_out/bench_err.log
python tools/pp_time.py C4 > gpurun_out/r2_pp_time.txt 2>&1
tail -3 gpurun_out/bench_err.log; cut -c1-300 gpurun_out/r2_bench_c4_exact.json; cat gpurun_out/r2_pp_time.txt; du -sh gpurun_out
