#!/bin/bash
# Final single-GPU evidence for profiles/ (final code): full ncu captures of the CVF kernel exported to CSV on the box (the .ncu-rep
# files are too big to bring back), the launch list of the bench command, clean bench lines, reference arm, PP / FGF / write-ceiling
# timings, smoke and the whole GPU test suite.
mkdir -p gpurun_out
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
bash tools/gpu_r2_final2.sh
