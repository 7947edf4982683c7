#!/bin/bash
# full ncu capture of the CVC kernel (shipped build and the grouped-load build) + guide and WTA kernels, exported to CSV on the box
mkdir -p gpurun_out
for v in 0 1; do
  ncu --set full --clock-control none --import-source on -k regex:cvc_both -s 2 -c 1 -f -o /tmp/r2_cvc_v$v \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --cvc-variant $v > gpurun_out/bench_under_ncu_cvc_$v.log 2>&1
  ncu -i /tmp/r2_cvc_v$v.ncu-rep --page raw --csv > gpurun_out/r2_cvc_v${v}_raw.csv 2>/dev/null
  ncu -i /tmp/r2_cvc_v$v.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2_cvc_v${v}_source.csv.gz
done
ncu --set full --clock-control none -k regex:"guide_kernel|wta_kernel|ingest_kernel" -s 6 -c 4 -f -o /tmp/r2_small \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_under_ncu_small.log 2>&1
ncu -i /tmp/r2_small.ncu-rep --page raw --csv > gpurun_out/r2_small_kernels_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
