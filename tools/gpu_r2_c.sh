#!/bin/bash
# round 2, run C: L1 prefetch variants + mixed ring protocol without register-held rows
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modes.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_modes.log 2>&1
echo "pytest modes exit $?" >> gpurun_out/pytest_modes.log
tail -3 gpurun_out/pytest_modes.log
rm -f gpurun_out/variants_c.txt
for cfg in "0 0" "0 5" "0 6" "1 0" "1 5" "1 4" "1 6" "1 5 128" "1 6 128" "0 5 128"; do
  set -- $cfg
  extra=""; [ -n "$3" ] && extra="--cta-threads $3"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --cvf-mode $1 --variant $2 $extra 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('mode',j['config']['cvf_mode'],'variant',j['config']['variant'],'$extra','ms/step',round(j['ms_per_step'],3),'cvf_kernel',round(j['roofline']['kernel_ms'],3),'frac',round(j['roofline']['frac'],4),'e2e',round(j['e2e']['ms_per_step'],3))
    else: print(l.rstrip()[:300])
" | tee -a gpurun_out/variants_c.txt
done
ncu --set full --clock-control none --import-source on -k regex:cvf_stream -s 2 -c 1 -f -o gpurun_out/cvf_prof_c_mode1 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --cvf-mode 1 --variant 5 > gpurun_out/bench_under_ncu_mode1.log 2>&1
