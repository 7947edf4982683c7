#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modes.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_modes.log 2>&1
echo "pytest modes exit $?" >> gpurun_out/pytest_modes.log
tail -3 gpurun_out/pytest_modes.log
rm -f gpurun_out/variants_e.txt
for cfg in "1 0" "1 7 224" "1 7 448" "1 8 416" "1 0 384" "0 0 384" "1 7 192" "1 8 192"; do
  set -- $cfg
  extra=""; [ -n "$3" ] && extra="--cta-threads $3"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --cvf-mode $1 --variant $2 $extra 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('mode',j['config']['cvf_mode'],'variant',j['config']['variant'],'$extra','ms/step',round(j['ms_per_step'],3),'cvf_kernel',round(j['roofline']['kernel_ms'],3),'frac',round(j['roofline']['frac'],4),'e2e',round(j['e2e']['ms_per_step'],3))
    else: print(l.rstrip()[:300])
" | tee -a gpurun_out/variants_e.txt
done
