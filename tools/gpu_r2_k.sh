#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modes.py -m gpu -q -x --timeout 600 -k "ragged or model" > gpurun_out/pytest_k.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_k.log
tail -4 gpurun_out/pytest_k.log
rm -f gpurun_out/variants_k.txt
for cfg in "1 0" "1 14" "1 15" "1 14 --cta-threads 128" "0 14"; do
  set -- $cfg
  m=$1; v=$2; shift; shift
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --cvf-mode $m --variant $v "$@" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('mode',j['config']['cvf_mode'],'variant',j['config']['variant'],'$*','ms/step',round(j['ms_per_step'],3),'cvf_kernel',round(j['roofline']['kernel_ms'],3),'frac',round(j['roofline']['frac'],4),'e2e',round(j['e2e']['ms_per_step'],3))
    else: print(l.rstrip()[:300])
" | tee -a gpurun_out/variants_k.txt
done
