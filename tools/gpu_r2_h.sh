#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pp.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_pp.log 2>&1
echo "pytest pp exit $?" >> gpurun_out/pytest_pp.log
tail -4 gpurun_out/pytest_pp.log
python tools/pp_time.py C4 2>&1 | tail -3 | tee gpurun_out/pp_time.txt
rm -f gpurun_out/variants_h.txt
for cfg in "0 0 --remap 1" "1 0 --remap 1" "0 0 --seg-rows 360" "1 0 --seg-rows 360" "1 0 --seg-rows 544" "1 0 --cta-threads 128" "1 0"; do
  set -- $cfg
  m=$1; v=$2; shift; shift
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --cvf-mode $m --variant $v "$@" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('mode',j['config']['cvf_mode'],'variant',j['config']['variant'],'$*','ms/step',round(j['ms_per_step'],3),'cvf_kernel',round(j['roofline']['kernel_ms'],3),'frac',round(j['roofline']['frac'],4),'e2e',round(j['e2e']['ms_per_step'],3), j['config']['stage_ms_last_step'])
    else: print(l.rstrip()[:300])
" | tee -a gpurun_out/variants_h.txt
done
