#!/bin/bash
mkdir -p gpurun_out
for v in 0 2 3; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --variant $v 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('variant',j['config']['variant'],'ms/step',round(j['ms_per_step'],3),'cvf_kernel',round(j['roofline']['kernel_ms'],3),'frac',round(j['roofline']['frac'],4), j['config']['stage_ms_last_step'])
" | tee -a gpurun_out/variants.txt
done
