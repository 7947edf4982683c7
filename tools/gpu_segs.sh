#!/bin/bash
for r in 1080 540 360 270 216 180 135; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --seg-rows $r 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('seg_rows',$r,'ms/step',round(j['ms_per_step'],3),'cvf_kernel',round(j['roofline']['kernel_ms'],3))
"
done
