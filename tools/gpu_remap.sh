#!/bin/bash
for x in 0 1; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --remap $x 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('remap',$x,'ms/step',round(j['ms_per_step'],3),'cvf_kernel',round(j['roofline']['kernel_ms'],3))
"
done
