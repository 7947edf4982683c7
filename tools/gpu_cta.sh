#!/bin/bash
for x in 128 64 96 32; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --cta-threads $x 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('cta_threads',$x,'ms/step',round(j['ms_per_step'],3),'cvf_kernel',round(j['roofline']['kernel_ms'],3))
"
done
python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -2
