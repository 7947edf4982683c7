#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modes.py tests/test_fgf.py tests/test_pp.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_i.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_i.log
tail -5 gpurun_out/pytest_i.log
python tools/pp_time.py C4 2>&1 | tail -2 | tee gpurun_out/pp_time.txt
rm -f gpurun_out/variants_i.txt
for cfg in "1 0" "1 10 --cta-threads 128" "1 11" "1 10 --cta-threads 96" "0 0" "0 11" "0 10 --cta-threads 128"; do
  set -- $cfg
  m=$1; v=$2; shift; shift
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --cvf-mode $m --variant $v "$@" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('mode',j['config']['cvf_mode'],'variant',j['config']['variant'],'$*','ms/step',round(j['ms_per_step'],3),'cvf_kernel',round(j['roofline']['kernel_ms'],3),'frac',round(j['roofline']['frac'],4),'e2e',round(j['e2e']['ms_per_step'],3))
    else: print(l.rstrip()[:300])
" | tee -a gpurun_out/variants_i.txt
done
# FGF timing at C4
python - <<'PY' 2>&1 | tee gpurun_out/fgf_time.txt
import numpy as np, sys
sys.path.insert(0,'.')
from primestereomatch_b200 import DispEst, capi, synth
W,H,D=1920,1080,128
l8,r8,_=synth.stereo_pair_u8(W,H,D)
with DispEst(l8,r8,D) as de:
    for s in (4,2,8):
        de.setSubsampleRate(s)
        ms=[]
        for _ in range(5):
            de.CostConst_GPU(); de.CostFilter_FGF_GPU(); de.DispSelect_GPU()
            ms.append(de.stage_ms(2))
        print(f"FGF s={s}: {np.mean(ms[1:]):.3f} ms (both views, C4)")
PY
