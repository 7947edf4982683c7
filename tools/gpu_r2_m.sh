#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_knockout.sh
timeout 600 python -m pytest tests/test_fgf.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_m.log 2>&1; tail -2 gpurun_out/pytest_m.log
python /dev/stdin <<'PY' 2>&1 | tee gpurun_out/fgf_time.txt
import numpy as np, sys
sys.path.insert(0,'.')
from primestereomatch_b200 import DispEst, capi, synth
W,H,D=1920,1080,128
l8,r8,_=synth.stereo_pair_u8(W,H,D)
with DispEst(l8,r8,D) as de:
    for s in (4,2,8):
        de.setSubsampleRate(s)
        ms=[]
        for _ in range(4):
            de.CostConst_GPU(); de.CostFilter_FGF_GPU(); de.DispSelect_GPU()
            ms.append(de.stage_ms(2))
        print(f"FGF s={s}: {np.mean(ms[1:]):.3f} ms (both views, C4)")
PY
