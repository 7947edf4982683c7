#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fgf.py tests/test_pp.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_l.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_l.log
tail -4 gpurun_out/pytest_l.log
python tools/pp_time.py C4 2>&1 | tail -2 | tee gpurun_out/pp_time.txt
cat > /tmp/fgf_time.py <<'PY'
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
python /tmp/fgf_time.py 2>&1 | tee gpurun_out/fgf_time.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fgf_|pp_" -c 60 --csv --log-file gpurun_out/fgf_launches.csv python /tmp/fgf_time.py > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/fgf_launches.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hdr=rows[hi]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
c=collections.OrderedDict()
for r in rows[hi+2:]:
    if len(r)>vi: c.setdefault(r[ki].split('(')[0],[]).append(float(r[vi].replace(',','')))
for k,v in c.items(): print(k, [round(x/1e3,1) for x in v[:12]])
PY
