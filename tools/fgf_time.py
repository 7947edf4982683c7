"""Times the Fast Guided Filter branch (psm_cost_filter_fgf, both views) on the C4 frame for s = 4, 2, 8 and the
whole CVC -> FGF -> WTA frame at s = 4.  Usage: python tools/fgf_time.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primestereomatch_b200 import DispEst, synth  # noqa: E402

W, H, D = 1920, 1080, 128
l8, r8, _ = synth.stereo_pair_u8(W, H, D)
with DispEst(l8, r8, D) as de:
    for s in (4, 2, 8):
        de.setSubsampleRate(s)
        ms, frame = [], []
        for _ in range(5):
            de.CostConst_GPU(); de.CostFilter_FGF_GPU(); de.DispSelect_GPU()
            ms.append(de.stage_ms(2))
            frame.append(de.stage_ms(0) + de.stage_ms(1) + de.stage_ms(2) + de.stage_ms(3))
        print(f"FGF s={s}: {np.mean(ms[1:]):.3f} ms (both views, C4); ingest+CVC+FGF+WTA {np.mean(frame[1:]):.3f} ms")
