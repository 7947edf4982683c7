"""Times the device post-processing stage (psm_post_process_device) on a BASELINE-size frame and checks a row band
against the CPU restatement.  Usage: python tools/pp_time.py [C3|C4]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primestereomatch_b200 import DispEst, capi, synth  # noqa: E402

W, H, D = {"C3": (1280, 720, 64), "C4": (1920, 1080, 128)}[sys.argv[1] if len(sys.argv) > 1 else "C4"]
l8, r8, _ = synth.stereo_pair_u8(W, H, D)
rng = np.random.default_rng(3)
r8 = np.clip(r8.astype(np.int16) + rng.integers(-6, 7, r8.shape), 0, 255).astype(np.uint8)
with DispEst(l8, r8, D) as de:
    de.CostConst_GPU(); de.CostFilter_GPU(); de.DispSelect_GPU()
    ld = de.lDisMap.copy()
    L = capi.lib()
    for _ in range(3):
        capi.check(L.psm_post_process_device(de.handle), de.handle)
    de.sync()
    ms = []
    for _ in range(5):
        capi.check(L.psm_post_process_device(de.handle), de.handle)
        ms.append(de.stage_ms(5))
    de.PostProcess_GPU()
    lpp = de.lDisMap.copy()
print(f"post-process (both views) {W}x{H}: {np.mean(ms):.3f} ms (min {min(ms):.3f}); pixels changed by the filter: {int((lpp != ld).sum())}")
from oracle import oracle as O  # noqa: E402  (checker)
y0, y1 = H // 2 - 20, H // 2 + 20
band = slice(y0 - 9, y1 + 9)
t0 = time.time()
want = O.post_process(O.u8_to_f32(l8[band]), ld[band])[9:-9]
print("band check vs CPU restatement:", "OK" if np.array_equal(lpp[y0:y1], want) else "MISMATCH", f"(CPU {time.time() - t0:.2f} s for {y1 - y0 + 18} rows)")
