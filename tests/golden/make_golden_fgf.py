"""Golden vectors for the Fast Guided Filter branch (reference src/fastguidedfilter.cpp, driven by
DispEst::CostFilter_FGF, src/DispEst.cpp:281-296) -- TEST INFRASTRUCTURE.

The reference code is followed call by call through the REAL cv2 primitives (blur, resize, multiply, subtract, add,
addWeighted, divide); cv::MatExpr chains are lowered the way OpenCV's matop.cpp lowers them (every a.mul(b) is a
rounded multiply; "X - Y + eps" with a non-zero scalar becomes addWeighted(X, 1, Y, -1, eps) -- double arithmetic, one
rounding; sums and differences of
products are plain add / subtract in source order).  cv::resize runs with IPP DISABLED: with IPP, INTER_LINEAR goes to
Intel's routine whose arithmetic differs from OpenCV's own by up to 1e-4 on these planes (build-dependent, like the
RGB2GRAY rounding); the C port (oracle/stereo_oracle.c::orc_fgf_*) follows OpenCV's own implementation.

Writes tests/golden/golden_fgf.json: sha256 of filtered slices and of the WTA maps for Teddy (s = 4, 2, 8).
Run: python tests/golden/make_golden_fgf.py"""
import hashlib
import json
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
f32 = np.float32
GIF_R_WIN, GIF_EPS = 8, float(f32(0.0001))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def box(I, r):
    return cv2.blur(I, (r, r))          # fastguidedfilter.cpp:5-10


class FastGuidedFilterColor:            # fastguidedfilter.cpp:121-198
    def __init__(self, I, r, eps, s):
        self.s = s
        self.r = 2 * (r // s) + 1       # :206-208
        self.orig = cv2.split(I)        # :131
        H, W = I.shape[:2]
        Is = cv2.resize(I, (W // s, H // s), interpolation=cv2.INTER_NEAREST)   # :132
        self.Ic = cv2.split(Is)
        r_ = self.r
        self.m = [box(c, r_) for c in self.Ic]                                   # :135-137
        mul, sub, add = cv2.multiply, cv2.subtract, cv2.add

        def var(a, b, with_eps):
            t = box(mul(self.Ic[a], self.Ic[b]), r_)
            mm = mul(self.m[a], self.m[b])
            return cv2.addWeighted(t, 1.0, mm, -1.0, eps) if with_eps else sub(t, mm)   # :144-149 (MatOp_AddEx -> addWeighted)
        rr, rg, rb = var(0, 0, True), var(0, 1, False), var(0, 2, False)
        gg, gb, bb = var(1, 1, True), var(1, 2, False), var(2, 2, True)
        inv = {
            "rr": sub(mul(gg, bb), mul(gb, gb)), "rg": sub(mul(gb, rb), mul(rg, bb)), "rb": sub(mul(rg, gb), mul(gg, rb)),
            "gg": sub(mul(rr, bb), mul(rb, rb)), "gb": sub(mul(rb, rg), mul(rr, gb)), "bb": sub(mul(rr, gg), mul(rg, rg)),
        }                                                                         # :152-157
        det = add(add(mul(inv["rr"], rr), mul(inv["rg"], rg)), mul(inv["rb"], rb))  # :159
        self.inv = {k: cv2.divide(v, det) for k, v in inv.items()}               # :161-166

    def filter(self, p):
        H, W = p.shape
        s, r_ = self.s, self.r
        mul, sub, add = cv2.multiply, cv2.subtract, cv2.add
        p2 = cv2.resize(p, (W // s, H // s), interpolation=cv2.INTER_NEAREST)    # :69
        mp = box(p2, r_)
        mIp = [box(mul(c, p2), r_) for c in self.Ic]
        cov = [sub(mIp[k], mul(self.m[k], mp)) for k in range(3)]                # :178-180
        iv = self.inv
        a_r = add(add(mul(iv["rr"], cov[0]), mul(iv["rg"], cov[1])), mul(iv["rb"], cov[2]))
        a_g = add(add(mul(iv["rg"], cov[0]), mul(iv["gg"], cov[1])), mul(iv["gb"], cov[2]))
        a_b = add(add(mul(iv["rb"], cov[0]), mul(iv["gb"], cov[1])), mul(iv["bb"], cov[2]))
        b = sub(sub(sub(mp, mul(a_r, self.m[0])), mul(a_g, self.m[1])), mul(a_b, self.m[2]))   # :186
        up = lambda x: cv2.resize(box(x, r_), (W, H), interpolation=cv2.INTER_LINEAR)           # :188-195
        ma = [up(a_r), up(a_g), up(a_b)]
        mb = up(b)
        q = add(add(mul(ma[0], self.orig[0]), mul(ma[1], self.orig[1])), mul(ma[2], self.orig[2]))
        return add(q, mb)                                                         # :196


def main():
    cv2.setNumThreads(1)
    cv2.ipp.setUseIPP(False)
    from oracle import oracle as O
    out = {"cv2": cv2.__version__, "ipp": False, "scenes": {}}
    for scene in ("Teddy",):
        s_ = scene.lower()
        l8 = cv2.imread(os.path.join(HERE, f"{s_}_im2.png"), cv2.IMREAD_COLOR)
        r8 = cv2.imread(os.path.join(HERE, f"{s_}_im6.png"), cv2.IMREAD_COLOR)
        l, r = O.u8_to_f32(l8), O.u8_to_f32(r8)
        D = 64
        _, _, lraw, rraw = O.cost_const(l, r, D)
        sc = {}
        for s in (4, 2, 8):
            fl = FastGuidedFilterColor(l, GIF_R_WIN, GIF_EPS, s)
            fr = FastGuidedFilterColor(r, GIF_R_WIN, GIF_EPS, s)
            lf = np.stack([fl.filter(lraw[d]) for d in range(D)])
            rf = np.stack([fr.filter(rraw[d]) for d in range(D)])
            sc[f"s{s}"] = {"lFilt": sha(lf), "rFilt": sha(rf), "l_d1": sha(lf[1]), "l_d20": sha(lf[20]), "l_d63": sha(lf[63]),
                           "lDis": sha(O.wta(lf)), "rDis": sha(O.wta(rf)),
                           "filt_min": float(lf.min()), "filt_max": float(lf.max())}
        out["scenes"][scene] = sc
    with open(os.path.join(HERE, "golden_fgf.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1)[:600])


if __name__ == "__main__":
    main()
