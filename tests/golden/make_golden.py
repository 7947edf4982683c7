#!/usr/bin/env python
"""Generate the golden fixtures that PIN the CPU oracle (run in the build container only).

The reference (PRiMEStereoMatch @ /root/reference) cannot be compiled here (no OpenCV C++
headers, no CL/cl.h), and it ships no tests or expected outputs.  What CAN be done here is to
run the reference's pthreads path *call by call through the real OpenCV primitives* that it
uses (python cv2 4.13.0 is installed): cv2.cvtColor / cv2.Sobel / cv2.boxFilter / cv2.multiply
in exactly the order of CVC.cpp:41-46, CVF.cpp:44-165, with the hand-written loops
(CVC.cpp:18-39, CVF.cpp:102-149, DispSel.cpp:83-109) restated in numpy float32/float64.
That cv2-driven run is the generator of these fixtures; tests/test_oracle.py then asserts the
plain-C oracle reproduces them bit for bit.

Outputs (all small, committed):
  tests/golden/<scene>_{im2,im6}.png      Middlebury inputs (copied: /root/reference is absent on the GPU box)
  tests/golden/<scene>_disp2.png, _occl.png  ground truth + non-occlusion mask for the %BP metric
  tests/golden/<scene>_{lDis,rDis}.png    golden u8 disparity maps (raw disparity index)
  tests/golden/golden.json                sha256 of every intermediate, %BP, primitive pins
  tests/golden/<scene>_crops.npz          small crops of raw cost / a / b / q at a few d
"""
import hashlib
import json
import os
import shutil
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/data"
D = 64
f32 = np.float32
EPS = f32(0.0001)  # GIF_EPS ComFunc.h:50


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_f32(path):
    u8 = cv2.imread(path, cv2.IMREAD_COLOR)  # BGR, StereoMatch.cpp:557
    return u8, u8.astype(f32) * f32(1 / f32(255.0))  # StereoMatch.cpp:193-197


def cvc_preprocess(img):
    g = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY)  # CVC.cpp:43 (on a BGR image, as the reference does)
    return cv2.Sobel(g, cv2.CV_32F, 1, 0, ksize=1)  # CVC.cpp:44


def cost4(lC, rC, lG, rG):
    """CVC.cpp:18-27, float32 left-to-right."""
    clr = (np.abs(lC[..., 0] - rC[..., 0]) + np.abs(lC[..., 1] - rC[..., 1])) + np.abs(lC[..., 2] - rC[..., 2])
    grd = np.abs(lG - rG)
    return f32(0.9) * clr + (f32(1) - f32(0.9)) * grd


def cost2(lC, lG):
    """CVC.cpp:30-39: BC_32F is a double literal -> double intermediates, one rounding."""
    lc = lC.astype(np.float64)
    clr = ((np.abs(lc[..., 0] - 1.0) + np.abs(lc[..., 1] - 1.0)) + np.abs(lc[..., 2] - 1.0)).astype(f32)
    grd = np.abs(lG.astype(np.float64) - 1.0).astype(f32)
    return f32(0.9) * clr + (f32(1) - f32(0.9)) * grd


def build_left(l, r, lg, rg, d):
    H, W, _ = l.shape
    c = np.empty((H, W), f32)
    c[:, d:] = cost4(l[:, d:], r[:, : W - d], lg[:, d:], rg[:, : W - d])
    c[:, :d] = cost2(l[:, :d], lg[:, :d])
    return c


def build_right(l, r, lg, rg, d):
    """CVC.cpp:151-179 with the caller's swapped arguments (DispEst.cpp:260)."""
    H, W, _ = l.shape
    c = np.empty((H, W), f32)
    c[:, : W - d] = cost4(l[:, : W - d], r[:, d:], lg[:, : W - d], rg[:, d:])
    c[:, W - d:] = cost2(l[:, W - d:], lg[:, W - d:])
    return c


def box(p):
    return cv2.boxFilter(p, -1, (8, 8))  # CVF.cpp:46 Size(GIF_R_WIN, GIF_R_WIN)


def cvf_preprocess(img):
    rgb = cv2.split(img)
    mean = [box(c) for c in rgb]
    var = []
    for c in range(3):
        for cp in range(c, 3):
            tmp = cv2.multiply(rgb[c], rgb[cp])
            v = box(tmp)
            tmp = cv2.multiply(mean[c], mean[cp])
            var.append(v - tmp)
    return rgb, mean, var


def guided(rgb, mean_I, var_I, p):
    mean_p = box(p)
    mean_Ip = [box(cv2.multiply(rgb[c], p)) for c in range(3)]
    cov = [mean_Ip[c] - cv2.multiply(mean_I[c], mean_p) for c in range(3)]
    c0, c1, c2 = cov
    a11 = var_I[0] + EPS; a12 = var_I[1]; a13 = var_I[2]
    a21 = var_I[1]; a22 = var_I[3] + EPS; a23 = var_I[4]
    a31 = var_I[2]; a32 = var_I[4]; a33 = var_I[5] + EPS
    DET = (a11 * (a33 * a22 - a32 * a23) - a21 * (a33 * a12 - a32 * a13)) + a31 * (a23 * a12 - a22 * a13)
    DET = f32(1) / DET
    a0 = DET * ((c0 * (a33 * a22 - a32 * a23) + c1 * (a31 * a23 - a33 * a21)) + c2 * (a32 * a21 - a31 * a22))
    a1 = DET * ((c0 * (a32 * a13 - a33 * a12) + c1 * (a33 * a11 - a31 * a13)) + c2 * (a31 * a12 - a32 * a11))
    a2 = DET * ((c0 * (a23 * a12 - a22 * a13) + c1 * (a21 * a13 - a23 * a11)) + c2 * (a22 * a11 - a21 * a12))
    a = [a0, a1, a2]
    b = mean_p.copy()
    for c in range(3):
        b = b - cv2.multiply(a[c], mean_I[c])
    q = box(b)
    for c in range(3):
        q = q + cv2.multiply(box(a[c]), rgb[c])
    return q, a, b


def bad_pixels(lDis, gt, occl, scale):
    """StereoMatch.cpp:275-311, MASK_NONOCC."""
    disp = cv2.convertScaleAbs(lDis, alpha=scale)  # :248 convertTo(CV_8U, scale_factor)
    e = cv2.absdiff(disp, gt)
    e[:, : D + 1] = 0
    thr = 4 * (127 // D)  # error_threshold*(CHAR_MAX/maxDis)
    _, e = cv2.threshold(e, thr, 255, cv2.THRESH_TOZERO)
    e = cv2.multiply(e, occl, scale=1 / 255.0)
    return float(np.count_nonzero(e)) * 100.0 / float(gt.size)


def main():
    cv2.setNumThreads(1)
    assert np.all(np.isfinite([1.0]))
    out = {"cv2": cv2.__version__, "D": D, "scenes": {}, "primitives": {}}

    # ---- primitive pins on seeded random planes (checked again live in tests when cv2 imports)
    rng = np.random.default_rng(20260924)
    plane = (rng.standard_normal((61, 83)) * np.exp(rng.uniform(-8, 3, (61, 83)))).astype(f32)
    img = rng.random((37, 53, 3), dtype=f32)
    out["primitives"]["box8_61x83"] = sha(box(plane))
    out["primitives"]["gray_37x53"] = sha(cv2.cvtColor(img, cv2.COLOR_RGB2GRAY))
    out["primitives"]["sobel_37x53"] = sha(cvc_preprocess(img))
    np.savez_compressed(os.path.join(HERE, "primitives.npz"), plane=plane, img=img,
                        box=box(plane), gray=cv2.cvtColor(img, cv2.COLOR_RGB2GRAY), sobel=cvc_preprocess(img))

    for scene in ("Cones", "Teddy"):
        s = scene.lower()
        for name in ("im2", "im6", "disp2", "occl"):
            shutil.copyfile(f"{REF}/{scene}/{name}.png", os.path.join(HERE, f"{s}_{name}.png"))
        _, l = load_f32(f"{REF}/{scene}/im2.png")
        _, r = load_f32(f"{REF}/{scene}/im6.png")
        H, W, _ = l.shape
        lg, rg = cvc_preprocess(l), cvc_preprocess(r)
        lraw = np.stack([build_left(l, r, lg, rg, d) for d in range(D)])
        rraw = np.stack([build_right(r, l, rg, lg, d) for d in range(D)])
        lgd = cvf_preprocess(l)
        rgd = cvf_preprocess(r)
        lf = np.empty_like(lraw); rf = np.empty_like(rraw)
        crops = {}
        ys, xs = slice(96, 128), slice(180, 244)
        for d in range(D):
            q, a, b = guided(*lgd, lraw[d])
            lf[d] = q
            if d in (1, 20, 63):
                crops[f"l_raw_d{d}"] = lraw[d][ys, xs]
                crops[f"l_a_d{d}"] = np.stack(a)[:, ys, xs]
                crops[f"l_b_d{d}"] = b[ys, xs]
                crops[f"l_q_d{d}"] = q[ys, xs]
                crops[f"l_a_top_d{d}"] = np.stack(a)[:, :12, :24]  # image corner: reflect-101 paths
                crops[f"l_q_top_d{d}"] = q[:12, :24]
                out["scenes"].setdefault(scene, {})[f"l_a_d{d}"] = sha(np.stack(a))
                out["scenes"][scene][f"l_b_d{d}"] = sha(b)
            rf[d] = guided(*rgd, rraw[d])[0]
        lDis = (np.argmin(lf[1:], axis=0) + 1).astype(np.uint8)  # DispSel.cpp:93-102 (first min = lowest d)
        rDis = (np.argmin(rf[1:], axis=0) + 1).astype(np.uint8)
        cv2.imwrite(os.path.join(HERE, f"{s}_lDis.png"), lDis)
        cv2.imwrite(os.path.join(HERE, f"{s}_rDis.png"), rDis)
        np.savez_compressed(os.path.join(HERE, f"{s}_crops.npz"), **crops)
        gt = cv2.imread(f"{REF}/{scene}/disp2.png", cv2.IMREAD_GRAYSCALE)
        occl = cv2.imread(f"{REF}/{scene}/occl.png", cv2.IMREAD_GRAYSCALE)
        sc = out["scenes"].setdefault(scene, {})
        sc.update({
            "W": W, "H": H,
            "lGrd": sha(lg), "rGrd": sha(rg),
            "lRaw": sha(lraw), "rRaw": sha(rraw),
            "lMean": sha(np.stack(lgd[1])), "lVar": sha(np.stack(lgd[2])),
            "rMean": sha(np.stack(rgd[1])), "rVar": sha(np.stack(rgd[2])),
            "lFilt": sha(lf), "rFilt": sha(rf),
            "lDis": sha(lDis), "rDis": sha(rDis),
            "bp_nonocc_left": bad_pixels(lDis, gt, occl, 4),
            "n_ties_left": int(np.sum(np.sum(lf[1:] == lf[1:].min(axis=0), axis=0) > 1)),
            "filt_min": float(lf.min()), "filt_max": float(lf.max()),
        })
        print(scene, json.dumps(sc, indent=1))
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    sys.exit(main())
