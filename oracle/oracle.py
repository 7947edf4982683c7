"""ctypes loader for the CPU oracle (oracle/libstereo_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build():
    """Compile the oracle with its Makefile (gcc, seconds)."""
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libstereo_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_u8_to_f32.argtypes = [_u8p, _f32p, C.c_size_t]
        L.orc_rgb2gray.argtypes = [_f32p, C.c_int, C.c_int, _f32p, C.c_int]
        L.orc_sobel_x.argtypes = [_f32p, C.c_int, C.c_int, _f32p]
        L.orc_cvc_preprocess.argtypes = [_f32p, C.c_int, C.c_int, _f32p, C.c_int]
        for fn in (L.orc_buildcv_left, L.orc_buildcv_right):
            fn.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, _f32p]
        L.orc_box8.argtypes = [_f32p, C.c_int, C.c_int, _f32p]
        L.orc_cvf_preprocess.argtypes = [_f32p, C.c_int, C.c_int, _f32p, _f32p, _f32p]
        L.orc_guided_filter.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, _f32p,
                                        C.c_void_p, C.c_void_p]
        L.orc_wta.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _u8p]
        L.orc_cost_const.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     _f32p, _f32p, _f32p, _f32p]
        L.orc_cost_filter.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                      _f32p, _f32p]
        L.orc_disp_select.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, _u8p, _u8p]
        L.orc_pipeline.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   _f32p, _f32p, _u8p, _u8p, _f64p]
        L.orc_cost_filter_fgf.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p]
        L.orc_cost_filter_fgf.restype = C.c_int
        L.orc_wmf.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, _u8p]
        L.orc_f32_to_u8x255.argtypes = [_f32p, _u8p, C.c_size_t]
        for name in ("orc_cost_const", "orc_cost_filter", "orc_disp_select", "orc_pipeline"):
            getattr(L, name).restype = C.c_int
        _LIB = L
    return _LIB


def _c(a, dt=np.float32):
    return np.ascontiguousarray(a, dtype=dt)


def u8_to_f32(img_u8):
    src = _c(img_u8, np.uint8)
    dst = np.empty(src.shape, np.float32)
    lib().orc_u8_to_f32(src.reshape(-1), dst.reshape(-1), src.size)
    return dst


def rgb2gray(img3, gray_mode=0):
    img3 = _c(img3)
    H, W, _ = img3.shape
    out = np.empty((H, W), np.float32)
    lib().orc_rgb2gray(img3, W, H, out, gray_mode)
    return out


def sobel_x(gray):
    gray = _c(gray)
    H, W = gray.shape
    out = np.empty((H, W), np.float32)
    lib().orc_sobel_x(gray, W, H, out)
    return out


def cvc_preprocess(img3, gray_mode=0):
    img3 = _c(img3)
    H, W, _ = img3.shape
    out = np.empty((H, W), np.float32)
    lib().orc_cvc_preprocess(img3, W, H, out, gray_mode)
    return out


def box8(src):
    src = _c(src)
    H, W = src.shape
    out = np.empty((H, W), np.float32)
    lib().orc_box8(src, W, H, out)
    return out


def cvf_preprocess(img3):
    """-> rgb[3,H,W], mean[3,H,W], var[6,H,W]"""
    img3 = _c(img3)
    H, W, _ = img3.shape
    rgb = np.empty((3, H, W), np.float32)
    mean = np.empty((3, H, W), np.float32)
    var = np.empty((6, H, W), np.float32)
    lib().orc_cvf_preprocess(img3, W, H, rgb, mean, var)
    return rgb, mean, var


def guided_filter(rgb, mean, var, p, want_ab=False):
    """-> q  (or (q, a[3,H,W], b[H,W]) with want_ab)"""
    p = _c(p).copy()
    H, W = p.shape
    if want_ab:
        a = np.empty((3, H, W), np.float32)
        b = np.empty((H, W), np.float32)
        lib().orc_guided_filter(_c(rgb), _c(mean), _c(var), W, H, p,
                                a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
        return p, a, b
    lib().orc_guided_filter(_c(rgb), _c(mean), _c(var), W, H, p, None, None)
    return p


def wta(vol):
    vol = _c(vol)
    D, H, W = vol.shape
    out = np.empty((H, W), np.uint8)
    lib().orc_wta(vol, W, H, D, out)
    return out


def cost_const(l, r, D, threads=8, gray_mode=0):
    """-> lGrd, rGrd, lVol[D,H,W], rVol[D,H,W]  (raw cost volumes)"""
    l, r = _c(l), _c(r)
    H, W, _ = l.shape
    lg = np.empty((H, W), np.float32)
    rg = np.empty((H, W), np.float32)
    lv = np.empty((D, H, W), np.float32)
    rv = np.empty((D, H, W), np.float32)
    rc = lib().orc_cost_const(l, r, W, H, D, threads, gray_mode, lg, rg, lv, rv)
    assert rc == 0
    return lg, rg, lv, rv


def cost_filter(l, r, lv, rv, threads=8):
    """filters copies of the volumes -> lVolF, rVolF"""
    l, r = _c(l), _c(r)
    lv, rv = _c(lv).copy(), _c(rv).copy()
    D, H, W = lv.shape
    rc = lib().orc_cost_filter(l, r, W, H, D, threads, lv, rv)
    assert rc == 0
    return lv, rv


def disp_select(lv, rv):
    lv, rv = _c(lv), _c(rv)
    D, H, W = lv.shape
    ld = np.empty((H, W), np.uint8)
    rd = np.empty((H, W), np.uint8)
    lib().orc_disp_select(lv, rv, W, H, D, ld, rd)
    return ld, rd


def pipeline(l, r, D, threads=8, gray_mode=0, keep_volumes=False):
    """Full CVC->CVF->WTA.  -> dict(lDis, rDis, times_ms=[cvc,cvf,wta], [lVol, rVol])"""
    l, r = _c(l), _c(r)
    H, W, _ = l.shape
    lv = np.empty((D, H, W), np.float32)
    rv = np.empty((D, H, W), np.float32)
    ld = np.empty((H, W), np.uint8)
    rd = np.empty((H, W), np.uint8)
    t = np.zeros(3, np.float64)
    rc = lib().orc_pipeline(l, r, W, H, D, threads, gray_mode, lv, rv, ld, rd, t)
    assert rc == 0
    out = {"lDis": ld, "rDis": rd, "times_ms": t.tolist()}
    if keep_volumes:
        out["lVol"], out["rVol"] = lv, rv
    return out


def post_process(img3_f32, disp, r=9):
    """PP::processDM's live code for one view (src/PP.cpp:414-422), un-clustered restatement (see stereo_oracle.c)."""
    img3 = _c(img3_f32)
    disp = _c(disp, np.uint8)
    H, W = disp.shape
    img8 = np.empty(img3.shape, np.uint8)
    lib().orc_f32_to_u8x255(img3.reshape(-1), img8.reshape(-1), img3.size)
    out = np.empty((H, W), np.uint8)
    lib().orc_wmf(disp, img8, W, H, r, out)
    return out


def cost_filter_fgf(l, r, lv, rv, s=4, threads=8):
    """DispEst::CostFilter_FGF (Fast Guided Filter, sub-sampling rate s) on copies of the volumes -> lVolF, rVolF"""
    l, r = _c(l), _c(r)
    lv, rv = _c(lv).copy(), _c(rv).copy()
    D, H, W = lv.shape
    rc = lib().orc_cost_filter_fgf(l, r, W, H, D, threads, s, lv, rv)
    assert rc == 0
    return lv, rv
