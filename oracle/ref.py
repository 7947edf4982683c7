"""ctypes loader for oracle/_ref/libstereo_ref.so -- the reference's OWN src/CVC.cpp, src/CVF.cpp and
src/DispSel.cpp compiled unmodified against oracle/shim/ (recipe: oracle/Makefile target `ref`).

TEST INFRASTRUCTURE ONLY: used by tests/ to pin the C port (oracle/stereo_oracle.c) to the reference's
compiled code, and by bench.py's CPU legs as the "reference" kind of cpu_baseline.  `available()`
is False where the library was never built (no /root/reference and no prebuilt copy)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libstereo_ref.so")
_LIB = None

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def available():
    if not os.path.exists(PATH) and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-C", _HERE, "-s", "ref"], check=False)
    return os.path.exists(PATH)


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise OSError(f"{PATH} not built (needs /root/reference; `make -C oracle ref`)")
        L = C.CDLL(PATH)
        vp = C.c_void_p
        L.ref_pipeline.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.ref_guided_filter.argtypes = [_f32p, C.c_int, C.c_int, _f32p, _f32p]
        L.ref_wta.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p]
        L.ref_buildcv.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]
        L.ref_post_process.argtypes = [_f32p, _u8p, C.c_int, C.c_int, C.c_int, _u8p, C.POINTER(C.c_int)]
        L.ref_build_info.restype = C.c_char_p
        _LIB = L
    return _LIB


def _c(a, dt=np.float32):
    return np.ascontiguousarray(a, dtype=dt)


def pipeline(l, r, D, threads=8, gray_mode=0, keep=False):
    """Whole CVC -> CVF -> WTA path through the reference's compiled operators.
    -> dict(lDis, rDis, times_ms[, lGrd, rGrd, lRaw, rRaw, lVol, rVol])"""
    l, r = _c(l), _c(r)
    H, W, _ = l.shape
    out = {"lDis": np.empty((H, W), np.uint8), "rDis": np.empty((H, W), np.uint8)}
    t = np.zeros(3, np.float64)
    if keep:
        for k in ("lGrd", "rGrd"):
            out[k] = np.empty((H, W), np.float32)
        for k in ("lRaw", "rRaw", "lVol", "rVol"):
            out[k] = np.empty((D, H, W), np.float32)
    p = lambda k: out[k].ctypes.data_as(C.c_void_p) if k in out else None
    rc = lib().ref_pipeline(l, r, W, H, D, threads, gray_mode, p("lGrd"), p("rGrd"), p("lRaw"), p("rRaw"),
                            p("lVol"), p("rVol"), p("lDis"), p("rDis"), t.ctypes.data_as(C.c_void_p))
    assert rc == 0
    out["times_ms"] = t.tolist()
    return out


def guided_filter(img3, p):
    img3, p = _c(img3), _c(p)
    H, W = p.shape
    q = np.empty((H, W), np.float32)
    lib().ref_guided_filter(img3, W, H, p, q)
    return q


def wta(vol, thread_variant=False, threads=8):
    vol = _c(vol)
    D, H, W = vol.shape
    out = np.empty((H, W), np.uint8)
    lib().ref_wta(vol, W, H, D, int(thread_variant), threads, out)
    return out


def buildcv(l, r, d, right=False, gray_mode=0):
    l, r = _c(l), _c(r)
    H, W, _ = l.shape
    out = np.empty((H, W), np.float32)
    lib().ref_buildcv(l, r, W, H, d, int(right), gray_mode, out)
    return out


def post_process(img3, disp, r=9):
    """One view of PP::processDM's live code (src/PP.cpp:414-422) through the reference's own JointWMF.h.
    -> (filtered u8 map, number of distinct 6-bit colours of the feature image).  The clustering inside is the
    reference's only when that number is <= 256 (oracle/shim kmeans stand-in)."""
    img3, disp = _c(img3), _c(disp, np.uint8)
    H, W = disp.shape
    out = np.empty((H, W), np.uint8)
    n = C.c_int(0)
    lib().ref_post_process(img3, disp, W, H, r, out, C.byref(n))
    return out, n.value
