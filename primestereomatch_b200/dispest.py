"""Host-side mirror of the reference's `DispEst` facade for the GPU compute mode.

Mirrors /root/reference/include/DispEst.h:21-109 (same method names, argument meaning and
return conventions: stage methods return 0 on success) with numpy arrays standing in for
cv::Mat.  Only the `_GPU` stage methods -- the path behind the `m` toggle -- are implemented,
and they are thin calls into the C-ABI (include/prime_stereo_b200.h); the CPU stage methods are
out of scope for this package and raise NotImplementedError.

The C++ twin that a maintainer links into the reference application is
primestereomatch_b200/host/DispEstB200.{h,cpp}; see INTEGRATION.md.
"""
import ctypes as C

import numpy as np

from . import capi

MAX_CPU_THREADS = 8  # reference include/ComFunc.h:52
OCV_DE, OCL_DE = 0, 1  # reference include/ComFunc.h:46-47


def device_count():
    """Role of openCLdevicepoll() (reference src/main.cpp:29): gates the `m` toggle."""
    return capi.lib().psm_device_count()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class DispEst:
    """DispEst(l, r, d, t, ocl) -- reference src/DispEst.cpp:10-143.

    l, r : HxWx3 images, both float32 in [0,1] (CV_32FC3) or both uint8 (CV_8UC3), BGR order.
    d    : maxDis.   t : CPU threads (kept for API compatibility).   ocl : use the GPU path.
    """

    def __init__(self, l, r, d, t=MAX_CPU_THREADS, ocl=True, device=0, d_begin=0, d_count=None):
        l, r = np.asarray(l), np.asarray(r)
        if l.dtype != r.dtype or l.shape != r.shape:
            # reference: "DE: Error - Left & Right images are of different types." then exit(1)
            raise ValueError("DE: Error - Left & Right images are of different types.")
        if l.ndim != 3 or l.shape[2] != 3 or l.dtype not in (np.float32, np.uint8):
            raise ValueError("DispEst expects HxWx3 float32 or uint8 images")
        self.hei, self.wid = int(l.shape[0]), int(l.shape[1])
        self.maxDis = int(d)
        self.threads = int(t)
        self.useOCL = bool(ocl)
        self.subsample_rate = 4
        self.lDisMap = np.zeros((self.hei, self.wid), np.uint8)  # DispEst.cpp:46-47
        self.rDisMap = np.zeros((self.hei, self.wid), np.uint8)
        self._lImg, self._rImg = l, r
        self._ctx = C.c_void_p()
        self._lib = None
        if self.useOCL:
            L = capi.lib()
            self._lib = L
            dc = self.maxDis - d_begin if d_count is None else d_count
            capi.check(L.psm_create_sharded(C.byref(self._ctx), self.wid, self.hei, self.maxDis,
                                            int(d_begin), int(dc), int(device)))
        self.d_begin, self.d_count = d_begin, (self.maxDis - d_begin if d_count is None else d_count)

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if self._lib is not None and self._ctx:
            self._lib.psm_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- reference public methods -----------------------------------------------------------
    def setInputImages(self, leftImg, rightImg):
        """DispEst.cpp:164-170 (asserts equal types, keeps the headers)."""
        leftImg, rightImg = np.asarray(leftImg), np.asarray(rightImg)
        assert leftImg.dtype == rightImg.dtype
        self._lImg, self._rImg = leftImg, rightImg
        return 0

    def setThreads(self, newThreads):
        """DispEst.cpp:172-179."""
        if newThreads > MAX_CPU_THREADS:
            return -1
        self.threads = int(newThreads)
        return 0

    def setSubsampleRate(self, newRate):
        self.subsample_rate = int(newRate)

    def set_option(self, key, value):
        capi.check(self._lib.psm_set_option(self._ctx, int(key), int(value)), self._ctx)

    def _need_gpu(self):
        if not self.useOCL or self._lib is None:
            raise RuntimeError("DispEst was constructed with ocl=False: the GPU stages are unavailable")

    def CostConst_GPU(self):
        """DispEst.cpp:272-276: upload both images, build both raw cost volumes on the device."""
        self._need_gpu()
        l = np.ascontiguousarray(self._lImg)
        r = np.ascontiguousarray(self._rImg)
        if l.shape != (self.hei, self.wid, 3) or r.shape != l.shape:
            raise ValueError("input image size differs from the size DispEst was constructed with")
        if l.dtype == np.uint8:
            capi.check(self._lib.psm_set_images_u8(self._ctx, _ptr(l), l.strides[0], _ptr(r), r.strides[0]), self._ctx)
        else:
            capi.check(self._lib.psm_set_images(self._ctx, _ptr(l), l.strides[0], _ptr(r), r.strides[0]), self._ctx)
        capi.check(self._lib.psm_cost_const(self._ctx), self._ctx)
        return 0

    def CostFilter_GPU(self):
        """DispEst.cpp:299-308."""
        self._need_gpu()
        capi.check(self._lib.psm_cost_filter(self._ctx), self._ctx)
        return 0

    def CostFilter_FGF_GPU(self):
        """Device twin of DispEst::CostFilter_FGF (DispEst.cpp:281-296): Fast Guided Filter at the sub-sampling rate
        set by setSubsampleRate (default 4, like StereoMatch.cpp:33)."""
        self._need_gpu()
        capi.check(self._lib.psm_cost_filter_fgf(self._ctx, int(self.subsample_rate)), self._ctx)
        return 0

    def DispSelect_GPU(self):
        """DispEst.cpp:323-328: WTA, results land in lDisMap / rDisMap."""
        self._need_gpu()
        capi.check(self._lib.psm_disp_select(self._ctx, _ptr(self.lDisMap), self.lDisMap.strides[0],
                                             _ptr(self.rDisMap), self.rDisMap.strides[0]), self._ctx)
        return 0

    def PostProcess_GPU(self):
        """DispEst.cpp:338-344 -> PP::processDM (PP.cpp:402-425): joint weighted-median filter of both maps; the
        reference runs it on the CPU even in GPU mode, here it is a device stage.  Results replace lDisMap / rDisMap."""
        self._need_gpu()
        capi.check(self._lib.psm_post_process(self._ctx, _ptr(self.lDisMap), self.lDisMap.strides[0],
                                              _ptr(self.rDisMap), self.rDisMap.strides[0]), self._ctx)
        return 0

    def CostConst(self):
        raise NotImplementedError("CPU stage (reference src/CVC.cpp) is not part of this package")

    CostConst_CPU = CostFilter = CostFilter_CPU = CostFilter_FGF = DispSelect_CPU = PostProcess_CPU = CostConst

    # -- parity / debugging helpers (role of DispEst::printCV, DispEst.cpp:181-194) ------------
    def read_cost_slice(self, view, d):
        out = np.empty((self.hei, self.wid), np.float32)
        capi.check(self._lib.psm_read_cost_slice(self._ctx, view, d, _ptr(out), out.strides[0]), self._ctx)
        return out

    def read_cost_volume(self, view):
        return np.stack([self.read_cost_slice(view, d) for d in range(self.d_begin, self.d_begin + self.d_count)])

    def write_cost_slice(self, view, d, src):
        src = np.ascontiguousarray(src, np.float32)
        capi.check(self._lib.psm_write_cost_slice(self._ctx, view, d, _ptr(src), src.strides[0]), self._ctx)

    def read_guide_plane(self, view, plane):
        out = np.empty((self.hei, self.wid), np.float32)
        capi.check(self._lib.psm_read_guide_plane(self._ctx, view, plane, _ptr(out), out.strides[0]), self._ctx)
        return out

    def read_ab_slice(self, view, d):
        a = np.empty((3, self.hei, self.wid), np.float32)
        b = np.empty((self.hei, self.wid), np.float32)
        capi.check(self._lib.psm_read_ab_slice(self._ctx, view, d, _ptr(a), _ptr(b)), self._ctx)
        return a, b

    def stage_ms(self, stage):
        ms = C.c_float()
        capi.check(self._lib.psm_stage_ms(self._ctx, stage, C.byref(ms)), self._ctx)
        return ms.value

    def launch_count(self):
        n = C.c_uint64()
        capi.check(self._lib.psm_launch_count(self._ctx, C.byref(n)), self._ctx)
        return n.value

    def sync(self):
        capi.check(self._lib.psm_sync(self._ctx), self._ctx)

    @property
    def handle(self):
        return self._ctx
