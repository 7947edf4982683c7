/*
 * stereo_oracle.h -- CPU ORACLE for the STEREO_GIF hot path (CVC -> CVF -> WTA).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (primestereomatch_b200/) never links, imports or calls anything in oracle/.
 *
 * It is a plain-C restatement of the reference's pthreads CPU path
 * (/root/reference: src/CVC.cpp, src/CVF.cpp, src/DispSel.cpp, driven the way
 * src/DispEst.cpp:222-270 drives CVC).  The reference itself cannot be compiled
 * in this image (include/ComFunc.h:33-38 needs <CL/cl.h> and <opencv2/opencv.hpp>,
 * neither is installed), so there is no oracle/_ref.  The OpenCV primitives the
 * reference calls (cvtColor, Sobel, boxFilter, multiply, split) are un-vendored
 * third-party code; they are restated here from OpenCV's published algorithm and
 * PINNED against python cv2 4.13.0 by tests/golden/make_golden.py + tests/test_oracle.py
 * (box filter, Sobel, multiply: bit-exact; RGB2GRAY: bit-exact vs this cv2 build).
 *
 * Layouts: images are interleaved 3-channel float (BGR order as cv::imread gives,
 * reference StereoMatch.cpp:557) with an element step; volumes are [d][y][x] float,
 * contiguous; disparity maps are u8 [y][x].
 */
#ifndef STEREO_ORACLE_H
#define STEREO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_GIF_R_WIN 8          /* reference include/ComFunc.h:49 */
#define ORC_GIF_EPS 0.0001f      /* reference include/ComFunc.h:50 */
#define ORC_MAX_CPU_THREADS 8    /* reference include/ComFunc.h:52 */

/* StereoMatch.cpp:193-197: convertTo(CV_32F, 1/255.0f) */
void orc_u8_to_f32(const uint8_t* src, float* dst, size_t n);

/* CVC.cpp:43  cvtColor(Img, GrdX, CV_RGB2GRAY) on a 3-channel f32 image.
 * gray_mode 0: cv2-4.13(+IPP/AVX2 dispatch) rounding  fma(c2,.114f, fma(c0,.299f, c1*.587f))
 * gray_mode 1: plain left-to-right  (c0*.299f + c1*.587f) + c2*.114f  (non-FMA OpenCV builds) */
void orc_rgb2gray(const float* img3, int W, int H, float* gray, int gray_mode);

/* CVC.cpp:44  Sobel(GrdX, GrdX, CV_32F, 1, 0, 1): G[x+1]-G[x-1], BORDER_REFLECT_101 */
void orc_sobel_x(const float* gray, int W, int H, float* grdx);

/* CVC.cpp:41-46 */
void orc_cvc_preprocess(const float* img3, int W, int H, float* grdx, int gray_mode);

/* CVC.cpp:122-149 (== buildCV_left_thread :48-83): one slice of the left volume */
void orc_buildcv_left(const float* lImg, const float* rImg, const float* lGrd, const float* rGrd,
                      int W, int H, int d, float* cost);
/* CVC.cpp:151-179 (== buildCV_right_thread :85-120); called with swapped images
 * exactly as DispEst.cpp:217,260 do */
void orc_buildcv_right(const float* lImg, const float* rImg, const float* lGrd, const float* rGrd,
                       int W, int H, int d, float* cost);

/* cv::boxFilter(src, dst, -1, Size(8,8)) as used at CVF.cpp:50,63,82,88,158,160:
 * normalised, anchor (4,4), BORDER_REFLECT_101, double accumulation in OpenCV's
 * RowSum<float,double> / ColumnSum<double,float> order. src may equal dst. */
void orc_box8(const float* src, int W, int H, float* dst);

/* CVF.cpp:44-70  planes are H*W each: rgb[3], mean[3], var[6] (contiguous, plane-major) */
void orc_cvf_preprocess(const float* img3, int W, int H, float* rgb, float* mean, float* var);

/* CVF.cpp:72-165 GuidedFilter_cv on one slice, in place (CVF.cpp:38).
 * a_out (3 planes) / b_out (1 plane) may be NULL; when given they receive the
 * coefficient planes (a: CVF.cpp:131-146, b: mean_p after CVF.cpp:152-155). */
void orc_guided_filter(const float* rgb, const float* mean, const float* var, int W, int H,
                       float* p_inout, float* a_out, float* b_out);

/* DispSel.cpp:83-109 */
void orc_wta(const float* vol, int W, int H, int D, uint8_t* disp);

/* DispEst.cpp:222-270 CostConst_CPU: preprocess both, then left volume in batches of
 * `threads` pthreads (one per d), then right volume likewise.  lGrd/rGrd (H*W) are outputs. */
int orc_cost_const(const float* lImg, const float* rImg, int W, int H, int D, int threads,
                   int gray_mode, float* lGrd, float* rGrd, float* lVol, float* rVol);

/* The CostFilter_CPU that DispEst.h:42 declares but never defines, reconstructed on the
 * model of CostConst_CPU with CVF::preprocess (CVF.cpp:44) + CVF::filterCV_thread (CVF.cpp:28). */
int orc_cost_filter(const float* lImg, const float* rImg, int W, int H, int D, int threads,
                    float* lVol, float* rVol);

/* DispEst.cpp:311-321 DispSelect_CPU */
int orc_disp_select(const float* lVol, const float* rVol, int W, int H, int D,
                    uint8_t* lDis, uint8_t* rDis);

/* Whole path with get_rt()-style per-stage timers (ComFunc.h:67-71); times_ms[3] = cvc,cvf,wta */
int orc_pipeline(const float* lImg, const float* rImg, int W, int H, int D, int threads,
                 int gray_mode, float* lVol, float* rVol, uint8_t* lDis, uint8_t* rDis,
                 double* times_ms);

/* Fast Guided Filter branch (src/fastguidedfilter.cpp via DispEst::CostFilter_FGF, DispEst.cpp:281-296), s = sub-sampling
 * rate (2, 4, 8), volumes filtered in place.  Pinned against cv2 (IPP off) by tests/golden/make_golden_fgf.py. */
void orc_box_k(const float* src, int W, int H, int K, float* dst);
int orc_cost_filter_fgf(const float* lImg, const float* rImg, int W, int H, int D, int threads, int s, float* lVol, float* rVol);

/* Post-processing (PP::processDM live code, src/PP.cpp:414-422 -> JointWMF::filter): see stereo_oracle.c.
 * PARITY UNPINNED for natural images (cv::kmeans feature clustering is RNG-seeded); this is the un-clustered filter. */
void orc_pp_weight_lut(uint32_t* lut /* 3*63*63 + 1 entries, 2^-22 fixed point */);
void orc_wmf(const uint8_t* disp, const uint8_t* img8_bgr, int W, int H, int r, uint8_t* out);
void orc_f32_to_u8x255(const float* src, uint8_t* dst, size_t n);

#ifdef __cplusplus
}
#endif
#endif
