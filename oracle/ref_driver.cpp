// ref_driver.cpp -- C-ABI driver over the reference's OWN compiled stage operators.
// TEST INFRASTRUCTURE ONLY (oracle/_ref/libstereo_ref.so; built by `make -C oracle ref`).
//
// The operators (CVC, CVF + GuidedFilter_cv, DispSel) are /root/reference/src/{CVC,CVF,DispSel}.cpp
// compiled unmodified against oracle/shim/.  This file only plays the part of DispEst
// (/root/reference/src/DispEst.cpp), which cannot be compiled here (it drags in the OpenCL
// classes): buffer ownership as in the ctor (DispEst.cpp:31-49), CostConst_CPU's pthread batching
// copied in spirit from DispEst.cpp:222-270, the CostFilter_CPU that include/DispEst.h:42 declares
// but no source file defines -- written on the same model with CVF::preprocess +
// CVF::filterCV_thread -- and DispSelect_CPU (DispEst.cpp:311-321).
#include "CVC.h"
#include "CVF.h"
#include "DispSel.h"
#include "JointWMF.h"   // vendored CUHK weighted-median filter, compiled unmodified from /root/reference/include

#include <vector>

namespace {

double now_ms()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// DispEst.cpp:235-251: batches of `threads` pthreads, one per d, joined before the next batch
template <typename TD>
void run_batched(std::vector<TD>& td, int n, int threads, void* (*entry)(void*))
{
    if (threads < 1) threads = 1;
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setdetachstate(&attr, PTHREAD_CREATE_JOINABLE);
    std::vector<pthread_t> th(n);
    void* status;
    for (int level = 0; level <= n / threads; ++level) {
        const int block_size = (level < n / threads) ? threads : (n % threads);
        for (int iter = 0; iter < block_size; ++iter) {
            const int d = level * threads + iter;
            pthread_create(&th[d], &attr, entry, (void*)&td[d]);
        }
        for (int iter = 0; iter < block_size; ++iter) pthread_join(th[level * threads + iter], &status);
    }
    pthread_attr_destroy(&attr);
}

struct Vols {
    std::vector<Mat> l, r;
    Vols(int W, int H, int D) : l(D), r(D)
    {
        for (int d = 0; d < D; ++d) { l[d] = Mat::zeros(H, W, CV_32FC1); r[d] = Mat::zeros(H, W, CV_32FC1); }  // DispEst.cpp:31-37
    }
};

void copy_out(const Mat& m, float* dst, int W, int H)
{
    for (int y = 0; y < H; ++y) memcpy(dst + (size_t)y * W, m.ptr<float>(y), (size_t)W * sizeof(float));
}
void copy_in(Mat& m, const float* src, int W, int H)
{
    for (int y = 0; y < H; ++y) memcpy(m.ptr<float>(y), src + (size_t)y * W, (size_t)W * sizeof(float));
}

void cost_const(Mat& lImg, Mat& rImg, Mat& lGrdX, Mat& rGrdX, Vols& v, int D, int threads)
{
    CVC constructor;
    constructor.preprocess(lImg, lGrdX);   // DispEst.cpp:232-233
    constructor.preprocess(rImg, rGrdX);
    std::vector<buildCV_TD> td(D);
    for (int d = 0; d < D; ++d) td[d] = {&lImg, &rImg, &lGrdX, &rGrdX, d, &v.l[d]};
    run_batched(td, D, threads, CVC::buildCV_left_thread);
    for (int d = 0; d < D; ++d) td[d] = {&rImg, &lImg, &rGrdX, &lGrdX, d, &v.r[d]};   // swapped, DispEst.cpp:260
    run_batched(td, D, threads, CVC::buildCV_right_thread);
}

void cost_filter(Mat& lImg, Mat& rImg, Vols& v, int D, int threads)
{
    CVF filter;
    Mat* imgs[2] = {&lImg, &rImg};
    std::vector<Mat>* vols[2] = {&v.l, &v.r};
    for (int view = 0; view < 2; ++view) {
        Mat Img_rgb[3], mean_Img[3], var_Img[6];
        filter.preprocess(*imgs[view], Img_rgb, mean_Img, var_Img);   // CVF.cpp:44-70, once per view
        std::vector<filterCV_TD> td(D);
        for (int d = 0; d < D; ++d) td[d] = {Img_rgb, mean_Img, var_Img, &(*vols[view])[d]};
        run_batched(td, D, threads, CVF::filterCV_thread);             // CVF.cpp:28-41
    }
}

}  // namespace

extern "C" {

// Whole path.  Images: interleaved BGR float H*W*3.  Outputs (any may be NULL): gradients H*W, raw volumes
// D*H*W (before filtering), filtered volumes D*H*W, u8 maps H*W, times_ms[3] = cvc, cvf, wta (get_rt()-style).
int ref_pipeline(const float* l, const float* r, int W, int H, int D, int threads, int gray_mode,
                 float* lGrd, float* rGrd, float* lRaw, float* rRaw, float* lVol, float* rVol,
                 unsigned char* lDis, unsigned char* rDis, double* times_ms)
{
    psm_shim_set_gray_mode(gray_mode);
    Mat lImg(H, W, CV_32FC3, (void*)l), rImg(H, W, CV_32FC3, (void*)r);
    Mat lGrdX, rGrdX;
    Vols v(W, H, D);
    const double t0 = now_ms();
    cost_const(lImg, rImg, lGrdX, rGrdX, v, D, threads);
    const double t1 = now_ms();
    if (lGrd) copy_out(lGrdX, lGrd, W, H);
    if (rGrd) copy_out(rGrdX, rGrd, W, H);
    for (int d = 0; d < D; ++d) {
        if (lRaw) copy_out(v.l[d], lRaw + (size_t)d * W * H, W, H);
        if (rRaw) copy_out(v.r[d], rRaw + (size_t)d * W * H, W, H);
    }
    const double t2 = now_ms();
    cost_filter(lImg, rImg, v, D, threads);
    const double t3 = now_ms();
    for (int d = 0; d < D; ++d) {
        if (lVol) copy_out(v.l[d], lVol + (size_t)d * W * H, W, H);
        if (rVol) copy_out(v.r[d], rVol + (size_t)d * W * H, W, H);
    }
    Mat lDisMap = Mat::zeros(H, W, CV_8UC1), rDisMap = Mat::zeros(H, W, CV_8UC1);   // DispEst.cpp:46-47
    DispSel selector;
    const double t4 = now_ms();
    selector.CVSelect(v.l.data(), (unsigned)D, lDisMap);   // DispEst.cpp:314-318
    selector.CVSelect(v.r.data(), (unsigned)D, rDisMap);
    const double t5 = now_ms();
    for (int y = 0; y < H; ++y) {
        if (lDis) memcpy(lDis + (size_t)y * W, lDisMap.ptr<unsigned char>(y), W);
        if (rDis) memcpy(rDis + (size_t)y * W, rDisMap.ptr<unsigned char>(y), W);
    }
    if (times_ms) { times_ms[0] = t1 - t0; times_ms[1] = t3 - t2; times_ms[2] = t5 - t4; }
    return 0;
}

// GuidedFilter_cv (CVF.cpp:72-165) on one caller-provided slice p (H*W), guide = interleaved BGR image.
int ref_guided_filter(const float* img, int W, int H, const float* p, float* q)
{
    Mat Img(H, W, CV_32FC3, (void*)img);
    CVF filter;
    Mat Img_rgb[3], mean_Img[3], var_Img[6];
    filter.preprocess(Img, Img_rgb, mean_Img, var_Img);
    Mat cost = Mat::zeros(H, W, CV_32FC1);
    copy_in(cost, p, W, H);
    filter.filterCV(Img_rgb, mean_Img, var_Img, cost);   // CVF.cpp:22-26
    copy_out(cost, q, W, H);
    return 0;
}

// DispSel::CVSelect and the pthread-per-row twin CVSelect_thread (DispSel.cpp:83-109 / :53-81) on a caller volume.
int ref_wta(const float* vol, int W, int H, int D, int use_thread_variant, int threads, unsigned char* dis)
{
    std::vector<Mat> cv(D);
    for (int d = 0; d < D; ++d) cv[d] = Mat(H, W, CV_32FC1, (void*)(vol + (size_t)d * W * H));
    Mat map = Mat::zeros(H, W, CV_8UC1);
    DispSel selector;
    if (use_thread_variant) selector.CVSelect_thread(cv.data(), (unsigned)D, map, threads);
    else selector.CVSelect(cv.data(), (unsigned)D, map);
    for (int y = 0; y < H; ++y) memcpy(dis + (size_t)y * W, map.ptr<unsigned char>(y), W);
    return 0;
}

// The non-pthread twins CVC::buildCV_left / buildCV_right (CVC.cpp:122-179) for one slice (parity of both code paths).
int ref_buildcv(const float* l, const float* r, int W, int H, int d, int right, int gray_mode, float* cost)
{
    psm_shim_set_gray_mode(gray_mode);
    Mat lImg(H, W, CV_32FC3, (void*)l), rImg(H, W, CV_32FC3, (void*)r);
    Mat lGrdX, rGrdX;
    CVC constructor;
    constructor.preprocess(lImg, lGrdX);
    constructor.preprocess(rImg, rGrdX);
    Mat c = Mat::zeros(H, W, CV_32FC1);
    if (right) constructor.buildCV_right(rImg, lImg, rGrdX, lGrdX, d, c);   // DispEst.cpp:217
    else constructor.buildCV_left(lImg, rImg, lGrdX, rGrdX, d, c);
    copy_out(c, cost, W, H);
    return 0;
}

// One view of PP::processDM's live code (src/PP.cpp:414-422): img.convertTo(CV_8UC3, 255) then
// JointWMF::filter(disp, img8UC3, MED_SZ/2 = 9) with its defaults (sigma 25.5, nI = nF = 256, "exp" weights).
// The feature clustering inside (JointWMF.h:586-591) calls cv::kmeans, which the shim can only stand in for
// (see shim/opencv2/opencv.hpp): the result is the reference's only when the image has <= 256 distinct 6-bit colours.
int ref_post_process(const float* img, const unsigned char* disp, int W, int H, int r, unsigned char* out, int* n_colours)
{
    Mat Img(H, W, CV_32FC3, (void*)img);
    Mat Img8;
    Img.convertTo(Img8, CV_8UC3, 255);
    if (n_colours) {   // distinct 6-bit colours (what JointWMF.h:552-568 counts)
        std::vector<unsigned char> seen(64 * 64 * 64, 0);
        int cnt = 0;
        for (int y = 0; y < H; ++y) {
            const unsigned char* p = Img8.ptr<unsigned char>(y);
            for (int x = 0; x < W; ++x) {
                const int k = ((p[3 * x] >> 2) * 64 + (p[3 * x + 1] >> 2)) * 64 + (p[3 * x + 2] >> 2);
                if (!seen[k]) { seen[k] = 1; ++cnt; }
            }
        }
        *n_colours = cnt;
    }
    Mat D = Mat::zeros(H, W, CV_8UC1);
    for (int y = 0; y < H; ++y) memcpy(D.ptr<unsigned char>(y), disp + (size_t)y * W, W);
    Mat res = JointWMF::filter(D, Img8, r);
    for (int y = 0; y < H; ++y) memcpy(out + (size_t)y * W, res.ptr<unsigned char>(y), W);
    return 0;
}

const char* ref_build_info(void)
{
    return "oracle/_ref: /root/reference/src/{CVC,CVF,DispSel}.cpp compiled unmodified against oracle/shim (g++ -O3 -std=c++11 -fopenmp -pthread)";
}

}  // extern "C"
