// DispEstB200.h -- C++ host facade with the reference's DispEst API surface
// (/root/reference/include/DispEst.h:21-109) whose "_GPU" stage methods run on the B200 C-ABI
// library (include/prime_stereo_b200.h) instead of the reference's OpenCL classes.
//
// Two ways to use it:
//   1. in the reference application: keep the reference's own DispEst.h/.cpp for the CPU stages and
//      patch the three _GPU bodies + ctor/dtor with the calls shown in INTEGRATION.md;
//   2. standalone (this header): compiles against OpenCV when its headers are present, otherwise
//      against the minimal cv::Mat stand-in below (this image has no OpenCV C++ headers).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "../../include/prime_stereo_b200.h"

#if !defined(PSM_NO_OPENCV) && defined(__has_include)
#if __has_include(<opencv2/core.hpp>)
#include <opencv2/core.hpp>
#define PSM_HAVE_OPENCV 1
#endif
#endif

#ifndef PSM_HAVE_OPENCV
// Minimal stand-in for the handful of cv::Mat members the facade touches.
#define CV_8U 0
#define CV_32F 5
#define CV_MAT_DEPTH_MASK 7
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
namespace cv {
class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    unsigned char* data = nullptr;
    Mat() = default;
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* ext, size_t ext_step) : rows(r), cols(c), step(ext_step), data((unsigned char*)ext), type_(type) {}
    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); std::memset(m.data, 0, m.step * r); return m; }
    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type;
        step = (size_t)c * elemSize();
        owner_.reset(new unsigned char[step * r], std::default_delete<unsigned char[]>());
        data = owner_.get();
    }
    int type() const { return type_; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize() const { return (size_t)channels() * ((type_ & CV_MAT_DEPTH_MASK) == CV_32F ? 4 : 1); }
    bool empty() const { return data == nullptr; }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + step * y); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + step * y); }
private:
    int type_ = 0;
    std::shared_ptr<unsigned char> owner_;
};
}  // namespace cv
#endif

#define MAX_CPU_THREADS 8  // reference include/ComFunc.h:52
#define OCV_DE 0
#define OCL_DE 1

// Role of openCLdevicepoll() (reference src/main.cpp:29): number of compute devices; 0 disables the `m` toggle.
inline int b200devicepoll() { return psm_device_count(); }

class DispEst {
public:
    // reference src/DispEst.cpp:10-143.  `ocl` == use the GPU path.
    DispEst(cv::Mat l, cv::Mat r, const int d, int t, bool ocl);
    ~DispEst(void);

    cv::Mat lDisMap;  // CV_8UC1, raw disparity index (reference DispEst.h:28-29)
    cv::Mat rDisMap;

    int setInputImages(cv::Mat l, cv::Mat r);
    int setThreads(unsigned int newThreads);
    void setSubsampleRate(unsigned int newRate) { subsample_rate = newRate; }
    int printCV(void);

    // GPU stages (the path behind the `m` toggle): 0 = success, as in the reference
    int CostConst_GPU();
    int CostFilter_GPU();
    int DispSelect_GPU();
    int PostProcess_GPU();

    // CPU stages are the reference's own code and are not part of this library
    int CostConst() { return not_built("CostConst"); }
    int CostConst_CPU() { return not_built("CostConst_CPU"); }
    int CostFilter() { return not_built("CostFilter"); }
    int CostFilter_CPU() { return not_built("CostFilter_CPU"); }
    int CostFilter_FGF() { return not_built("CostFilter_FGF"); }
    int DispSelect_CPU() { return not_built("DispSelect_CPU"); }
    int PostProcess_CPU() { return not_built("PostProcess_CPU"); }

    // stage timers of the last frame in ms (role of cvc_time/cvf_time/dispsel_time, StereoMatch.cpp:209-241)
    float stage_ms(int stage) const { float ms = -1.f; psm_stage_ms(ctx, stage, &ms); return ms; }
    const char* last_error() const { return psm_last_error(ctx); }

private:
    int not_built(const char* what) const
    {
        std::fprintf(stderr, "DispEst::%s: CPU stage not built into the B200 facade (use the reference's own)\n", what);
        return -1;
    }
    cv::Mat lImg, rImg;
    int hei, wid, maxDis, threads;
    bool useOCL;
    unsigned int subsample_rate = 4;
    psm_ctx* ctx = nullptr;
};
