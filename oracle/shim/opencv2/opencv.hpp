/*
 * opencv2/opencv.hpp -- MINIMAL STAND-IN for the OpenCV C++ API.  TEST INFRASTRUCTURE ONLY.
 *
 * Purpose: compile the reference's own src/CVC.cpp, src/CVF.cpp and src/DispSel.cpp UNMODIFIED,
 * from where they lie under /root/reference, into oracle/_ref/libstereo_ref.so (recipe:
 * oracle/Makefile, target `ref`).  The image has no OpenCV C++ headers, so
 * this header supplies exactly the slice of the API those three files touch:
 *
 *   cv::Mat   rows, cols, ptr<T>(y), at<T>(y,x), zeros(), shared ownership on copy/assign,
 *             Mat - Mat, -=, +=                                   (CVF.cpp:94,154,162; DispSel.cpp:105)
 *   cv::Size, cv::split, cv::boxFilter, cv::multiply, cv::cvtColor, cv::Sobel
 *             (CVC.cpp:43-44; CVF.cpp:47,50,62-64,82,87-88,93,153,158,160-161)
 *   CV_32F, CV_32FC1, CV_32FC3, CV_8UC1, CV_RGB2GRAY
 *
 * The five image primitives are implemented in opencv_shim.cpp by calling the C primitives of
 * oracle/stereo_oracle.c, whose arithmetic is pinned bit-for-bit against python cv2 4.13.0
 * (tests/test_oracle.py, tests/golden/make_golden.py).  Everything else the reference does --
 * loop structure, operation order, float/double promotion, thread entry points -- is the
 * reference's own compiled code.
 */
#ifndef PSM_SHIM_OPENCV_HPP
#define PSM_SHIM_OPENCV_HPP

#include <algorithm>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

namespace cv {

typedef unsigned char uchar;

#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 63) + 1)

enum { CV_RGB2GRAY = 7, COLOR_RGB2GRAY = 7 };
enum { CV_TERMCRIT_ITER = 1, CV_TERMCRIT_EPS = 2, KMEANS_PP_CENTERS = 2 };

struct Scalar {
    double val[4];
    Scalar(double v0 = 0, double v1 = 0, double v2 = 0, double v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
};
struct TermCriteria {
    int type, maxCount;
    double epsilon;
    TermCriteria(int t = 0, int n = 0, double e = 0) : type(t), maxCount(n), epsilon(e) {}
};

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};

class Mat {
public:
    int rows, cols;
    size_t step;      // bytes per row
    uchar* data;

    Mat() : rows(0), cols(0), step(0), data(nullptr), type_(0) {}
    Mat(int r, int c, int type) : rows(0), cols(0), step(0), data(nullptr), type_(0) { create(r, c, type); }
    Mat(Size sz, int type) : rows(0), cols(0), step(0), data(nullptr), type_(0) { create(sz.height, sz.width, type); }
    // header over caller-owned memory (no ownership), like cv::Mat(rows, cols, type, void*, step)
    Mat(int r, int c, int type, void* ext, size_t ext_step = 0)
        : rows(r), cols(c), step(ext_step ? ext_step : (size_t)c * elem(type)), data(static_cast<uchar*>(ext)), type_(type) {}

    void create(int r, int c, int type)
    {
        if (data && r == rows && c == cols && type == type_) return;  // cv::Mat::create keeps a fitting buffer
        rows = r; cols = c; type_ = type;
        step = (size_t)c * elem(type);
        buf_ = std::shared_ptr<uchar>(new uchar[step * (size_t)(r > 0 ? r : 1)], std::default_delete<uchar[]>());
        data = buf_.get();
    }
    static Mat zeros(int r, int c, int type)
    {
        Mat m(r, c, type);
        std::memset(m.data, 0, m.step * (size_t)r);
        return m;
    }
    Mat clone() const
    {
        Mat m(rows, cols, type_);
        for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * elem(type_));
        return m;
    }
    int type() const { return type_; }
    int depth() const { return CV_MAT_DEPTH(type_); }
    int channels() const { return CV_MAT_CN(type_); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    Mat& operator=(const Scalar& s);                                   // fill (1-channel)
    void convertTo(Mat& dst, int rtype, double alpha = 1.0, double beta = 0.0) const;   // 8U <-> 32S, 32F -> 8U (rounded, saturated)
    template <typename T> T* ptr(int y, int x) { return reinterpret_cast<T*>(data + (size_t)y * step) + x; }
    template <typename T> const T* ptr(int y, int x) const { return reinterpret_cast<const T*>(data + (size_t)y * step) + x; }
    size_t elemSize() const { return elem(type_); }

    template <typename T> T* ptr(int y = 0) { return reinterpret_cast<T*>(data + (size_t)y * step); }
    template <typename T> const T* ptr(int y = 0) const { return reinterpret_cast<const T*>(data + (size_t)y * step); }
    template <typename T> T& at(int y, int x) { return reinterpret_cast<T*>(data + (size_t)y * step)[x]; }
    template <typename T> const T& at(int y, int x) const { return reinterpret_cast<const T*>(data + (size_t)y * step)[x]; }

    Mat& operator-=(const Mat& o);   // element-wise IEEE float ops (cv::subtract / cv::add on CV_32F)
    Mat& operator+=(const Mat& o);

private:
    static size_t elem(int type) { return (size_t)(CV_MAT_DEPTH(type) == CV_8U ? 1 : 4) * CV_MAT_CN(type); }
    int type_;
    std::shared_ptr<uchar> buf_;
};

Mat operator-(const Mat& a, const Mat& b);

void split(const Mat& src, Mat* mv);
void split(const Mat& src, std::vector<Mat>& mv);     // 1-channel sources only (JointWMF.h:100 on the disparity map)
void merge(const std::vector<Mat>& mv, Mat& dst);
void minMaxLoc(const Mat& src, double* minVal, double* maxVal = nullptr);   // declared for JointWMF.h:695 (float-image path, never called here)     // 1-channel only (JointWMF.h:151)
/* cv::kmeans is un-vendored, RNG-seeded OpenCV code (k-means++ on cv::theRNG()): it CANNOT be pinned.  The stand-in is
 * exact when K equals the number of samples (every sample its own cluster, labels = identity) -- the only case the parity
 * tests use (JointWMF.h:570 clamps nF to the number of distinct 6-bit colours) -- and otherwise a deterministic
 * farthest-point seeding followed by two Lloyd iterations, documented as NOT the reference's clustering. */
double kmeans(const Mat& samples, int K, Mat& labels, TermCriteria crit, int attempts, int flags, Mat& centers);
void boxFilter(const Mat& src, Mat& dst, int ddepth, Size ksize);
void multiply(const Mat& a, const Mat& b, Mat& dst, double scale = 1.0, int dtype = -1);
void cvtColor(const Mat& src, Mat& dst, int code, int dstCn = 0);
void Sobel(const Mat& src, Mat& dst, int ddepth, int dx, int dy, int ksize = 3);

}  // namespace cv

/* selects the RGB2GRAY rounding of the OpenCV build being imitated (see oracle/stereo_oracle.h:
 * 0 = FMA form of cv2 4.13 SIMD/IPP builds, 1 = plain left-to-right form) */
extern "C" void psm_shim_set_gray_mode(int mode);

#endif
