// opencv_shim.cpp -- the five image primitives behind oracle/shim/opencv2/opencv.hpp.
// TEST INFRASTRUCTURE ONLY (part of oracle/_ref).  Each primitive forwards to the C restatement in
// oracle/stereo_oracle.c, which is pinned against python cv2 (tests/test_oracle.py).
#include <opencv2/opencv.hpp>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../stereo_oracle.h"

namespace {
int g_gray_mode = 0;

void need(bool ok, const char* what)
{
    if (!ok) { std::fprintf(stderr, "opencv shim: unsupported use: %s\n", what); std::abort(); }
}

// contiguous copy of a 1-channel float Mat (the C primitives take packed planes)
std::vector<float> packed(const cv::Mat& m)
{
    const int cn = m.channels();
    std::vector<float> v((size_t)m.rows * m.cols * cn);
    for (int y = 0; y < m.rows; ++y) std::memcpy(&v[(size_t)y * m.cols * cn], m.ptr<float>(y), (size_t)m.cols * cn * sizeof(float));
    return v;
}
void unpack(const std::vector<float>& v, cv::Mat& dst, int rows, int cols)
{
    dst.create(rows, cols, CV_32FC1);
    for (int y = 0; y < rows; ++y) std::memcpy(dst.ptr<float>(y), &v[(size_t)y * cols], (size_t)cols * sizeof(float));
}
}  // namespace

extern "C" void psm_shim_set_gray_mode(int mode) { g_gray_mode = mode; }

namespace cv {

Mat& Mat::operator-=(const Mat& o)
{
    need(type() == CV_32FC1 && o.type() == CV_32FC1 && rows == o.rows && cols == o.cols, "Mat -= Mat (CV_32FC1, same size)");
    for (int y = 0; y < rows; ++y) {
        float* a = ptr<float>(y);
        const float* b = o.ptr<float>(y);
        for (int x = 0; x < cols; ++x) a[x] = a[x] - b[x];
    }
    return *this;
}

Mat& Mat::operator+=(const Mat& o)
{
    need(type() == CV_32FC1 && o.type() == CV_32FC1 && rows == o.rows && cols == o.cols, "Mat += Mat (CV_32FC1, same size)");
    for (int y = 0; y < rows; ++y) {
        float* a = ptr<float>(y);
        const float* b = o.ptr<float>(y);
        for (int x = 0; x < cols; ++x) a[x] = a[x] + b[x];
    }
    return *this;
}

Mat operator-(const Mat& a, const Mat& b)
{
    need(a.type() == CV_32FC1 && b.type() == CV_32FC1 && a.rows == b.rows && a.cols == b.cols, "Mat - Mat (CV_32FC1, same size)");
    Mat r(a.rows, a.cols, CV_32FC1);
    for (int y = 0; y < a.rows; ++y) {
        const float* pa = a.ptr<float>(y);
        const float* pb = b.ptr<float>(y);
        float* pr = r.ptr<float>(y);
        for (int x = 0; x < a.cols; ++x) pr[x] = pa[x] - pb[x];
    }
    return r;
}

void split(const Mat& src, Mat* mv)
{
    need(src.type() == CV_32FC3, "split(CV_32FC3)");
    for (int c = 0; c < 3; ++c) mv[c].create(src.rows, src.cols, CV_32FC1);
    for (int y = 0; y < src.rows; ++y) {
        const float* s = src.ptr<float>(y);
        for (int c = 0; c < 3; ++c) {
            float* d = mv[c].ptr<float>(y);
            for (int x = 0; x < src.cols; ++x) d[x] = s[3 * x + c];
        }
    }
}

void boxFilter(const Mat& src, Mat& dst, int ddepth, Size ksize)
{
    need(src.type() == CV_32FC1 && ddepth == -1 && ksize.width == ORC_GIF_R_WIN && ksize.height == ORC_GIF_R_WIN,
         "boxFilter(CV_32FC1, -1, Size(8,8))");
    const int rows = src.rows, cols = src.cols;
    if (src.step == (size_t)cols * sizeof(float)) {   // packed rows: filter straight into the destination (orc_box8 allows src == dst)
        const Mat keep = src;                          // dst may be the same header as src
        dst.create(rows, cols, CV_32FC1);
        orc_box8(keep.ptr<float>(0), cols, rows, dst.ptr<float>(0));
        return;
    }
    std::vector<float> in = packed(src), out(in.size());
    orc_box8(in.data(), cols, rows, out.data());
    unpack(out, dst, rows, cols);
}

void multiply(const Mat& a, const Mat& b, Mat& dst, double scale, int dtype)
{
    need(a.type() == CV_32FC1 && b.type() == CV_32FC1 && a.rows == b.rows && a.cols == b.cols && scale == 1.0 && dtype == -1,
         "multiply(CV_32FC1, CV_32FC1)");
    const int rows = a.rows, cols = a.cols;
    Mat out = (dst.data == a.data || dst.data == b.data) ? dst : Mat();   // in-place use is legal (CVF.cpp:161)
    out.create(rows, cols, CV_32FC1);
    for (int y = 0; y < rows; ++y) {
        const float* pa = a.ptr<float>(y);
        const float* pb = b.ptr<float>(y);
        float* pd = out.ptr<float>(y);
        for (int x = 0; x < cols; ++x) pd[x] = pa[x] * pb[x];
    }
    dst = out;
}

void cvtColor(const Mat& src, Mat& dst, int code, int)
{
    need(src.type() == CV_32FC3 && code == CV_RGB2GRAY, "cvtColor(CV_32FC3, CV_RGB2GRAY)");
    std::vector<float> in = packed(src), out((size_t)src.rows * src.cols);
    orc_rgb2gray(in.data(), src.cols, src.rows, out.data(), g_gray_mode);
    unpack(out, dst, src.rows, src.cols);
}

void Sobel(const Mat& src, Mat& dst, int ddepth, int dx, int dy, int ksize)
{
    need(src.type() == CV_32FC1 && ddepth == CV_32F && dx == 1 && dy == 0 && ksize == 1, "Sobel(CV_32F, 1, 0, 1)");
    std::vector<float> in = packed(src), out(in.size());
    orc_sobel_x(in.data(), src.cols, src.rows, out.data());
    unpack(out, dst, src.rows, src.cols);
}

}  // namespace cv
