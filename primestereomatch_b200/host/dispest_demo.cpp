// dispest_demo.cpp -- drives the C++ DispEst facade exactly like StereoMatch::compute does
// (reference src/StereoMatch.cpp:198-241): setInputImages -> CostConst_GPU -> CostFilter_GPU ->
// DispSelect_GPU -> PostProcess_GPU.  Raw I/O so it needs no image library:
//   dispest_demo W H D left.f32 right.f32 out_left.u8 out_right.u8
#include <cstdio>
#include <vector>

#include "DispEstB200.h"

static bool read_all(const char* path, void* dst, size_t n)
{
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    const size_t got = std::fread(dst, 1, n, f);
    std::fclose(f);
    return got == n;
}

int main(int argc, char** argv)
{
    if (argc != 8) { std::fprintf(stderr, "usage: %s W H D left.f32 right.f32 outL.u8 outR.u8\n", argv[0]); return 2; }
    const int W = std::atoi(argv[1]), H = std::atoi(argv[2]), D = std::atoi(argv[3]);
    if (b200devicepoll() == 0) { std::fprintf(stderr, "no CUDA device\n"); return 3; }
    cv::Mat l(H, W, CV_32FC3), r(H, W, CV_32FC3);
    if (!read_all(argv[4], l.data, l.step * H) || !read_all(argv[5], r.data, r.step * H)) { std::fprintf(stderr, "bad input\n"); return 2; }
    DispEst* SMDE = new DispEst(l, r, D, MAX_CPU_THREADS, true);
    SMDE->setInputImages(l, r);
    SMDE->setThreads(MAX_CPU_THREADS);
    SMDE->setSubsampleRate(4);
    int rc = SMDE->CostConst_GPU();
    if (!rc) rc = SMDE->CostFilter_GPU();
    if (!rc) rc = SMDE->DispSelect_GPU();
    if (!rc) rc = SMDE->PostProcess_GPU();
    if (rc) { std::fprintf(stderr, "stage failed (%d): %s\n", rc, SMDE->last_error()); return 1; }
    std::printf("CVC %.3f ms  CVF %.3f ms  DispSel %.3f ms  PP %.3f ms\n", SMDE->stage_ms(1), SMDE->stage_ms(2), SMDE->stage_ms(3), SMDE->stage_ms(5));
    FILE* f = std::fopen(argv[6], "wb"); std::fwrite(SMDE->lDisMap.data, 1, (size_t)W * H, f); std::fclose(f);
    f = std::fopen(argv[7], "wb"); std::fwrite(SMDE->rDisMap.data, 1, (size_t)W * H, f); std::fclose(f);
    delete SMDE;
    return 0;
}
