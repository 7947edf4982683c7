// DispEstB200.cpp -- see DispEstB200.h.  Every method cites the reference code it stands in for.
#include "DispEstB200.h"

#include <cassert>

// reference src/DispEst.cpp:10-143: type check (exit(1) on mismatch), output maps, and -- instead of
// the OpenCL context / queue / 12 cl_mem buffers / three run-time compiled programs -- one psm_ctx.
DispEst::DispEst(cv::Mat l, cv::Mat r, const int d, int t, bool ocl)
    : lImg(l), rImg(r), maxDis(d), threads(t), useOCL(ocl)
{
    hei = lImg.rows;
    wid = lImg.cols;
    if (lImg.type() != rImg.type()) {
        std::printf("DE: Error - Left & Right images are of different types.\n");
        std::exit(1);
    }
    lDisMap = cv::Mat::zeros(hei, wid, CV_8UC1);
    rDisMap = cv::Mat::zeros(hei, wid, CV_8UC1);
    if (useOCL) {
        if (psm_create(&ctx, wid, hei, maxDis, 0) != PSM_OK) {
            // reference logs OpenCL set-up failures and carries on (DispEst.cpp:69-79); so do we, with
            // the GPU stages then returning an error code instead of computing anything on the CPU
            std::fprintf(stderr, "Failed to create the B200 context: %s\n", psm_last_error(nullptr));
            ctx = nullptr;
        }
    }
}

// reference src/DispEst.cpp:145-162
DispEst::~DispEst(void)
{
    if (ctx) psm_destroy(ctx);
}

// reference src/DispEst.cpp:164-170
int DispEst::setInputImages(cv::Mat leftImg, cv::Mat rightImg)
{
    assert(leftImg.type() == rightImg.type());
    lImg = leftImg;
    rImg = rightImg;
    return 0;
}

// reference src/DispEst.cpp:172-179
int DispEst::setThreads(unsigned int newThreads)
{
    if (newThreads > MAX_CPU_THREADS) return -1;
    threads = newThreads;
    return 0;
}

// reference src/DispEst.cpp:181-194 dumps every slice as PNG; here: slice statistics on stdout
int DispEst::printCV(void)
{
    if (!ctx) return -1;
    cv::Mat s(hei, wid, CV_32FC1);
    for (int v = 0; v < 2; ++v)
        for (int i = 0; i < maxDis; ++i) {
            if (int rc = psm_read_cost_slice(ctx, v, i, s.ptr<float>(), s.step)) return rc;
            double sum = 0;
            for (int y = 0; y < hei; ++y)
                for (int x = 0; x < wid; ++x) sum += s.ptr<float>(y)[x];
            std::printf("%cCV%d mean %.6f\n", v ? 'r' : 'l', i, sum / ((double)hei * wid));
        }
    return 0;
}

// reference src/DispEst.cpp:272-276 -> CVC_cl::buildCV (src/CVC_cl.cpp:93-211): upload + build both volumes
int DispEst::CostConst_GPU()
{
    if (!ctx) return -1;
    // the context was sized at construction: a differently shaped or typed Mat from setInputImages must not reach
    // the 2-D copy (the reference re-creates DispEst on every dataset change, StereoMatch.cpp:601-602)
    if (lImg.rows != hei || lImg.cols != wid || rImg.rows != hei || rImg.cols != wid || lImg.type() != rImg.type() ||
        lImg.channels() != 3) {
        std::fprintf(stderr, "DispEst::CostConst_GPU: input images must be %dx%d, 3 channels, equal types\n", wid, hei);
        return -1;
    }
    int rc;
    if ((lImg.type() & CV_MAT_DEPTH_MASK) == CV_32F)
        rc = psm_set_images(ctx, lImg.ptr<float>(), lImg.step, rImg.ptr<float>(), rImg.step);
    else
        rc = psm_set_images_u8(ctx, lImg.ptr<uint8_t>(), lImg.step, rImg.ptr<uint8_t>(), rImg.step);
    if (rc) return rc;
    return psm_cost_const(ctx);
}

// reference src/DispEst.cpp:299-308 -> CVF_cl::preprocess + CVF_cl::filterCV per view
int DispEst::CostFilter_GPU()
{
    if (!ctx) return -1;
    return psm_cost_filter(ctx);
}

// reference src/DispEst.cpp:323-328 -> DispSel_cl::CVSelect (src/DispSel_cl.cpp:69-139)
int DispEst::DispSelect_GPU()
{
    if (!ctx) return -1;
    return psm_disp_select(ctx, lDisMap.ptr<uint8_t>(), lDisMap.step, rDisMap.ptr<uint8_t>(), rDisMap.step);
}

// reference src/DispEst.cpp:338-344 -> PP::processDM (src/PP.cpp:402-425): the joint weighted-median filter of
// both maps.  The reference runs it on the CPU even in GPU mode; here it is a device stage (psm_pp.cuh).
int DispEst::PostProcess_GPU()
{
    if (!ctx) return -1;
    return psm_post_process(ctx, lDisMap.ptr<uint8_t>(), lDisMap.step, rDisMap.ptr<uint8_t>(), rDisMap.step);
}
