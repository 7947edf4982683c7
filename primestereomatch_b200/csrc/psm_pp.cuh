// psm_pp.cuh -- post-processing: joint weighted-median filter of the disparity maps on the device.
//
// Reference: PP::processDM (/root/reference/src/PP.cpp:402-425), whose live code is
//   img.convertTo(img8UC3, CV_8UC3, 255);  disp = JointWMF::filter(disp, img8UC3, MED_SZ/2 = 9)
// (vendored include/JointWMF.h:81-155, filterCore :173-390) and which the reference runs on the CPU even in
// its GPU mode (src/DispEst.cpp:338-344).
//
// What is computed (exactly what the CPU restatement orc_wmf (test infrastructure, oracle/) restates, and what the reference's
// JointWMF computes whenever the feature image has <= 256 distinct 6-bit colours): for every pixel p
//   out(p) = min { v : sum_{q in window, I_q <= v} w(p,q)  >=  sum_{q in window, I_q > v} w(p,q) }
// window = (2r+1)^2 clipped at the image border, w = exp(-|c_p - c_q|^2 / (2 s^2)) on the 6-bit colours
// c = (B,G,R) >> 2, s = 25.5/256*64, held as 2^-22 fixed point (table indexed by the integer squared colour
// distance, computed on the host with the reference's float expression).  The reference additionally clusters
// the colours to 256 indices with cv::kmeans (RNG-seeded, un-vendored): parity for natural images is unpinned
// by the reference itself; this kernel needs no clustering -- it evaluates the weights directly.
//
// B200 mapping (nothing like the reference's serial column scan with necklace tables): one WARP per pixel,
//   the 361 window taps dealt round-robin to the 32 lanes (12 rounds), each tap one 4-byte load of a packed
//   u32 image (disparity << 24 | R6 << 16 | G6 << 8 | B6);
//   per round the lanes holding the same disparity combine their weights (__match_any_sync + __reduce_add_sync) and one
//   lane adds the sum to a 256-bin histogram in shared memory (1 KB per warp): integer sums, order-independent;
//   then each lane scans 8 bins, a warp prefix sum finds the first bin where 2*cum >= total.
#pragma once
#include "psm_common.cuh"

namespace psm {

constexpr int kPpRadius = 9;                 // MED_SZ / 2, reference include/PP.h:12
constexpr int kPpMaxD2 = 3 * 63 * 63;        // largest squared distance between two 6-bit colours
constexpr int kPpWarps = 8;                  // warps (pixels in flight) per CTA

// guide planes (float, BGR planar) + u8 disparity map -> packed u32 image
__global__ void pp_pack_kernel(const float* __restrict__ I0, const float* __restrict__ I1, const float* __restrict__ I2, int Wp,
                               const uint8_t* __restrict__ disp, int W, int H, uint32_t* __restrict__ packed)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W || y >= H) return;
    const size_t o = (size_t)y * Wp + x;
    // convertTo(CV_8UC3, 255): saturate_cast<uchar>(cvRound(v * 255.0f)), round half to even (PP.cpp:416-417)
    auto q6 = [](float v) {
        int q = __float2int_rn(fmul(v, 255.0f));
        q = q < 0 ? 0 : (q > 255 ? 255 : q);
        return (uint32_t)(q >> 2);           // JointWMF.h:546,559-561: 8 bit -> 6 bit
    };
    const uint32_t b = q6(I0[o]), g = q6(I1[o]), r = q6(I2[o]);
    packed[(size_t)y * W + x] = ((uint32_t)disp[(size_t)y * W + x] << 24) | (r << 16) | (g << 8) | b;
}

__global__ void __launch_bounds__(kPpWarps * 32) pp_wmf_kernel(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ lut,
                                                              int W, int H, uint8_t* __restrict__ out)
{
    __shared__ uint32_t hist[kPpWarps][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* h = hist[warp];
#pragma unroll
    for (int k = 0; k < 8; ++k) h[lane * 8 + k] = 0;
    __syncwarp();
    const long npix = (long)W * H;
    for (long pix = (long)blockIdx.x * kPpWarps + warp; pix < npix; pix += (long)gridDim.x * kPpWarps) {
        const int y = (int)(pix / W), x = (int)(pix - (long)y * W);
        const uint32_t cp = __ldg(packed + pix);
        const int pb = cp & 63, pg = (cp >> 8) & 63, pr = (cp >> 16) & 63;
        uint32_t total = 0;
        constexpr int kSide = 2 * kPpRadius + 1, kTaps = kSide * kSide;
        // the 361 window taps are dealt to the 32 lanes round-robin (12 rounds); a lane's tap is (i / 19, i % 19)
        for (int base = 0; base < kTaps; base += 32) {
            const int i = base + lane;
            const int dy = i / kSide, dx = i - dy * kSide;
            const int yy = y - kPpRadius + dy, xx = x - kPpRadius + dx;
            const bool ok = i < kTaps && yy >= 0 && yy < H && xx >= 0 && xx < W;
            uint32_t w = 0;
            unsigned dq = 0x100u + lane;      // a key no valid tap has
            if (ok) {
                const uint32_t cq = __ldg(packed + (size_t)yy * W + xx);
                const int d0 = pb - (int)(cq & 63), d1 = pg - (int)((cq >> 8) & 63), d2 = pr - (int)((cq >> 16) & 63);
                w = __ldg(lut + (d0 * d0 + d1 * d1 + d2 * d2));
                dq = cq >> 24;
            }
            total += w;
            // combine the taps that carry the same disparity (neighbouring pixels usually do) before touching shared memory:
            // every lane learns its peer set in one MATCH, the peers add their weights with one REDUX, the lowest peer owns the bin
            const unsigned peers = __match_any_sync(0xffffffffu, dq);
            const uint32_t s = __reduce_add_sync(peers, w);
            if (ok && lane == __ffs(peers) - 1) h[dq] += s;
        }
        __syncwarp();
        total = __reduce_add_sync(0xffffffffu, total);
        // lane l owns bins 8l .. 8l+7
        uint32_t b8[8], mine = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { b8[k] = h[lane * 8 + k]; mine += b8[k]; h[lane * 8 + k] = 0; }
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        // first lane whose inclusive sum reaches half of the total (2*cum >= total; sums < 2^31)
        const unsigned reach = __ballot_sync(0xffffffffu, 2u * incl >= total);
        const int first = __ffs(reach) - 1;   // total > 0 (the centre tap has weight 1), so lane 31 always reaches
        if (lane == first) {
            uint32_t cum = incl - mine;
            int v = lane * 8;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                cum += b8[k];
                if (2u * cum >= total) { v = lane * 8 + k; break; }
            }
            out[pix] = (uint8_t)(v > 255 ? 255 : v);
        }
        __syncwarp();
    }
}

}  // namespace psm
