// psm_fgf.cuh -- the Fast Guided Filter branch (CostFilter_FGF) on the device.
//
// Reference: FastGuidedFilterColor (/root/reference/src/fastguidedfilter.cpp:121-198) as driven by
// DispEst::CostFilter_FGF (src/DispEst.cpp:281-296): FastGuidedFilter(img, GIF_R_WIN = 8, GIF_EPS, s); every
// slice is NN-subsampled by s, guided-filtered there with a K x K box, K = 2*(8/s)+1 (:206-208), and the four
// coefficient means are bilinearly up-sampled and applied at full resolution.  It is what the reference's CPU
// branch runs today (src/StereoMatch.cpp:214); it is NOT the full-resolution GIF of psm_cvf_stream.cuh.
//
// Numerics follow the CPU restatement orc_fgf_* (test infrastructure, oracle/) (pinned against cv2 with IPP off, tests/golden/
// make_golden_fgf.py): box sums in fp64 (exact: <= 81 terms), * 1/(K*K) in double, one rounding; every float
// operation a separate IEEE op in source order; the diagonal variance terms (x - y + eps) in double with one
// rounding (cv::addWeighted); true divisions; OpenCV's own INTER_NN / INTER_LINEAR index and weight formulas
// (tables built on the host by the same float / double expressions, see psm_capi.cu).
//
// Work split (nothing like the reference's per-slice Mat pipeline of ~40 OpenCV calls):
//   fgf_guide_kernel : per view, per low-res pixel: channel means, (Sigma + eps I)^-1          (once per frame)
//   fgf_lowres_kernel: per slice tile, both low-res stages in shared memory (separable fp64 box sums): p and I*p box
//                      means -> a (3), b -> their box means                                      (reads 1/s^2 of p)
//   fgf_up_kernel    : per full-res pixel group (4 px, 128-bit store): bilinear up-sampling of the four means and
//                      q = ((ma_r*I_r + ma_g*I_g) + ma_b*I_b) + mb, written over p (p is only read at the sample
//                      points by the earlier kernels, so the filter runs in place)
// HBM traffic per view ~ 4*V written + V/s^2 * 4 read + low-res planes: the pass is write-bound like CVC.
#pragma once
#include "psm_common.cuh"

namespace psm {

struct FgfGeom {
    int W, H, Wp, w2, h2, K, s;
    double ifx, ify;   // 1 / (w2 / W), 1 / (h2 / H): INTER_NN source index = min(floor(x * ifx), W - 1)
};

__device__ __forceinline__ int fgf_nn(int i, double inv, int n)
{
    const int v = (int)floor((double)i * inv);
    return v < n - 1 ? v : n - 1;
}

// low-res planes of one view: Ic[3], m[3], inv[6] (rr rg rb gg gb bb), each h2 x w2, contiguous
__global__ void fgf_guide_kernel(const float* __restrict__ I /* 3 full-res planes, pitch Wp */, size_t plane, FgfGeom g, float eps,
                                 float* __restrict__ lo /* 12 planes */)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= g.w2 || y >= g.h2) return;
    const int a = g.K / 2;
    const size_t n2 = (size_t)g.w2 * g.h2, o = (size_t)y * g.w2 + x;
    double sm[3] = {0, 0, 0}, sp[6] = {0, 0, 0, 0, 0, 0};
    for (int dy = -a; dy <= a; ++dy) {
        const int yy = fgf_nn(reflect101(y + dy, g.h2), g.ify, g.H);
        for (int dx = -a; dx <= a; ++dx) {
            const int xx = fgf_nn(reflect101(x + dx, g.w2), g.ifx, g.W);
            const size_t q = (size_t)yy * g.Wp + xx;
            const float c0 = I[q], c1 = I[plane + q], c2 = I[2 * plane + q];
            sm[0] += (double)c0; sm[1] += (double)c1; sm[2] += (double)c2;
            sp[0] += (double)fmul(c0, c0); sp[1] += (double)fmul(c0, c1); sp[2] += (double)fmul(c0, c2);
            sp[3] += (double)fmul(c1, c1); sp[4] += (double)fmul(c1, c2); sp[5] += (double)fmul(c2, c2);
        }
    }
    const double scale = 1.0 / ((double)g.K * g.K);
    const float m0 = (float)(sm[0] * scale), m1 = (float)(sm[1] * scale), m2 = (float)(sm[2] * scale);
    float bx[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) bx[k] = (float)(sp[k] * scale);
    // fastguidedfilter.cpp:144-149 (diagonal: addWeighted in double, one rounding)
    const float rr = (float)(((double)bx[0] - (double)fmul(m0, m0)) + (double)eps);
    const float rg = fsub(bx[1], fmul(m0, m1));
    const float rb = fsub(bx[2], fmul(m0, m2));
    const float gg = (float)(((double)bx[3] - (double)fmul(m1, m1)) + (double)eps);
    const float gb = fsub(bx[4], fmul(m1, m2));
    const float bb = (float)(((double)bx[5] - (double)fmul(m2, m2)) + (double)eps);
    const float irr = fsub(fmul(gg, bb), fmul(gb, gb));   // :152-157
    const float irg = fsub(fmul(gb, rb), fmul(rg, bb));
    const float irb = fsub(fmul(rg, gb), fmul(gg, rb));
    const float igg = fsub(fmul(rr, bb), fmul(rb, rb));
    const float igb = fsub(fmul(rb, rg), fmul(rr, gb));
    const float ibb = fsub(fmul(rr, gg), fmul(rg, rg));
    const float det = fadd(fadd(fmul(irr, rr), fmul(irg, rg)), fmul(irb, rb));   // :159
    const int ys = fgf_nn(y, g.ify, g.H), xs = fgf_nn(x, g.ifx, g.W);
    const size_t q = (size_t)ys * g.Wp + xs;
    lo[o] = I[q]; lo[n2 + o] = I[plane + q]; lo[2 * n2 + o] = I[2 * plane + q];
    lo[3 * n2 + o] = m0; lo[4 * n2 + o] = m1; lo[5 * n2 + o] = m2;
    lo[6 * n2 + o] = __fdiv_rn(irr, det); lo[7 * n2 + o] = __fdiv_rn(irg, det); lo[8 * n2 + o] = __fdiv_rn(irb, det);   // :161-166
    lo[9 * n2 + o] = __fdiv_rn(igg, det); lo[10 * n2 + o] = __fdiv_rn(igb, det); lo[11 * n2 + o] = __fdiv_rn(ibb, det);
}

// Both low-resolution stages of one slice tile in ONE kernel (shared-memory tiles, separable fp64 box sums):
//   region R1 = tile + 2h halo (h = K/2): p and I*p at the NN sample points, widened once into shared memory (doubles)
//   horizontal K-sums -> vertical K-sums -> box means on R2 = tile + h halo -> a (3), b there (floats, shared memory)
//   horizontal / vertical K-sums of a, b -> their box means on the tile -> global memory (4 planes per slice)
// Cells outside the low-res image hold the value of the BORDER_REFLECT_101 pixel; because reflect-101 is an even
// symmetry and the fp64 sums are exact, the box around a halo cell equals the box around the pixel it mirrors, so the
// second stage sees exactly cv::blur's border handling of the first stage's output.
constexpr int kFgfTX = 32, kFgfTY = 16, kFgfThreads = 256;

__host__ __device__ inline size_t fgf_smem_bytes(int K)
{
    const int h = K / 2;
    // row strides are made odd: the sliding sums run one thread per row, so consecutive lanes are one row stride apart
    const size_t r1 = (size_t)((kFgfTX + 4 * h) | 1) * (kFgfTY + 4 * h);    // inputs (4 planes of doubles)
    const size_t hs = (size_t)((kFgfTX + 2 * h) | 1) * (kFgfTY + 4 * h);    // horizontal sums (4 planes of doubles)
    const size_t r2 = (size_t)((kFgfTX + 2 * h) | 1) * (kFgfTY + 2 * h);    // a, b (4 planes of floats)
    return (r1 + hs) * 4 * sizeof(double) + r2 * 4 * sizeof(float);
}

__global__ void __launch_bounds__(kFgfThreads) fgf_lowres_kernel(const float* __restrict__ vol, size_t vplane, FgfGeom g,
                                                                 const float* __restrict__ lo, float* __restrict__ mean)
{
    extern __shared__ double fsm[];
    const int h = g.K / 2, K = g.K;
    const int W1n = kFgfTX + 4 * h, H1 = kFgfTY + 4 * h;     // input region (W1n columns used, odd stride W1)
    const int W2n = kFgfTX + 2 * h, H2 = kFgfTY + 2 * h;     // a,b region
    const int W1 = W1n | 1, W2 = W2n | 1;
    constexpr int kTXs = kFgfTX | 1;                          // stage-2 row stride
    double* in = fsm;                                        // [4][H1][W1]
    double* hs = in + (size_t)4 * W1 * H1;                   // [4][H1][W2]  (stage 2 reuses it as [4][H2][kTXs])
    float* ab = reinterpret_cast<float*>(hs + (size_t)4 * W2 * H1);   // [4][H2][W2]
    const int x0 = blockIdx.x * kFgfTX, y0 = blockIdx.y * kFgfTY, dl = blockIdx.z;
    const size_t n2 = (size_t)g.w2 * g.h2;
    const float* p = vol + (size_t)dl * vplane;
    const double scale = 1.0 / ((double)K * K);
    const int tx0 = threadIdx.x & 31, ty0 = threadIdx.x >> 5;   // 32 x 8 thread grid: rows by warps, columns by lanes (no runtime division)
    constexpr int kRowsPerPass = kFgfThreads / 32;
    // ---- inputs: p2 = resize(p, INTER_NN) (:69) and I*p2 (:175-177) on R1, widened once
    for (int ry = ty0; ry < H1; ry += kRowsPerPass) {
        const int ly = reflect101(y0 - 2 * h + ry, g.h2);
        const float* prow = p + (size_t)fgf_nn(ly, g.ify, g.H) * g.Wp;
        for (int rx = tx0; rx < W1n; rx += 32) {
            const int lx = reflect101(x0 - 2 * h + rx, g.w2);
            const float pv = __ldg(prow + fgf_nn(lx, g.ifx, g.W));
            const size_t q = (size_t)ly * g.w2 + lx;
            const int i = ry * W1 + rx;
            in[i] = (double)pv;
            in[(size_t)W1 * H1 + i] = (double)fmul(__ldg(lo + q), pv);
            in[(size_t)2 * W1 * H1 + i] = (double)fmul(__ldg(lo + n2 + q), pv);
            in[(size_t)3 * W1 * H1 + i] = (double)fmul(__ldg(lo + 2 * n2 + q), pv);
        }
    }
    __syncthreads();
    // Box sums as SLIDING fp64 sums (+ entering, - leaving): two shared-memory reads per output instead of K.  The sums
    // are exact (<= 81 floats of bounded range), so the sliding form gives the same bits as the direct form.
    // ---- stage 1, horizontal: one thread per (plane, row) slides along the W2 outputs
    for (int task = threadIdx.x; task < 4 * H1; task += kFgfThreads) {
        const double* src = in + (size_t)task * W1;      // task = plane * H1 + row
        double* dst = hs + (size_t)task * W2;
        double sacc = 0;
        for (int d = 0; d < K; ++d) sacc += src[d];
        dst[0] = sacc;
        for (int x = 1; x < W2n; ++x) { sacc = (sacc + src[x + K - 1]) - src[x - 1]; dst[x] = sacc; }
    }
    __syncthreads();
    // ---- stage 1, vertical: one thread per (plane, column) slides down the H2 outputs -> box means (floats) into `ab`
    for (int task = threadIdx.x; task < 4 * W2n; task += kFgfThreads) {
        const int k = task / W2n, rx = task - k * W2n;
        const double* src = hs + (size_t)k * H1 * W2 + rx;
        float* dst = ab + (size_t)k * W2 * H2 + rx;
        double sacc = 0;
        for (int d = 0; d < K; ++d) sacc += src[(size_t)d * W2];
        dst[0] = (float)(sacc * scale);
        for (int y = 1; y < H2; ++y) {
            sacc = (sacc + src[(size_t)(y + K - 1) * W2]) - src[(size_t)(y - 1) * W2];
            dst[(size_t)y * W2] = (float)(sacc * scale);
        }
    }
    __syncthreads();
    // ---- a, b on R2 from the four means of every cell, in place (fastguidedfilter.cpp:178-186)
    for (int ry = ty0; ry < H2; ry += kRowsPerPass) {
        const int ly = reflect101(y0 - h + ry, g.h2);
        for (int rx = tx0; rx < W2n; rx += 32) {
            const int i = ry * W2 + rx;
            const float mp = ab[i], mIp0 = ab[(size_t)W2 * H2 + i], mIp1 = ab[(size_t)2 * W2 * H2 + i], mIp2 = ab[(size_t)3 * W2 * H2 + i];
            const int lx = reflect101(x0 - h + rx, g.w2);
            const size_t o = (size_t)ly * g.w2 + lx;
            const float m0 = __ldg(lo + 3 * n2 + o), m1 = __ldg(lo + 4 * n2 + o), m2 = __ldg(lo + 5 * n2 + o);
            const float c0 = fsub(mIp0, fmul(m0, mp)), c1 = fsub(mIp1, fmul(m1, mp)), c2 = fsub(mIp2, fmul(m2, mp));
            const float irr = __ldg(lo + 6 * n2 + o), irg = __ldg(lo + 7 * n2 + o), irb = __ldg(lo + 8 * n2 + o);
            const float igg = __ldg(lo + 9 * n2 + o), igb = __ldg(lo + 10 * n2 + o), ibb = __ldg(lo + 11 * n2 + o);
            const float ar = fadd(fadd(fmul(irr, c0), fmul(irg, c1)), fmul(irb, c2));
            const float ag = fadd(fadd(fmul(irg, c0), fmul(igg, c1)), fmul(igb, c2));
            const float ab_ = fadd(fadd(fmul(irb, c0), fmul(igb, c1)), fmul(ibb, c2));
            const float bb = fsub(fsub(fsub(mp, fmul(ar, m0)), fmul(ag, m1)), fmul(ab_, m2));
            ab[i] = ar; ab[(size_t)W2 * H2 + i] = ag; ab[(size_t)2 * W2 * H2 + i] = ab_; ab[(size_t)3 * W2 * H2 + i] = bb;
        }
    }
    __syncthreads();
    // ---- stage 2, horizontal: (plane, row) tasks slide along the TX outputs of the tile
    for (int task = threadIdx.x; task < 4 * H2; task += kFgfThreads) {
        const float* src = ab + (size_t)task * W2;
        double* dst = hs + (size_t)task * kTXs;
        double sacc = 0;
        for (int d = 0; d < K; ++d) sacc += (double)src[d];
        dst[0] = sacc;
        for (int x = 1; x < kFgfTX; ++x) { sacc = (sacc + (double)src[x + K - 1]) - (double)src[x - 1]; dst[x] = sacc; }
    }
    __syncthreads();
    // ---- stage 2, vertical: (plane, column) tasks slide down the TY outputs -> means staged in `in` as floats [TY][TX][4]
    float* stage = reinterpret_cast<float*>(in);
    for (int task = threadIdx.x; task < 4 * kFgfTX; task += kFgfThreads) {
        const int k = task / kFgfTX, tx = task - k * kFgfTX;
        const double* src = hs + (size_t)k * H2 * kTXs + tx;
        double sacc = 0;
        for (int d = 0; d < K; ++d) sacc += src[(size_t)d * kTXs];
        stage[(0 * kFgfTX + tx) * 4 + k] = (float)(sacc * scale);
        for (int y = 1; y < kFgfTY; ++y) {
            sacc = (sacc + src[(size_t)(y + K - 1) * kTXs]) - src[(size_t)(y - 1) * kTXs];
            stage[(y * kFgfTX + tx) * 4 + k] = (float)(sacc * scale);
        }
    }
    __syncthreads();
    // ---- (mean_a_r, mean_a_g, mean_a_b, mean_b) of the tile, interleaved per pixel, coalesced 128-bit stores (:188-191)
    for (int ty = ty0; ty < kFgfTY; ty += kRowsPerPass) {
        const int x = x0 + tx0, y = y0 + ty;
        if (x >= g.w2 || y >= g.h2) continue;
        reinterpret_cast<float4*>(mean)[(size_t)dl * n2 + (size_t)y * g.w2 + x] = reinterpret_cast<const float4*>(stage)[ty * kFgfTX + tx0];
    }
}

// up-sampling plan (OpenCV's own INTER_LINEAR formula, built on the host): per destination column / row the two source
// indices and the float weight of the second one, 16 bytes each so that one 128-bit load fetches an entry
struct alignas(16) FgfTap { int i0, i1; float f, pad; };
struct FgfPlan { const FgfTap* x; const FgfTap* y; };

__global__ void __launch_bounds__(128) fgf_up_kernel(float* __restrict__ vol, size_t vplane, const float* __restrict__ I, size_t plane,
                                                     FgfGeom g, FgfPlan pl, const float* __restrict__ mean)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y, dl = blockIdx.z;
    if (x4 >= g.W) return;
    const size_t n2 = (size_t)g.w2 * g.h2;
    const float4* m = reinterpret_cast<const float4*>(mean) + (size_t)dl * n2;   // (ma_r, ma_g, ma_b, mb) per low-res pixel
    const FgfTap ty = pl.y[y];
    const float fy = ty.f, b0 = fsub(1.f, fy);
    const float4* row0 = m + (size_t)ty.i0 * g.w2;
    const float4* row1 = m + (size_t)ty.i1 * g.w2;
    const size_t o4 = (size_t)y * g.Wp + x4;
    float i0[4], i1[4], i2[4];
    if (x4 + 3 < g.Wp) {   // rows are padded to a multiple of 4 floats: a full 128-bit read is always inside the row
        const float4 a = __ldg(reinterpret_cast<const float4*>(I + o4)), b = __ldg(reinterpret_cast<const float4*>(I + plane + o4)),
                     c = __ldg(reinterpret_cast<const float4*>(I + 2 * plane + o4));
        i0[0] = a.x; i0[1] = a.y; i0[2] = a.z; i0[3] = a.w; i1[0] = b.x; i1[1] = b.y; i1[2] = b.z; i1[3] = b.w;
        i2[0] = c.x; i2[1] = c.y; i2[2] = c.z; i2[3] = c.w;
    }
    float q[4];
    int c0 = -1, c1 = -1;
    float4 s00 = make_float4(0.f, 0.f, 0.f, 0.f), s01 = s00, s10 = s00, s11 = s00;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x = x4 + j;
        if (x >= g.W) { q[j] = 0.f; continue; }
        const FgfTap tx = pl.x[x];
        const float fx = tx.f, a0 = fsub(1.f, fx);
        // consecutive pixels share their low-res column pair (s >= 2) or shift it by one: reload only what changed
        if (tx.i0 != c0 || tx.i1 != c1) {
            if (tx.i0 == c1) { s00 = s01; s10 = s11; }
            else { s00 = __ldg(row0 + tx.i0); s10 = __ldg(row1 + tx.i0); }
            s01 = __ldg(row0 + tx.i1); s11 = __ldg(row1 + tx.i1);
            c0 = tx.i0; c1 = tx.i1;
        }
        // rows after columns, like cv::resize: h = S0*(1-fx) + S1*fx per source row, then h0*(1-fy) + h1*fy
#define PSM_FGF_UP(c) fadd(fmul(fadd(fmul(s00.c, a0), fmul(s01.c, fx)), b0), fmul(fadd(fmul(s10.c, a0), fmul(s11.c, fx)), fy))
        const float ur = PSM_FGF_UP(x), ug = PSM_FGF_UP(y), ub = PSM_FGF_UP(z), ubb = PSM_FGF_UP(w);
#undef PSM_FGF_UP
        q[j] = fadd(fadd(fadd(fmul(ur, i0[j]), fmul(ug, i1[j])), fmul(ub, i2[j])), ubb);   // :196
    }
    float* dst = vol + (size_t)dl * vplane + o4;
    if (x4 + 3 < g.W) __stcs(reinterpret_cast<float4*>(dst), make_float4(q[0], q[1], q[2], q[3]));
    else for (int j = 0; j < 4 && x4 + j < g.W; ++j) dst[j] = q[j];
}

}  // namespace psm
