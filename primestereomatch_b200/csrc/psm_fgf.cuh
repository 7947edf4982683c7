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
//   fgf_ab_kernel    : per slice, per low-res pixel: p and I*p box means -> a (3), b            (reads 1/s^2 of p)
//   fgf_mean_kernel  : per slice: K x K box means of a, b at low resolution
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

// a (3), b of every owned slice at low resolution: ab[(slice * 4 + k) * n2 + pixel]
__global__ void fgf_ab_kernel(const float* __restrict__ vol, size_t vplane, FgfGeom g, const float* __restrict__ lo, float* __restrict__ ab)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, dl = blockIdx.z;
    if (x >= g.w2 || y >= g.h2) return;
    const int a = g.K / 2;
    const size_t n2 = (size_t)g.w2 * g.h2, o = (size_t)y * g.w2 + x;
    const float* p = vol + (size_t)dl * vplane;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int dy = -a; dy <= a; ++dy) {
        const int ly = reflect101(y + dy, g.h2);
        const int yy = fgf_nn(ly, g.ify, g.H);
        for (int dx = -a; dx <= a; ++dx) {
            const int lx = reflect101(x + dx, g.w2);
            const float pv = __ldg(p + (size_t)yy * g.Wp + fgf_nn(lx, g.ifx, g.W));   // resize(p, INTER_NN), :69
            const size_t q = (size_t)ly * g.w2 + lx;
            s0 += (double)pv;
            s1 += (double)fmul(__ldg(lo + q), pv);
            s2 += (double)fmul(__ldg(lo + n2 + q), pv);
            s3 += (double)fmul(__ldg(lo + 2 * n2 + q), pv);
        }
    }
    const double scale = 1.0 / ((double)g.K * g.K);
    const float mp = (float)(s0 * scale), mIp0 = (float)(s1 * scale), mIp1 = (float)(s2 * scale), mIp2 = (float)(s3 * scale);
    const float m0 = lo[3 * n2 + o], m1 = lo[4 * n2 + o], m2 = lo[5 * n2 + o];
    const float c0 = fsub(mIp0, fmul(m0, mp)), c1 = fsub(mIp1, fmul(m1, mp)), c2 = fsub(mIp2, fmul(m2, mp));   // :178-180
    const float irr = lo[6 * n2 + o], irg = lo[7 * n2 + o], irb = lo[8 * n2 + o], igg = lo[9 * n2 + o], igb = lo[10 * n2 + o], ibb = lo[11 * n2 + o];
    const float ar = fadd(fadd(fmul(irr, c0), fmul(irg, c1)), fmul(irb, c2));   // :182-184
    const float ag = fadd(fadd(fmul(irg, c0), fmul(igg, c1)), fmul(igb, c2));
    const float ab_ = fadd(fadd(fmul(irb, c0), fmul(igb, c1)), fmul(ibb, c2));
    const float b = fsub(fsub(fsub(mp, fmul(ar, m0)), fmul(ag, m1)), fmul(ab_, m2));   // :186
    float* out = ab + (size_t)dl * 4 * n2 + o;
    out[0] = ar; out[n2] = ag; out[2 * n2] = ab_; out[3 * n2] = b;
}

// K x K box means of the four coefficient planes of every slice
__global__ void fgf_mean_kernel(const float* __restrict__ ab, FgfGeom g, float* __restrict__ mean)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= g.w2 || y >= g.h2) return;
    const int a = g.K / 2;
    const size_t n2 = (size_t)g.w2 * g.h2;
    const float* src = ab + (size_t)blockIdx.z * n2;   // blockIdx.z = slice * 4 + plane
    double s = 0;
    for (int dy = -a; dy <= a; ++dy) {
        const float* row = src + (size_t)reflect101(y + dy, g.h2) * g.w2;
        for (int dx = -a; dx <= a; ++dx) s += (double)__ldg(row + reflect101(x + dx, g.w2));
    }
    mean[(size_t)blockIdx.z * n2 + (size_t)y * g.w2 + x] = (float)(s * (1.0 / ((double)g.K * g.K)));
}

// up-sampling plan of one axis: i0, i1 (low-res indices) and the float weight f of i1 (OpenCV's own formula, built on the host)
struct FgfPlan { const int* x0; const int* x1; const float* fx; const int* y0; const int* y1; const float* fy; };

__global__ void __launch_bounds__(128) fgf_up_kernel(float* __restrict__ vol, size_t vplane, const float* __restrict__ I, size_t plane,
                                                     FgfGeom g, FgfPlan pl, const float* __restrict__ mean)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y, dl = blockIdx.z;
    if (x4 >= g.W) return;
    const size_t n2 = (size_t)g.w2 * g.h2;
    const float* m = mean + (size_t)dl * 4 * n2;
    const int r0 = pl.y0[y], r1 = pl.y1[y];
    const float fy = pl.fy[y], b0 = fsub(1.f, fy);
    float q[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x = x4 + j;
        if (x >= g.W) { q[j] = 0.f; continue; }
        const int c0 = pl.x0[x], c1 = pl.x1[x];
        const float fx = pl.fx[x], a0 = fsub(1.f, fx);
        float up[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float* mk = m + (size_t)k * n2;
            const float h0 = fadd(fmul(__ldg(mk + (size_t)r0 * g.w2 + c0), a0), fmul(__ldg(mk + (size_t)r0 * g.w2 + c1), fx));
            const float h1 = fadd(fmul(__ldg(mk + (size_t)r1 * g.w2 + c0), a0), fmul(__ldg(mk + (size_t)r1 * g.w2 + c1), fx));
            up[k] = fadd(fmul(h0, b0), fmul(h1, fy));   // rows after columns, like cv::resize
        }
        const size_t o = (size_t)y * g.Wp + x;
        q[j] = fadd(fadd(fadd(fmul(up[0], I[o]), fmul(up[1], I[plane + o])), fmul(up[2], I[2 * plane + o])), up[3]);   // :196
    }
    float* dst = vol + (size_t)dl * vplane + (size_t)y * g.Wp + x4;
    if (x4 + 3 < g.W) __stcs(reinterpret_cast<float4*>(dst), make_float4(q[0], q[1], q[2], q[3]));
    else for (int j = 0; j < 4 && x4 + j < g.W; ++j) dst[j] = q[j];
}

}  // namespace psm
