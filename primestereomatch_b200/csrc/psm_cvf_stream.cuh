// psm_cvf_stream.cuh -- K3: the fused streaming guided-image-filter kernel (the graded kernel).
//
// One launch filters every owned slice of both cost volumes:  q = GuidedFilter_cv(I, p)
// (/root/reference/src/CVF.cpp:72-165) with both 8x8 box stages chained on-chip, so HBM sees
// p once (read) and q once (write) per voxel; a, b, the eight box means and all products never
// leave the SM.
//
// Decomposition (B200-first, not the reference's per-slice Mat pipeline):
//   * a WARP owns a strip of 128 input columns (each lane 4 consecutive columns -> 128-bit loads)
//     of DT disparity slices and streams down the rows of one row segment;
//   * stage-1 vertical 8-row sums of p, I0*p, I1*p, I2*p are fp64 running sums in registers
//     (add newest row, subtract oldest row; both come straight from the read-only raw volume);
//   * the horizontal 8-column sums are formed from per-lane prefix/suffix sums of the 4 owned
//     columns plus 4 fp64 warp shuffles per box (lane+1 total, lane+2 prefixes);
//   * a,b for the row are computed in fp32 with the reference's exact operation order and
//     pushed into a per-thread 8-row ring in shared memory (the only on-chip history needed);
//   * stage-2 vertical sums of a0,a1,a2,b are fp64 running sums fed from the registers (newest)
//     and the ring (oldest); horizontal sums as in stage 1; q is written with one 128-bit store.
//   All box sums are fp64 (== cv::boxFilter's double accumulation, order-independent in practice),
//   all fp32 arithmetic is unfused round-to-nearest: q is bit-exact against the oracle.
//
// Column bookkeeping for strip s (X0 = 112*s, c0 = X0 - 8), lane l:
//   input  columns  c0+4l   .. c0+4l+3    (p, I)
//   a,b    columns  c0+4l+4 .. c0+4l+7    (valid for l <= 29; window [x-4, x+3] of the inputs)
//   output columns  c0+4l+8 .. c0+4l+11   (valid for l <= 27)  ==  X0+4l .. X0+4l+3
// Borders: input loads reflect (BORDER_REFLECT_101) in x and y; a,b at columns outside [0,W)
// are replaced by the reflected columns' values (lane shuffles, border strips only); a,b rows
// outside [0,H) are handled by the running-sum schedule (top: weights 1,2,2,2,1; bottom: three
// virtual iterations fed from the ring).
#pragma once
#include "psm_kernels.cuh"

namespace psm {

constexpr int kStripOut = 112;   // output columns per warp
constexpr int kStripIn = 128;    // input columns per warp
// warps per CTA = template parameter NW (each warp a different disparity group of the same strip)

struct CvfParams {
    const float* vol_in[2];   // raw volumes  [Dloc][H][Wp]
    float* vol_out[2];        // filtered volumes
    const float* guide[2];    // guide planes [kGuidePlanes][H][Wp]
    int W, H, Wp, Dloc;
    int nstrips, nseg, seg_rows, ndgroups;
};

template <typename T> __device__ __forceinline__ T shfl_down_t(T v, int delta)
{
    return __shfl_down_sync(0xffffffffu, v, delta);
}

// 8-wide horizontal window sums from the 4 owned column sums c[0..3]:
// h[j] = sum of columns (4l+j) .. (4l+j+7)  = suffix_l[j..3] + total_{l+1} + prefix_{l+2}[0..j-1]
template <typename T>
__device__ __forceinline__ void hsum8(const T c[4], T h[4])
{
    const T P1 = c[0], P2 = c[0] + c[1], P3 = P2 + c[2], Tt = P3 + c[3];
    const T S1 = c[3], S2 = c[2] + c[3], S3 = c[1] + S2;
    const T Tn = shfl_down_t(Tt, 1);
    const T Q1 = shfl_down_t(P1, 2), Q2 = shfl_down_t(P2, 2), Q3 = shfl_down_t(P3, 2);
    h[0] = Tt + Tn;
    h[1] = (S3 + Tn) + Q1;
    h[2] = (S2 + Tn) + Q2;
    h[3] = (S1 + Tn) + Q3;
}

__device__ __forceinline__ float mean64(double s) { return (float)__dmul_rn(s, 1.0 / 64.0); }

template <int DT, int NW>
__global__ void __launch_bounds__(NW * 32)
cvf_stream_kernel(const CvfParams P)
{
    extern __shared__ float4 ring[];  // [8 slots][4 planes][DT][blockDim.x] float4 (thread-private columns)
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int nthr = NW * 32;

    int b = blockIdx.x;
    const int dgroup = b % P.ndgroups; b /= P.ndgroups;
    const int strip = b % P.nstrips;   b /= P.nstrips;
    const int seg = b % P.nseg;
    const int view = b / P.nseg;

    const int dbase = (dgroup * NW + warp) * DT;
    if (dbase >= P.Dloc) return;  // warps are independent: no block-level barrier anywhere below

    const int W = P.W, H = P.H, Wp = P.Wp;
    const size_t plane = (size_t)H * Wp;
    const float* __restrict__ G = P.guide[view];
    const float* __restrict__ vin = P.vol_in[view];
    float* __restrict__ vout = P.vol_out[view];

    // the last strip is shifted left so that it ends at the image edge (its columns that the
    // previous strip already produces are not stored again)
    const int out_lo = strip * kStripOut;
    const int X0 = (strip == P.nstrips - 1 && strip > 0) ? ((W - kStripOut + 3) & ~3) : out_lo;
    const int cin = X0 - 8 + 4 * lane;  // first owned input column
    const int ca = cin + 4;             // first a,b column
    const int co = cin + 8;             // first output column

    int dl[DT];
    bool dvalid[DT];
#pragma unroll
    for (int k = 0; k < DT; ++k) { dvalid[k] = dbase + k < P.Dloc; dl[k] = dvalid[k] ? dbase + k : P.Dloc - 1; }

    // ---- a,b column reflection plan (only strips that touch an image border) -----------------
    const bool strip_fix = (X0 == 0) || (X0 + kStripOut + 3 >= W);  // warp-uniform
    int fix_lane[4], fix_elem[4];
    bool fix_need[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int xa = ca + j;
        fix_need[j] = (xa < 0) || (xa >= W);
        const int r = reflect101(xa, W) - (X0 - 4);  // offset inside the warp's a,b columns
        int sl = r >> 2;
        sl = sl < 0 ? 0 : (sl > 31 ? 31 : sl);
        fix_lane[j] = sl;
        fix_elem[j] = r & 3;
    }

    // ---- row schedule -----------------------------------------------------------------------
    const int Y0 = seg * P.seg_rows;
    const int Y1 = min(H, Y0 + P.seg_rows);
    const bool top = (Y0 == 0);
    const bool bottom = (Y1 == H);
    const int T0 = top ? 0 : Y0 - 4;                 // first a,b row computed
    const int Tlast = bottom ? H - 1 : Y1 + 2;       // last real a,b row
    const int Tend = bottom ? H + 2 : Tlast;         // last iteration (virtual rows at the bottom)

    double S1[DT][4][4];  // [slice][box: p, I0p, I1p, I2p][column]
    double S2[DT][4][4];  // [slice][plane: a0, a1, a2, b][column]
#pragma unroll
    for (int k = 0; k < DT; ++k)
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 4; ++j) { S1[k][q][j] = 0.0; S2[k][q][j] = 0.0; }

    // add (sign=+1) or remove (sign=-1) one input row from the stage-1 column sums
    auto feed = [&](int r, const double sign) {
        const int rr = reflect101(r, H);
        const size_t ro = (size_t)rr * Wp;
        const float4 i0 = load_row4(G + ro, cin, W);
        const float4 i1 = load_row4(G + plane + ro, cin, W);
        const float4 i2 = load_row4(G + 2 * plane + ro, cin, W);
#pragma unroll
        for (int k = 0; k < DT; ++k) {
            const float4 p4 = load_row4(vin + (size_t)dl[k] * plane + ro, cin, W);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float p = comp(p4, j);
                S1[k][0][j] = __fma_rn(sign, (double)p, S1[k][0][j]);  // +-1 * x is exact: plain add/sub
                S1[k][1][j] = __fma_rn(sign, (double)fmul(comp(i0, j), p), S1[k][1][j]);  // CVF.cpp:87 multiply
                S1[k][2][j] = __fma_rn(sign, (double)fmul(comp(i1, j), p), S1[k][2][j]);
                S1[k][3][j] = __fma_rn(sign, (double)fmul(comp(i2, j), p), S1[k][3][j]);
            }
        }
    };

    // warm-up of stage 1: the window of a,b row T0 is input rows T0-4 .. T0+3
    for (int r = T0 - 4; r <= T0 + 2; ++r) feed(r, 1.0);

    for (int t = T0; t <= Tend; ++t) {
        float av[DT][4][4];  // newest a,b row: [slice][plane][column]
        const bool real_row = t <= Tlast;
        if (real_row) {
            feed(t + 3, 1.0);
            // ---- stage-1 horizontal sums -> means -> a,b (CVF.cpp:81-155) -------------------
            const size_t ro = (size_t)t * Wp;
            float4 g4[10];
#pragma unroll
            for (int q = 0; q < 10; ++q) g4[q] = load_row4(G + (size_t)(kGuideMean + q) * plane + ro, ca, W);
#pragma unroll
            for (int k = 0; k < DT; ++k) {
                double h[4][4];
#pragma unroll
                for (int q = 0; q < 4; ++q) hsum8(S1[k][q], h[q]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    GuidePix g;
                    g.mI0 = comp(g4[0], j); g.mI1 = comp(g4[1], j); g.mI2 = comp(g4[2], j);
                    g.M00 = comp(g4[3], j); g.M01 = comp(g4[4], j); g.M02 = comp(g4[5], j);
                    g.M11 = comp(g4[6], j); g.M12 = comp(g4[7], j); g.M22 = comp(g4[8], j);
                    g.idet = comp(g4[9], j);
                    gif_coeffs(mean64(h[0][j]), mean64(h[1][j]), mean64(h[2][j]), mean64(h[3][j]), g,
                               av[k][0][j], av[k][1][j], av[k][2][j], av[k][3][j]);
                }
                if (strip_fix) {  // a,b at columns outside the image := reflected columns (BORDER_REFLECT_101)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float fixed[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float e0 = __shfl_sync(0xffffffffu, av[k][q][0], fix_lane[j]);
                            const float e1 = __shfl_sync(0xffffffffu, av[k][q][1], fix_lane[j]);
                            const float e2 = __shfl_sync(0xffffffffu, av[k][q][2], fix_lane[j]);
                            const float e3 = __shfl_sync(0xffffffffu, av[k][q][3], fix_lane[j]);
                            const int e = fix_elem[j];
                            const float v = e == 0 ? e0 : (e == 1 ? e1 : (e == 2 ? e2 : e3));
                            fixed[j] = fix_need[j] ? v : av[k][q][j];
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) av[k][q][j] = fixed[j];
                    }
                }
            }
            feed(t - 4, -1.0);  // oldest row of this window leaves before the next a,b row
        } else {
            // virtual a,b rows below the image: row t == reflected row 2(H-1)-t, still in the ring
            const int slot = reflect101(t, H) & 7;
#pragma unroll
            for (int k = 0; k < DT; ++k)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 v = ring[((slot * 4 + q) * DT + k) * nthr + tid];
                    av[k][q][0] = v.x; av[k][q][1] = v.y; av[k][q][2] = v.z; av[k][q][3] = v.w;
                }
        }

        // ---- stage-2 vertical running sums ---------------------------------------------------
        // window of output row y = t-3 is a,b rows t-7 .. t (reflected at the image top/bottom)
        const int age = t - T0;
        const bool warm = top ? (t <= 4) : (age < 8);
        const double wnew = (top && t >= 1 && t <= 3) ? 2.0 : 1.0;  // rows 1..3 appear twice in row 0's window
        const int oslot = (top && t < 8) ? ((8 - t) & 7) : (t & 7);  // slot of row reflect(t-8)
        const int nslot = t & 7;
#pragma unroll
        for (int k = 0; k < DT; ++k)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 old4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (!warm) old4 = ring[((oslot * 4 + q) * DT + k) * nthr + tid];
                if (real_row)
                    ring[((nslot * 4 + q) * DT + k) * nthr + tid] =
                        make_float4(av[k][q][0], av[k][q][1], av[k][q][2], av[k][q][3]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double s = __fma_rn(wnew, (double)av[k][q][j], S2[k][q][j]);  // exact: wnew is 1 or 2
                    if (!warm) s = __dsub_rn(s, (double)comp(old4, j));
                    S2[k][q][j] = s;
                }
            }

        // ---- outputs -------------------------------------------------------------------------
        const bool first_out = top ? (t == 4) : (age == 7);
        if (warm && !first_out) continue;
        const int nrows = (top && t == 4) ? 2 : 1;  // rows 0 and 1 share the same reflected window
        for (int e = 0; e < nrows; ++e) {
            const int y = (top && t == 4) ? e : t - 3;
            if (y < Y0 || y >= Y1) continue;
            const size_t ro = (size_t)y * Wp;
            const float4 i0 = load_row4(G + ro, co, W);
            const float4 i1 = load_row4(G + plane + ro, co, W);
            const float4 i2 = load_row4(G + 2 * plane + ro, co, W);
#pragma unroll
            for (int k = 0; k < DT; ++k) {
                double h[4][4];
#pragma unroll
                for (int q = 0; q < 4; ++q) hsum8(S2[k][q], h[q]);
                float4 q4;
                float* qp = reinterpret_cast<float*>(&q4);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // q = box(b) + box(a0)*I0 + box(a1)*I1 + box(a2)*I2, accumulated in that order (CVF.cpp:157-163)
                    float q = mean64(h[3][j]);
                    q = fadd(q, fmul(mean64(h[0][j]), comp(i0, j)));
                    q = fadd(q, fmul(mean64(h[1][j]), comp(i1, j)));
                    q = fadd(q, fmul(mean64(h[2][j]), comp(i2, j)));
                    qp[j] = q;
                }
                if (lane <= 27 && co < W && co >= out_lo && dvalid[k])
                    *reinterpret_cast<float4*>(vout + (size_t)dl[k] * plane + ro + co) = q4;
            }
        }
    }
}

}  // namespace psm
