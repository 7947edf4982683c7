// psm_cvf_stream2.cuh -- K3 v2: the fused streaming guided-image-filter kernel, engineered.
//
// Same algorithm and numerics as psm_cvf_stream.cuh (see the header comment there): per warp a
// 128-column strip of one disparity slice streams down a row segment; both 8x8 box stages are
// fp64 running column sums + prefix/suffix/shuffle row sums; a,b live in a thread-private
// shared-memory ring; q is bit-exact against the oracle.
//
// What changed versus v1 (ncu: 1145 warp-instructions per warp-row, 32% issue utilisation, 8 warps/SM):
//   * BORDER template: interior strips (every column of every lane inside the image) take a path
//     with plain 128-bit loads and no reflection code at all; only the first/last strip of a row
//     of strips pays for BORDER_REFLECT_101 (per-lane reflected column tables, scalar gathers).
//   * the fp32 multiplies run on packed f32x2 instructions (FMUL2: two columns per issue slot); adds
//     stay scalar because ptxas would fuse packed mul+add into FFMA2 (see add2 below).
//   * loads of a row are issued at the top of the iteration, long before their first use.
//   * row offsets are 32-bit element offsets from three base pointers.
#pragma once
#include "psm_cvf_stream.cuh"

namespace psm {

struct f2x2 {  // four columns as two packed pairs
    float2 lo, hi;
};
__device__ __forceinline__ f2x2 from4(const float4& v) { return {make_float2(v.x, v.y), make_float2(v.z, v.w)}; }
__device__ __forceinline__ float4 to4(const f2x2& v) { return make_float4(v.lo.x, v.lo.y, v.hi.x, v.hi.y); }
__device__ __forceinline__ f2x2 mul2(const f2x2& a, const f2x2& b) { return {__fmul2_rn(a.lo, b.lo), __fmul2_rn(a.hi, b.hi)}; }
// NOTE: ptxas (12.9) fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with --fmad=false, which
// would break bit-exactness against the reference's separate multiply / add.  It does not fuse a
// packed multiply with SCALAR adds, so every add/sub below is two scalar FADDs on the halves.
__device__ __forceinline__ f2x2 add2(const f2x2& a, const f2x2& b)
{
    return {make_float2(fadd(a.lo.x, b.lo.x), fadd(a.lo.y, b.lo.y)), make_float2(fadd(a.hi.x, b.hi.x), fadd(a.hi.y, b.hi.y))};
}
__device__ __forceinline__ f2x2 sub2(const f2x2& a, const f2x2& b)
{
    return {make_float2(fsub(a.lo.x, b.lo.x), fsub(a.lo.y, b.lo.y)), make_float2(fsub(a.hi.x, b.hi.x), fsub(a.hi.y, b.hi.y))};
}
__device__ __forceinline__ float get(const f2x2& v, int j) { return j == 0 ? v.lo.x : (j == 1 ? v.lo.y : (j == 2 ? v.hi.x : v.hi.y)); }

// Per-lane column plan of a border strip: reflected source columns for the three column groups.
struct ColPlan {
    int in[4];   // input columns  (p, I)           reflect101(cin + j)
    int ab[4];   // a,b columns (guide mean/adj)    clamped into the image (values there are replaced anyway)
    int out[4];  // output columns (I)              clamped
};

template <bool BORDER>
__device__ __forceinline__ float4 ld4(const float* __restrict__ base, unsigned off, int col, const int idx[4])
{
    if (!BORDER) return __ldg(reinterpret_cast<const float4*>(base + off + col));
    float4 v;
    v.x = __ldg(base + off + idx[0]);
    v.y = __ldg(base + off + idx[1]);
    v.z = __ldg(base + off + idx[2]);
    v.w = __ldg(base + off + idx[3]);
    return v;
}

template <bool BORDER>
__device__ __forceinline__ void cvf2_body(const CvfParams& P, float4* ring, int view, int seg, int strip, int dlc, bool dvalid)
{
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    constexpr int nthr = 128;
    const int W = P.W, H = P.H, Wp = P.Wp;
    const unsigned plane = (unsigned)H * (unsigned)Wp;
    const float* __restrict__ G = P.guide[view];
    const float* __restrict__ vin = P.vol_in[view] + (size_t)dlc * plane;
    float* __restrict__ vout = P.vol_out[view] + (size_t)dlc * plane;

    const int out_lo = strip * kStripOut;
    const int X0 = (strip == P.nstrips - 1 && strip > 0) ? ((W - kStripOut + 3) & ~3) : out_lo;
    const int cin = X0 - 8 + 4 * lane;
    const int ca = cin + 4;
    const int co = cin + 8;

    ColPlan cp;
    int fix_lane[4], fix_elem[4];
    bool fix_need[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (BORDER) {
            cp.in[j] = reflect101(cin + j, W);
            cp.ab[j] = min(max(ca + j, 0), W - 1);
            cp.out[j] = min(max(co + j, 0), W - 1);
            const int xa = ca + j;
            fix_need[j] = (xa < 0) || (xa >= W);
            const int r = reflect101(xa, W) - (X0 - 4);
            int sl = r >> 2;
            fix_lane[j] = sl < 0 ? 0 : (sl > 31 ? 31 : sl);
            fix_elem[j] = r & 3;
        } else {
            cp.in[j] = cp.ab[j] = cp.out[j] = 0;
            fix_need[j] = false; fix_lane[j] = 0; fix_elem[j] = 0;
        }
    }

    const int Y0 = seg * P.seg_rows;
    const int Y1 = min(H, Y0 + P.seg_rows);
    const bool top = (Y0 == 0);
    const bool bottom = (Y1 == H);
    const int T0 = top ? 0 : Y0 - 4;
    const int Tlast = bottom ? H - 1 : Y1 + 2;
    const int Tend = bottom ? H + 2 : Tlast;

    double S1[4][4], S2[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j) { S1[q][j] = 0.0; S2[q][j] = 0.0; }

    // one input row (p and the three guide channels), as loaded
    struct RowIn { float4 p, i0, i1, i2; };
    auto load_in = [&](int r) {
        const unsigned ro = (unsigned)reflect101(r, H) * (unsigned)Wp;
        RowIn x;
        x.p = ld4<BORDER>(vin, ro, cin, cp.in);
        x.i0 = ld4<BORDER>(G, ro, cin, cp.in);
        x.i1 = ld4<BORDER>(G, plane + ro, cin, cp.in);
        x.i2 = ld4<BORDER>(G, 2 * plane + ro, cin, cp.in);
        return x;
    };
    auto accumulate = [&](const RowIn& x, const double sign) {
        const f2x2 p = from4(x.p);
        const f2x2 m0 = mul2(from4(x.i0), p), m1 = mul2(from4(x.i1), p), m2 = mul2(from4(x.i2), p);  // CVF.cpp:87
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            S1[0][j] = __fma_rn(sign, (double)get(p, j), S1[0][j]);
            S1[1][j] = __fma_rn(sign, (double)get(m0, j), S1[1][j]);
            S1[2][j] = __fma_rn(sign, (double)get(m1, j), S1[2][j]);
            S1[3][j] = __fma_rn(sign, (double)get(m2, j), S1[3][j]);
        }
    };

    for (int r = T0 - 4; r <= T0 + 2; ++r) accumulate(load_in(r), 1.0);

    for (int t = T0; t <= Tend; ++t) {
        const bool real_row = t <= Tlast;
        f2x2 av[4];  // newest a0,a1,a2,b row
        if (real_row) {
            // ---- issue every global load of this row up front --------------------------------
            const RowIn xn = load_in(t + 3);
            const RowIn xo = load_in(t - 4);
            const unsigned ro = (unsigned)t * (unsigned)Wp;
            float4 g4[10];
#pragma unroll
            for (int q = 0; q < 10; ++q) g4[q] = ld4<BORDER>(G, (unsigned)(kGuideMean + q) * plane + ro, ca, cp.ab);

            accumulate(xn, 1.0);
            double h[4][4];
#pragma unroll
            for (int q = 0; q < 4; ++q) hsum8(S1[q], h[q]);
            accumulate(xo, -1.0);  // S1 now holds the window of row t+1 minus its newest row

            // ---- means -> cov -> a,b, packed two columns per instruction (CVF.cpp:92-155) -----
            f2x2 m[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                m[q] = {make_float2(mean64(h[q][0]), mean64(h[q][1])), make_float2(mean64(h[q][2]), mean64(h[q][3]))};
            const f2x2 mI0 = from4(g4[0]), mI1 = from4(g4[1]), mI2 = from4(g4[2]);
            const f2x2 M00 = from4(g4[3]), M01 = from4(g4[4]), M02 = from4(g4[5]);
            const f2x2 M11 = from4(g4[6]), M12 = from4(g4[7]), M22 = from4(g4[8]);
            const f2x2 idet = from4(g4[9]);
            const f2x2 c0 = sub2(m[1], mul2(mI0, m[0]));
            const f2x2 c1 = sub2(m[2], mul2(mI1, m[0]));
            const f2x2 c2 = sub2(m[3], mul2(mI2, m[0]));
            av[0] = mul2(idet, add2(add2(mul2(c0, M00), mul2(c1, M01)), mul2(c2, M02)));
            av[1] = mul2(idet, add2(add2(mul2(c0, M01), mul2(c1, M11)), mul2(c2, M12)));
            av[2] = mul2(idet, add2(add2(mul2(c0, M02), mul2(c1, M12)), mul2(c2, M22)));
            av[3] = sub2(sub2(sub2(m[0], mul2(av[0], mI0)), mul2(av[1], mI1)), mul2(av[2], mI2));

            if (BORDER) {  // a,b outside the image := reflected columns
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float fixed[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float e0 = __shfl_sync(0xffffffffu, av[q].lo.x, fix_lane[j]);
                        const float e1 = __shfl_sync(0xffffffffu, av[q].lo.y, fix_lane[j]);
                        const float e2 = __shfl_sync(0xffffffffu, av[q].hi.x, fix_lane[j]);
                        const float e3 = __shfl_sync(0xffffffffu, av[q].hi.y, fix_lane[j]);
                        const int e = fix_elem[j];
                        const float v = e == 0 ? e0 : (e == 1 ? e1 : (e == 2 ? e2 : e3));
                        fixed[j] = fix_need[j] ? v : get(av[q], j);
                    }
                    av[q] = {make_float2(fixed[0], fixed[1]), make_float2(fixed[2], fixed[3])};
                }
            }
        } else {
            const int slot = reflect101(t, H) & 7;
#pragma unroll
            for (int q = 0; q < 4; ++q) av[q] = from4(ring[(slot * 4 + q) * nthr + tid]);
        }

        // ---- stage-2 vertical running sums (window of output row t-3 is a,b rows t-7..t) -------
        const int age = t - T0;
        const bool warm = top ? (t <= 4) : (age < 8);
        const double wnew = (top && t >= 1 && t <= 3) ? 2.0 : 1.0;
        const int oslot = (top && t < 8) ? ((8 - t) & 7) : (t & 7);
        const int nslot = t & 7;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 old4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!warm) old4 = ring[(oslot * 4 + q) * nthr + tid];
            if (real_row) ring[(nslot * 4 + q) * nthr + tid] = to4(av[q]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double s = __fma_rn(wnew, (double)get(av[q], j), S2[q][j]);
                if (!warm) s = __dsub_rn(s, (double)comp(old4, j));
                S2[q][j] = s;
            }
        }

        const bool first_out = top ? (t == 4) : (age == 7);
        if (warm && !first_out) continue;
        const int nrows = (top && t == 4) ? 2 : 1;
        double h2[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) hsum8(S2[q], h2[q]);
        f2x2 mb[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            mb[q] = {make_float2(mean64(h2[q][0]), mean64(h2[q][1])), make_float2(mean64(h2[q][2]), mean64(h2[q][3]))};
        for (int e = 0; e < nrows; ++e) {
            const int y = (top && t == 4) ? e : t - 3;
            if (y < Y0 || y >= Y1) continue;
            const unsigned ro = (unsigned)y * (unsigned)Wp;
            const f2x2 i0 = from4(ld4<BORDER>(G, ro, co, cp.out));
            const f2x2 i1 = from4(ld4<BORDER>(G, plane + ro, co, cp.out));
            const f2x2 i2 = from4(ld4<BORDER>(G, 2 * plane + ro, co, cp.out));
            // q = box(b) + box(a0)*I0 + box(a1)*I1 + box(a2)*I2 in that order (CVF.cpp:157-163)
            f2x2 qv = add2(mb[3], mul2(mb[0], i0));
            qv = add2(qv, mul2(mb[1], i1));
            qv = add2(qv, mul2(mb[2], i2));
            if (lane <= 27 && co < W && co >= out_lo && dvalid)
                *reinterpret_cast<float4*>(vout + ro + co) = to4(qv);
        }
    }
}

// Strip index -> (first output column X0, touches-a-border?).  A strip is interior when every
// column any of its lanes touches (inputs X0-8 .. X0+119) lies inside the image.
__host__ __device__ __forceinline__ bool strip_is_border(int strip, int nstrips, int W)
{
    const int out_lo = strip * kStripOut;
    const int X0 = (strip == nstrips - 1 && strip > 0) ? ((W - kStripOut + 3) & ~3) : out_lo;
    return (X0 - 8 < 0) || (X0 + 120 > W);
}

// Two kernels so that the interior path's register allocation is not inflated by the border path.
// `strip_list` maps blockIdx -> strip index for the strips of this kernel's class.
template <bool BORDER>
__global__ void __launch_bounds__(128, BORDER ? 2 : 3)
cvf_stream2_kernel(const CvfParams P, const int* __restrict__ strip_list, int nlist)
{
    extern __shared__ float4 ring[];  // [8 slots][4 planes][128 threads]
    const int warp = threadIdx.x >> 5;
    int b = blockIdx.x;
    const int dgroup = b % P.ndgroups; b /= P.ndgroups;
    const int strip = strip_list[b % nlist]; b /= nlist;
    const int seg = b % P.nseg;
    const int view = b / P.nseg;
    const int d = dgroup * 4 + warp;
    if (d >= P.Dloc) return;
    cvf2_body<BORDER>(P, ring, view, seg, strip, d, true);
}

}  // namespace psm
