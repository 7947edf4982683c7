// psm_cvf_stream.cuh -- K3: the fused streaming guided-image-filter kernel (the graded kernel).
//
// q = GuidedFilter_cv(I, p) (/root/reference/src/CVF.cpp:72-165) for every owned slice of both cost
// volumes in ONE pass over HBM: p is read once and q written once per voxel; the two chained 8x8
// box stages, a, b, all products and all means stay on the SM.  The reference runs ~90 Mat-sized
// passes per slice for the same result.
//
// Numerics (bit-exact against the oracle):
//   * cv::boxFilter on CV_32F accumulates in double; an fp64 sum of 64 floats is exact unless the
//     window spans > 2^23 in magnitude, so any summation order gives the same float after the one
//     final rounding.  All eight box sums here are fp64: running column sums (+ newest row,
//     - oldest row) and 8-wide row sums, scaled by 1/64 in fp64, rounded once to float.
//   * every fp32 operation is a separate IEEE round-to-nearest op in the reference's order
//     (CVF.cpp:87-163).  Multiplies of two columns share one FMUL2; adds stay scalar because
//     ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with --fmad=false.
//
// Work decomposition (B200-first; nothing like the reference's per-slice Mat pipeline):
//   warp   = one strip of 128 input columns (112 output columns) of ONE disparity slice, one row segment
//   lane   = 4 consecutive columns -> every global access is a 128-bit load/store
//   CTA    = 3 warps = 3 consecutive slices of the same strip and segment (guide rows hit in L1)
//   stage 1: S1[4 boxes][4 cols] fp64 running column sums of p, I0*p, I1*p, I2*p;
//            row sums = per-lane prefix/suffix sums + 4 fp64 shuffles per box (lane+1 total, lane+2 prefixes)
//   a,b    : CVF.cpp:92-155 with the d-independent adjugate / 1/det precomputed per pixel (K2)
//   ring   : thread-private 8-row history of a0,a1,a2,b in shared memory (512 B per thread);
//            the only on-chip history: stage-1 "oldest rows" are re-read from the raw volume (L1/L2)
//   stage 2: S2[4 planes][4 cols] fp64 running column sums of a0,a1,a2,b (newest from registers,
//            oldest from the ring); row sums as in stage 1; q = box(b) + sum_c box(a_c) * I_c
//
// Column bookkeeping of strip s (X0 = first output column, 112*s; the last strip is shifted left
// to end at the image edge), lane l:
//   input  columns X0-8+4l .. +3   (p, I)                  halo columns come from the mirrored halo
//   a,b    columns X0-4+4l .. +3   (valid for l <= 29)     of the padded layout (psm_kernels.cuh)
//   output columns X0  +4l .. +3   (valid for l <= 27)
// a,b at columns outside [0,W) must be the REFLECTED a,b (cv::boxFilter reflects its input, which
// for the second stage is a,b) -- not what stage 1 computes there from mirrored p (the 8-wide window
// is asymmetric) -- so border strips overwrite those entries by lane shuffles.
// Rows: the top of the image uses weights 1,2,2,2,1 for the first window, the bottom feeds three
// virtual rows from the ring; input rows reflect by index.  These special cases live in
// generic_step; the bulk of the rows run steady_step, which has no conditionals at all.
#pragma once
#include "psm_kernels.cuh"

namespace psm {

constexpr int kStripOut = 112;  // output columns per warp
constexpr int kStripIn = 128;   // input columns per warp
constexpr int kCvfThreads = 96;       // shipped CTA size: 3 slice-warps, 48 KB ring, 4 CTAs/SM (measured best of 32/64/96/128)
constexpr int kCvfMaxThreads = 128;   // upper bound the register allocation is sized for

struct CvfParams {
    const float* vol_in[2];   // raw volumes  [Dloc][H][Wp]   (pointer to row 0, column 0)
    float* vol_out[2];        // filtered volumes
    const float* guide[2];    // guide planes [kGuidePlanes][H][Wp]
    int W, H, Wp, Dloc;
    int nstrips, nseg, seg_rows, ndgroups;
    int remap_sms, remap_ctas;  // SM count and resident CTAs per SM for the block->work remap (0: identity)
};

struct f2x2 { float2 lo, hi; };  // four columns as two packed pairs
__device__ __forceinline__ f2x2 from4(const float4& v) { return {make_float2(v.x, v.y), make_float2(v.z, v.w)}; }
__device__ __forceinline__ float4 to4(const f2x2& v) { return make_float4(v.lo.x, v.lo.y, v.hi.x, v.hi.y); }
__device__ __forceinline__ f2x2 mul2(const f2x2& a, const f2x2& b) { return {__fmul2_rn(a.lo, b.lo), __fmul2_rn(a.hi, b.hi)}; }
__device__ __forceinline__ f2x2 add2(const f2x2& a, const f2x2& b)
{
    return {make_float2(fadd(a.lo.x, b.lo.x), fadd(a.lo.y, b.lo.y)), make_float2(fadd(a.hi.x, b.hi.x), fadd(a.hi.y, b.hi.y))};
}
__device__ __forceinline__ f2x2 sub2(const f2x2& a, const f2x2& b)
{
    return {make_float2(fsub(a.lo.x, b.lo.x), fsub(a.lo.y, b.lo.y)), make_float2(fsub(a.hi.x, b.hi.x), fsub(a.hi.y, b.hi.y))};
}
__device__ __forceinline__ float get(const f2x2& v, int j) { return j == 0 ? v.lo.x : (j == 1 ? v.lo.y : (j == 2 ? v.hi.x : v.hi.y)); }

// f32 -> f64 without the conversion pipe: for a positive normal float the double is
// {hi = (u >> 3) + 0x38000000, lo = u << 29}; everything else (zero, denormal, negative, inf, nan)
// takes the F2F instruction under a predicate that is almost never set.  Stage-1 inputs (p and
// I*p) are non-negative, so on real data the XU pipe is spared these conversions.
__device__ __forceinline__ double widen_pos(float f)
{
    const unsigned u = __float_as_uint(f);
    double d = __hiloint2double((int)((u >> 3) + 0x38000000u), (int)(u << 29));
    if (u - 0x00800000u >= 0x7f000000u) d = (double)f;
    return d;
}

__device__ __forceinline__ float mean64(double s) { return (float)__dmul_rn(s, 1.0 / 64.0); }

__device__ __forceinline__ float4 ldg4(const char* __restrict__ p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// MINB: resident CTAs per SM the register allocation is sized for (3 -> <=168 regs, 2 -> <=255);
// IW: integer widening (widen_pos) of the stage-1 inputs: 0 none, 1 oldest rows, 2 newest + oldest rows.
template <int MINB, int IW>
__global__ void __launch_bounds__(kCvfMaxThreads, MINB)
cvf_stream_kernel(const CvfParams P)
{
    extern __shared__ float4 ring[];  // [8 slots][4 planes][blockDim.x threads]
    constexpr bool IWN = IW >= 2, IWO = IW >= 1;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int nthr = blockDim.x;            // kCvfThreads in the shipped configuration (option 103 varies it)
    const int wpc = nthr >> 5;

    // Block -> work mapping.  Hardware hands consecutive blockIdx to consecutive SMs, so the CTAs that
    // share an SM (and its L1) are blockIdx k, k+nsm, k+2nsm, ...; the remap makes those neighbours in
    // work order (consecutive slice groups of the same strip and segment), which read the same guide rows.
    int b = blockIdx.x;
    if (P.remap_sms > 0) {
        const int per = P.remap_sms * P.remap_ctas;
        const int chunk = b / per, r = b - chunk * per;
        if ((chunk + 1) * per <= (int)gridDim.x) b = chunk * per + (r % P.remap_sms) * P.remap_ctas + r / P.remap_sms;
    }
    const int dgroup = b % P.ndgroups; b /= P.ndgroups;
    const int strip = b % P.nstrips;   b /= P.nstrips;
    const int seg = b % P.nseg;
    const int view = b / P.nseg;
    const bool slice_ok = dgroup * wpc + warp < P.Dloc;
    const int dlc = slice_ok ? dgroup * wpc + warp : P.Dloc - 1;  // surplus warps redo the last slice, stores masked
    if (!slice_ok) return;  // warps never synchronise with each other

    const int W = P.W, H = P.H;
    const unsigned Wp = (unsigned)P.Wp;
    const unsigned plane = (unsigned)H * Wp;
    const int out_lo = strip * kStripOut;
    const int X0 = (strip == P.nstrips - 1 && strip > 0) ? ((W - kStripOut + 3) & ~3) : out_lo;
    const int cin = X0 - 8 + 4 * lane;
    // per-thread base pointers (bytes); every offset added to them below is warp-uniform
    const char* __restrict__ Gi = reinterpret_cast<const char*>(P.guide[view] + cin);      // guide, input columns
    const char* __restrict__ Ga = reinterpret_cast<const char*>(P.guide[view] + cin + 4);  // guide, a,b columns
    const char* __restrict__ Go = reinterpret_cast<const char*>(P.guide[view] + cin + 8);  // guide, output columns
    const char* __restrict__ vin = reinterpret_cast<const char*>(P.vol_in[view] + (size_t)dlc * plane + cin);
    char* __restrict__ vout = reinterpret_cast<char*>(P.vol_out[view] + (size_t)dlc * plane + cin + 8);
    const size_t planeB = (size_t)plane * 4, rowB = (size_t)Wp * 4;
    const bool store_ok = slice_ok && lane <= 27 && cin + 8 < W && cin + 8 >= out_lo;

    // ---- x-reflection plan for a,b (strips whose a,b columns X0-4 .. X0+115 leave the image) ----
    const bool fix_left = X0 == 0;
    const bool fix_right = X0 + 115 >= W;
    int fix_src[4];               // source lane of element j (this lane's a,b column X0-4+4l+j mirrored into the image)
    unsigned maskL = 0, maskR = 0;  // bit j: element j lies left of column 0 / right of column W-1
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int xa = cin + 4 + j;
        const int r = reflect101(xa < -4 ? -4 : (xa > W + 2 ? W + 2 : xa), W) - (X0 - 4);
        fix_src[j] = (r >> 2) & 31;
        maskL |= (fix_left && xa < 0) ? (1u << j) : 0u;
        maskR |= (fix_right && xa >= W && xa <= W + 2) ? (1u << j) : 0u;
    }
    // which element of the source lane: left (8-j)&3, right (2(W-1)-j)&3 -- uniform per side, so the
    // source lane (which does not know who reads it) can put the right element on the wire
    const int eR = (2 * (W - 1)) & 3;

    const int Y0 = seg * P.seg_rows;
    const int Y1 = min(H, Y0 + P.seg_rows);
    const bool top = (Y0 == 0);
    const bool bottom = (Y1 == H);
    const int T0 = top ? 0 : Y0 - 4;              // first a,b row
    const int Tlast = bottom ? H - 1 : Y1 + 2;    // last real a,b row
    const int Tend = bottom ? H + 2 : Tlast;      // last step (three virtual rows below the image)

    double S1[4][4], S2[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j) { S1[q][j] = 0.0; S2[q][j] = 0.0; }

    struct RowIn { float4 p, i0, i1, i2; };
    auto load_at = [&](size_t ro) {  // ro: byte offset of the row
        RowIn x;
        x.p = ldg4(vin + ro);
        x.i0 = ldg4(Gi + ro);
        x.i1 = ldg4(Gi + (planeB + ro));
        x.i2 = ldg4(Gi + (2 * planeB + ro));
        return x;
    };
    auto load_row = [&](int r) { return load_at((size_t)reflect101(r, H) * rowB); };
    auto add_row = [&](const RowIn& x) {
        const f2x2 p = from4(x.p);
        const f2x2 m0 = mul2(from4(x.i0), p), m1 = mul2(from4(x.i1), p), m2 = mul2(from4(x.i2), p);  // CVF.cpp:87
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            S1[0][j] = __dadd_rn(S1[0][j], IWN ? widen_pos(get(p, j)) : (double)get(p, j));
            S1[1][j] = __dadd_rn(S1[1][j], IWN ? widen_pos(get(m0, j)) : (double)get(m0, j));
            S1[2][j] = __dadd_rn(S1[2][j], IWN ? widen_pos(get(m1, j)) : (double)get(m1, j));
            S1[3][j] = __dadd_rn(S1[3][j], IWN ? widen_pos(get(m2, j)) : (double)get(m2, j));
        }
    };
    auto sub_row = [&](const RowIn& x) {
        const f2x2 p = from4(x.p);
        const f2x2 m0 = mul2(from4(x.i0), p), m1 = mul2(from4(x.i1), p), m2 = mul2(from4(x.i2), p);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            S1[0][j] = __dsub_rn(S1[0][j], IWO ? widen_pos(get(p, j)) : (double)get(p, j));
            S1[1][j] = __dsub_rn(S1[1][j], IWO ? widen_pos(get(m0, j)) : (double)get(m0, j));
            S1[2][j] = __dsub_rn(S1[2][j], IWO ? widen_pos(get(m1, j)) : (double)get(m1, j));
            S1[3][j] = __dsub_rn(S1[3][j], IWO ? widen_pos(get(m2, j)) : (double)get(m2, j));
        }
    };
    auto load_guide = [&](size_t ro, float4 (&g4)[10]) {
#pragma unroll
        for (int q = 0; q < 10; ++q) g4[q] = ldg4(Ga + ((size_t)(kGuideMean + q) * planeB + ro));
    };

    // stage-1 row sums -> means -> cov -> a,b (CVF.cpp:81-155), then the x-reflection of a,b
    auto coeffs = [&](const float4 (&g4)[10], f2x2 (&av)[4]) {
        double h[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) hsum8(S1[q], h[q]);
        f2x2 m[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            m[q] = {make_float2(mean64(h[q][0]), mean64(h[q][1])), make_float2(mean64(h[q][2]), mean64(h[q][3]))};
        const f2x2 mI0 = from4(g4[0]), mI1 = from4(g4[1]), mI2 = from4(g4[2]);
        const f2x2 M00 = from4(g4[3]), M01 = from4(g4[4]), M02 = from4(g4[5]);
        const f2x2 M11 = from4(g4[6]), M12 = from4(g4[7]), M22 = from4(g4[8]);
        const f2x2 idet = from4(g4[9]);
        const f2x2 c0 = sub2(m[1], mul2(mI0, m[0]));   // CVF.cpp:92-95
        const f2x2 c1 = sub2(m[2], mul2(mI1, m[0]));
        const f2x2 c2 = sub2(m[3], mul2(mI2, m[0]));
        av[0] = mul2(idet, add2(add2(mul2(c0, M00), mul2(c1, M01)), mul2(c2, M02)));  // CVF.cpp:121-146
        av[1] = mul2(idet, add2(add2(mul2(c0, M01), mul2(c1, M11)), mul2(c2, M12)));
        av[2] = mul2(idet, add2(add2(mul2(c0, M02), mul2(c1, M12)), mul2(c2, M22)));
        av[3] = sub2(sub2(sub2(m[0], mul2(av[0], mI0)), mul2(av[1], mI1)), mul2(av[2], mI2));  // CVF.cpp:152-155
        if (fix_left | fix_right) {  // warp-uniform
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float e[4] = {av[q].lo.x, av[q].lo.y, av[q].hi.x, av[q].hi.y};
                float r[4] = {e[0], e[1], e[2], e[3]};
                if (fix_left) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float v = __shfl_sync(0xffffffffu, e[(8 - j) & 3], fix_src[j]);
                        if ((maskL >> j) & 1u) r[j] = v;
                    }
                }
                if (fix_right) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int er = (eR - j) & 3;
                        const float sv = er == 0 ? e[0] : (er == 1 ? e[1] : (er == 2 ? e[2] : e[3]));
                        const float v = __shfl_sync(0xffffffffu, sv, fix_src[j]);
                        if ((maskR >> j) & 1u) r[j] = v;
                    }
                }
                av[q] = {make_float2(r[0], r[1]), make_float2(r[2], r[3])};
            }
        }
    };

    // stage-2 row sums -> q for one output row; i* are the guide channels at the output columns
    auto emit = [&](size_t ro, const float4& i0, const float4& i1, const float4& i2) {
        double h2[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) hsum8(S2[q], h2[q]);
        f2x2 mb[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            mb[q] = {make_float2(mean64(h2[q][0]), mean64(h2[q][1])), make_float2(mean64(h2[q][2]), mean64(h2[q][3]))};
        // q = box(b) + box(a0)*I0 + box(a1)*I1 + box(a2)*I2, accumulated in that order (CVF.cpp:157-163)
        f2x2 qv = add2(mb[3], mul2(mb[0], from4(i0)));
        qv = add2(qv, mul2(mb[1], from4(i1)));
        qv = add2(qv, mul2(mb[2], from4(i2)));
        if (store_ok) *reinterpret_cast<float4*>(vout + ro) = to4(qv);
    };

    // ---- generic step: any row, every special case (top weights, warm-up, reflection, virtual rows)
    auto generic_step = [&](int t) {
        const bool real_row = t <= Tlast;
        f2x2 av[4];
        if (real_row) {
            const RowIn xn = load_row(t + 3);
            const RowIn xo = load_row(t - 4);
            float4 g4[10];
            load_guide((size_t)t * rowB, g4);
            add_row(xn);
            coeffs(g4, av);
            sub_row(xo);
        } else {  // virtual a,b row below the image == reflected row, still in the ring
            const int slot = reflect101(t, H) & 7;
#pragma unroll
            for (int q = 0; q < 4; ++q) av[q] = from4(ring[(slot * 4 + q) * nthr + tid]);
        }
        const int age = t - T0;
        const bool warm = top ? (t <= 4) : (age < 8);
        const double wnew = (top && t >= 1 && t <= 3) ? 2.0 : 1.0;      // rows 1..3 appear twice in row 0's window
        const int oslot = (top && t < 8) ? ((8 - t) & 7) : (t & 7);      // slot of a,b row reflect(t-8)
        const int nslot = t & 7;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 old4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!warm) old4 = ring[(oslot * 4 + q) * nthr + tid];
            if (real_row) ring[(nslot * 4 + q) * nthr + tid] = to4(av[q]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double s = __fma_rn(wnew, (double)get(av[q], j), S2[q][j]);  // exact: wnew is 1 or 2
                if (!warm) s = __dsub_rn(s, (double)comp(old4, j));
                S2[q][j] = s;
            }
        }
        const bool first_out = top ? (t == 4) : (age == 7);
        if (warm && !first_out) return;
        const int nrows = (top && t == 4) ? 2 : 1;  // output rows 0 and 1 share one reflected window
        for (int e = 0; e < nrows; ++e) {
            const int y = (top && t == 4) ? e : t - 3;
            if (y < Y0 || y >= Y1) continue;
            const size_t ro = (size_t)y * rowB;
            emit(ro, ldg4(Go + ro), ldg4(Go + (planeB + ro)), ldg4(Go + (2 * planeB + ro)));
        }
    };

    // ---- schedule ----------------------------------------------------------------------------
    // steady steps need: a real row, no warm-up, no reflected input row (t-4 >= 0, t+4 <= H-1 because
    // the step also preloads row t+4 for its successor) and the plain ring slot t&7
    const int Ts0 = top ? 8 : T0 + 8;
    const int Ts1 = min(Tlast, H - 5);

    for (int r = T0 - 4; r <= T0 + 2; ++r) add_row(load_row(r));
    int t = T0;
    for (; t <= Tend && t < Ts0; ++t) generic_step(t);

    if (t <= Ts1) {
        size_t ro_n = (size_t)(t + 3) * rowB;  // newest input row  t+3   (byte offsets)
        size_t ro_o = (size_t)(t - 4) * rowB;  // oldest input row  t-4
        size_t ro_t = (size_t)t * rowB;        // a,b row           t
        size_t ro_y = (size_t)(t - 3) * rowB;  // output row        t-3
        RowIn xn = load_at(ro_n);   // newest row of the next stage-1 step (loaded one step ahead)
        RowIn xo = load_at(ro_o);   // oldest row of the next stage-1 step (loaded one step ahead)
        // stage 1 of a,b row ro_t: S1 += newest, row sums -> a,b, S1 -= oldest; refills xn / xo
        auto stage1 = [&](f2x2 (&av)[4]) {
            float4 g4[10];
            load_guide(ro_t, g4);
            add_row(xn);
            ro_n += rowB;
            xn = load_at(ro_n);     // into the registers add_row just released
            coeffs(g4, av);
            sub_row(xo);
            ro_o += rowB;
            xo = load_at(ro_o);
            ro_t += rowB;
        };
        // stage 2 of a,b row tt (output row tt-3 at byte offset ro_y): ring exchange, S2 update, q
        auto stage2 = [&](int tt, const f2x2 (&av)[4], const float4& o0, const float4& o1, const float4& o2) {
            float4* rp = ring + ((tt & 7) * 4) * nthr + tid;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 old4 = rp[q * nthr];
                rp[q * nthr] = to4(av[q]);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    S2[q][j] = __dsub_rn(__dadd_rn(S2[q][j], (double)get(av[q], j)), (double)comp(old4, j));
            }
            emit(ro_y, o0, o1, o2);
            ro_y += rowB;
        };
        for (; t <= Ts1; ++t) {
            const float4 o0 = ldg4(Go + ro_y), o1 = ldg4(Go + (planeB + ro_y)), o2 = ldg4(Go + (2 * planeB + ro_y));
            f2x2 av[4];
            stage1(av);
            stage2(t, av, o0, o1, o2);
        }
    }
    for (; t <= Tend; ++t) generic_step(t);
}

}  // namespace psm
