// psm_cvf_stream3.cuh -- K3 v3: fused streaming guided-image-filter kernel (the graded kernel).
//
// q = GuidedFilter_cv(I, p) (/root/reference/src/CVF.cpp:72-165) for every owned slice of both cost
// volumes in ONE pass over HBM: p is read once, q written once; the two chained 8x8 box stages,
// a, b, all products and means stay on the SM.  Algorithm and numerics are those of v1/v2
// (psm_cvf_stream.cuh has the derivation): all eight box sums are fp64 (== cv::boxFilter's double
// accumulation, order-independent in practice), all fp32 math is unfused IEEE RN, q is bit-exact.
//
// Work decomposition:
//   warp   = one 128-column strip (112 output columns) of one disparity slice, one row segment
//   lane   = 4 consecutive columns (128-bit loads/stores)
//   CTA    = 4 warps = 4 consecutive slices of the same strip/segment (guide rows hit in L1)
//   stage 1: fp64 running column sums of p, I0*p, I1*p, I2*p  (+ newest row, - oldest row)
//            8-wide row sums = lane prefix/suffix sums + 4 fp64 shuffles per box
//   a,b    : fp32, reference operation order; FMUL2 packs the multiplies of two columns
//   ring   : thread-private 8-row history of a0,a1,a2,b in shared memory (512 B / thread)
//   stage 2: fp64 running column sums of a0,a1,a2,b fed from registers (newest) and ring (oldest)
//
// v3 = v2 + (ncu on v2: 775 instructions per warp-row of which ~190 were border/warm-up/reflect
// bookkeeping; 35% of stall samples on the first use of the row loads):
//   * the row loop is split into generic steps (segment warm-up, image top/bottom, virtual rows)
//     and STEADY steps with no conditionals, no reflection and running row offsets;
//   * the newest input row of step t+1 is loaded during step t (software pipelining), all other
//     loads of a step are issued at its top.
#pragma once
#include "psm_cvf_stream2.cuh"

namespace psm {

template <bool BORDER>
__device__ __forceinline__ void cvf3_body(const CvfParams& P, float4* ring, int view, int seg, int strip, int dlc)
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
    const int cin = X0 - 8 + 4 * lane;  // input columns
    const int ca = cin + 4;             // a,b columns
    const int co = cin + 8;             // output columns
    const bool store_ok = lane <= 27 && co < W && co >= out_lo;

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
            const int sl = r >> 2;
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
    const int T0 = top ? 0 : Y0 - 4;              // first a,b row
    const int Tlast = bottom ? H - 1 : Y1 + 2;    // last real a,b row
    const int Tend = bottom ? H + 2 : Tlast;      // last step (three virtual rows below the image)

    double S1[4][4], S2[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j) { S1[q][j] = 0.0; S2[q][j] = 0.0; }

    struct RowIn { float4 p, i0, i1, i2; };
    auto load_at = [&](unsigned ro) {
        RowIn x;
        x.p = ld4<BORDER>(vin, ro, cin, cp.in);
        x.i0 = ld4<BORDER>(G, ro, cin, cp.in);
        x.i1 = ld4<BORDER>(G, plane + ro, cin, cp.in);
        x.i2 = ld4<BORDER>(G, 2 * plane + ro, cin, cp.in);
        return x;
    };
    auto load_row = [&](int r) { return load_at((unsigned)reflect101(r, H) * (unsigned)Wp); };
    auto add_row = [&](const RowIn& x) {
        const f2x2 p = from4(x.p);
        const f2x2 m0 = mul2(from4(x.i0), p), m1 = mul2(from4(x.i1), p), m2 = mul2(from4(x.i2), p);  // CVF.cpp:87
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            S1[0][j] = __dadd_rn(S1[0][j], (double)get(p, j));
            S1[1][j] = __dadd_rn(S1[1][j], (double)get(m0, j));
            S1[2][j] = __dadd_rn(S1[2][j], (double)get(m1, j));
            S1[3][j] = __dadd_rn(S1[3][j], (double)get(m2, j));
        }
    };
    auto sub_row = [&](const RowIn& x) {
        const f2x2 p = from4(x.p);
        const f2x2 m0 = mul2(from4(x.i0), p), m1 = mul2(from4(x.i1), p), m2 = mul2(from4(x.i2), p);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            S1[0][j] = __dsub_rn(S1[0][j], (double)get(p, j));
            S1[1][j] = __dsub_rn(S1[1][j], (double)get(m0, j));
            S1[2][j] = __dsub_rn(S1[2][j], (double)get(m1, j));
            S1[3][j] = __dsub_rn(S1[3][j], (double)get(m2, j));
        }
    };
    auto load_guide = [&](unsigned ro, float4 (&g4)[10]) {
#pragma unroll
        for (int q = 0; q < 10; ++q) g4[q] = ld4<BORDER>(G, (unsigned)(kGuideMean + q) * plane + ro, ca, cp.ab);
    };

    // stage-1 row sums -> means -> cov -> a,b  (CVF.cpp:81-155), then the x-reflection of a,b
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
        const f2x2 c0 = sub2(m[1], mul2(mI0, m[0]));
        const f2x2 c1 = sub2(m[2], mul2(mI1, m[0]));
        const f2x2 c2 = sub2(m[3], mul2(mI2, m[0]));
        av[0] = mul2(idet, add2(add2(mul2(c0, M00), mul2(c1, M01)), mul2(c2, M02)));
        av[1] = mul2(idet, add2(add2(mul2(c0, M01), mul2(c1, M11)), mul2(c2, M12)));
        av[2] = mul2(idet, add2(add2(mul2(c0, M02), mul2(c1, M12)), mul2(c2, M22)));
        av[3] = sub2(sub2(sub2(m[0], mul2(av[0], mI0)), mul2(av[1], mI1)), mul2(av[2], mI2));
        if (BORDER) {
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
    };

    // stage-2 row sums -> q for one output row; i* are the guide channels at the output columns
    auto emit = [&](unsigned ro, const float4& i0, const float4& i1, const float4& i2) {
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
        if (store_ok) *reinterpret_cast<float4*>(vout + ro + co) = to4(qv);
    };

    // ---- generic step: any row, any special case (top weights, warm-up, reflection, virtual rows)
    auto generic_step = [&](int t) {
        const bool real_row = t <= Tlast;
        f2x2 av[4];
        if (real_row) {
            const RowIn xn = load_row(t + 3);
            const RowIn xo = load_row(t - 4);
            float4 g4[10];
            load_guide((unsigned)t * (unsigned)Wp, g4);
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
            const unsigned ro = (unsigned)y * (unsigned)Wp;
            emit(ro, ld4<BORDER>(G, ro, co, cp.out), ld4<BORDER>(G, plane + ro, co, cp.out),
                 ld4<BORDER>(G, 2 * plane + ro, co, cp.out));
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
        unsigned ro_n = (unsigned)(t + 3) * (unsigned)Wp;  // newest input row  t+3
        unsigned ro_o = (unsigned)(t - 4) * (unsigned)Wp;  // oldest input row  t-4
        unsigned ro_t = (unsigned)t * (unsigned)Wp;        // a,b row           t
        unsigned ro_y = (unsigned)(t - 3) * (unsigned)Wp;  // output row        t-3
        RowIn xn = load_at(ro_n);
        for (; t <= Ts1; ++t) {
            // all loads of this step, and the newest row of the next one
            const RowIn xo = load_at(ro_o);
            float4 g4[10];
            load_guide(ro_t, g4);
            const float4 o0 = ld4<BORDER>(G, ro_y, co, cp.out);
            const float4 o1 = ld4<BORDER>(G, plane + ro_y, co, cp.out);
            const float4 o2 = ld4<BORDER>(G, 2 * plane + ro_y, co, cp.out);
            ro_n += Wp;
            const RowIn xnext = load_at(ro_n);

            add_row(xn);
            f2x2 av[4];
            coeffs(g4, av);
            sub_row(xo);

            float4* rp = ring + ((t & 7) * 4) * nthr + tid;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 old4 = rp[q * nthr];
                rp[q * nthr] = to4(av[q]);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    S2[q][j] = __dsub_rn(__dadd_rn(S2[q][j], (double)get(av[q], j)), (double)comp(old4, j));
            }
            emit(ro_y, o0, o1, o2);

            xn = xnext;
            ro_o += Wp; ro_t += Wp; ro_y += Wp;
        }
    }
    for (; t <= Tend; ++t) generic_step(t);
}

template <bool BORDER>
__global__ void __launch_bounds__(128, BORDER ? 2 : 3)
cvf_stream3_kernel(const CvfParams P, const int* __restrict__ strip_list, int nlist)
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
    cvf3_body<BORDER>(P, ring, view, seg, strip, d);
}

}  // namespace psm
