// psm_kernels.cuh -- ingest, cost-volume construction (CVC), guide precompute, WTA and the
// unfused "naive" CVF cross-check kernels.  The fused streaming CVF kernel is in
// psm_cvf_stream.cuh.  Layout: every plane / volume row has pitch Wp floats and is stored with a
// column halo: kPadLeft columns before column 0 and at least kPadRight after column W-1 (so a
// pointer to (row, col 0) may be indexed from -kPadLeft .. W+kPadRight-1); pitch and halo are
// multiples of 4 floats, so 4-column groups are 16-byte aligned.  The halos of the cost volumes
// and of the three guide channels hold the BORDER_REFLECT_101 mirror of the row (written by
// pad_cols_kernel), which lets the streaming CVF kernel use plain 128-bit loads everywhere.
// Planes are [H][Wp], volumes [d_local][H][Wp]; "W4" below is W rounded up to 4.
#pragma once
#include "psm_common.cuh"

namespace psm {

constexpr int kPadLeft = 32;   // floats before column 0 (128 B: column 0 of every row is cache-line aligned);
                               // the last 8 of them are the mirrored left halo
constexpr int kPadRight = 12;  // minimum halo columns after column W4-1 (the last strip reads up to W+10)

// Row pitch (floats) for an image of width W: left pad + max(W4 + right halo, one full strip),
// rounded up to a multiple of 32 floats so that every row starts on a 128-byte line.
__host__ __device__ inline int pitch_for_width(int W)
{
    const int w4 = (W + 3) & ~3;
    const int right = (w4 + kPadRight > 120) ? w4 + kPadRight : 120;
    return kPadLeft + ((right + 31) & ~31);
}

// Fill the column halos of `nrows` rows with the BORDER_REFLECT_101 mirror of the row:
// row[-k] = row[k] (k = 1..8), row[W+k] = row[W-2-k] (k = 0..7), zeros beyond.
__global__ void pad_cols_kernel(float* __restrict__ base, size_t nrows, int W, int Wp)
{
    const size_t row = (size_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    float* r = base + row * (size_t)Wp;
    const int lane = threadIdx.x & 31;
    const int right = Wp - kPadLeft - W;  // halo columns after W-1
    for (int k = lane; k < kPadLeft + right; k += 32) {
        const int x = k < kPadLeft ? -(k + 1) : W + (k - kPadLeft);
        const bool mirrored = (x < 0 && x >= -8) || (x >= W && x < W + 8);
        r[x] = mirrored ? r[reflect101(x, W)] : 0.f;
    }
}

// ------------------------------------------------------------------------------------------
// Ingest: interleaved BGR (float or u8) -> 3 planar channels + x-gradient of the gray image.
// Restates StereoMatch.cpp:193-197 (u8 * (1/255.0f)), CVF.cpp:47 (split) and CVC.cpp:41-46
// (cvtColor RGB2GRAY on the BGR image, Sobel dx=1 ksize=1 with BORDER_REFLECT_101).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(uint8_t v) { return fmul((float)v, 1 / 255.0f); }

template <typename T>
__device__ __forceinline__ float gray_at(const T* __restrict__ row, int x, int gray_mode)
{
    const float c0 = to_f32(row[3 * x]), c1 = to_f32(row[3 * x + 1]), c2 = to_f32(row[3 * x + 2]);
    if (gray_mode == 0) return __fmaf_rn(c2, 0.114f, __fmaf_rn(c0, 0.299f, fmul(c1, 0.587f)));
    return fadd(fadd(fmul(c0, 0.299f), fmul(c1, 0.587f)), fmul(c2, 0.114f));
}

template <typename T>
__global__ void ingest_kernel(const T* __restrict__ src, size_t step_bytes, int W, int H, int Wp,
                              float* __restrict__ I0, float* __restrict__ I1, float* __restrict__ I2,
                              float* __restrict__ grd, int gray_mode, int* __restrict__ guide_flag)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= ((W + 3) & ~3) || y >= H) return;
    const size_t o = (size_t)y * Wp + x;
    if (x >= W) { I0[o] = 0.f; I1[o] = 0.f; I2[o] = 0.f; grd[o] = 0.f; return; }
    const T* row = reinterpret_cast<const T*>(reinterpret_cast<const char*>(src) + (size_t)y * step_bytes);
    const float v0 = to_f32(row[3 * x]), v1 = to_f32(row[3 * x + 1]), v2 = to_f32(row[3 * x + 2]);
    I0[o] = v0;
    I1[o] = v1;
    I2[o] = v2;
    // the integer widening in guide_kernel / the CVF kernel assumes I >= +0, finite, and I * I finite (images are in
    // [0,1] by contract, DispEst data contract SURVEY 8b); a negative, non-finite or >= 2^63 value raises the per-view
    // flag and selects their conversion-pipe paths (bit patterns compare like the magnitudes; the sign bit sorts last)
    if (max(max(__float_as_uint(v0), __float_as_uint(v1)), __float_as_uint(v2)) >= 0x5f000000u) atomicOr(guide_flag, 1);
    const float gr = gray_at(row, reflect101(x + 1, W), gray_mode);
    const float gl = gray_at(row, reflect101(x - 1, W), gray_mode);
    grd[o] = fsub(gr, gl);
}

// ------------------------------------------------------------------------------------------
// K1 CVC.  One thread = 4 consecutive pixels of one row of ONE view; it walks the owned
// disparities keeping the matched window of the other image in registers (one new scalar load
// per plane per disparity; GROUPED builds take one 128-bit load per plane per FOUR disparities when
// the shard starts at a multiple of 4 -- same speed, more registers, kept for A/B) and writes one
// 128-bit store per slice.
//   left  volume (SIGN=-1): x >= d     ? cost4(L[x], R[x-d]) : border(L[x])   CVC.cpp:122-149
//   right volume (SIGN=+1): x <  W - d ? cost4(R[x], L[x+d]) : border(R[x])   CVC.cpp:151-179
// cost4 = CVC.cpp:18-27 (float), border = CVC.cpp:30-39 (double intermediates, BC_32F = 1.0).
// ------------------------------------------------------------------------------------------
struct CvcParams {
    const float* self[4];   // planes of the view being built: c0,c1,c2,grd
    const float* other[4];  // planes of the matched view
    float* vol;             // [d_count][H][Wp]
    int W, H, Wp, d_begin, d_count;
    int fold_halo;          // 1: this kernel also writes the mirrored column halo (W >= 32)
};

__device__ __forceinline__ float cost4(float l0, float l1, float l2, float lg,
                                       float r0, float r1, float r2, float rg)
{
    const float clr = fadd(fadd(fabsf(fsub(l0, r0)), fabsf(fsub(l1, r1))), fabsf(fsub(l2, r2)));
    const float grd = fabsf(fsub(lg, rg));
    return fadd(fmul(0.9f, clr), fmul(fsub(1.0f, 0.9f), grd));
}

__device__ __forceinline__ float cost_border(float l0, float l1, float l2, float lg)
{
    const double s = __dadd_rn(__dadd_rn(fabs(__dsub_rn((double)l0, 1.0)), fabs(__dsub_rn((double)l1, 1.0))),
                               fabs(__dsub_rn((double)l2, 1.0)));
    const float clr = (float)s;
    const float grd = (float)fabs(__dsub_rn((double)lg, 1.0));
    return fadd(fmul(0.9f, clr), fmul(fsub(1.0f, 0.9f), grd));
}

template <int SIGN, int GROUPED>
__device__ __forceinline__ void cvc_body(const CvcParams& P);

// both volumes in ONE launch (blockIdx.z = view): twice the CTAs in flight and no tail between the two views.
// MINB = resident CTAs per SM the register budget is cut for; GROUPED: 0 = shipped (interior fast path + scalar-load window for the
// border warps), 1 = border / all warps use grouped 128-bit window loads when the shard is aligned (A/B), 2 = no fast path (round-1 loop).
// blockIdx.z = chunk * 2 + view: with chunk > 0 (P2.chunk slices per CTA instead of all owned slices) the grid walks the volume
// chunk by chunk -- z is the slowest block index -- so that at any moment the GPU writes a few neighbouring slices instead of one
// 2 KB run in each of 128 slices 8 MB apart (DRAM page locality of the 2.1 GB store stream).
struct CvcParams2 { CvcParams v[2]; int chunk; };
template <int MINB, int GROUPED>
__global__ void __launch_bounds__(128, MINB) cvc_both_kernel(const CvcParams2 P2)
{
    const int view = blockIdx.z & 1, chunk = blockIdx.z >> 1;
    CvcParams P = P2.v[view];
    if (P2.chunk > 0) {
        const int first = chunk * P2.chunk;
        P.vol += (size_t)first * P.H * P.Wp;
        P.d_begin += first;
        P.d_count = min(P2.chunk, P.d_count - first);
    }
    if (view == 0) cvc_body<-1, GROUPED>(P);
    else cvc_body<+1, GROUPED>(P);
}

template <int SIGN, int GROUPED>
__device__ __forceinline__ void cvc_body(const CvcParams& P)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    if (x4 >= ((P.W + 3) & ~3)) return;
    const int W = P.W;
    const size_t ro = (size_t)y * P.Wp;

    float s[4][4], bord[4];  // [plane][pixel]
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(P.self[c] + ro + x4));
        s[c][0] = v.x; s[c][1] = v.y; s[c][2] = v.z; s[c][3] = v.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) bord[j] = cost_border(s[0][j], s[1][j], s[2][j], s[3][j]);

    const float* o0 = P.other[0] + ro;
    const float* o1 = P.other[1] + ro;
    const float* o2 = P.other[2] + ro;
    const float* o3 = P.other[3] + ro;
    auto clampx = [W](int x) { return x < 0 ? 0 : (x >= W ? W - 1 : x); };

    int d = P.d_begin;
    const size_t slice = (size_t)P.H * P.Wp;
    float* out = P.vol + ro + x4;
    // mirrored column halo (see pad_cols_kernel), written here for W >= 32: pixel x in [1,8] also goes to
    // column -x, pixel x in [W-9, W-2] also to column 2(W-1)-x.  Only the edge threads of a row take part.
    int halo_off[4];
    bool edge = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x = x4 + j;
        halo_off[j] = 0;
        if (P.fold_halo && x < W) {
            if (x >= 1 && x <= 8) halo_off[j] = -x - x4;
            else if (x >= W - 9 && x <= W - 2) halo_off[j] = 2 * (W - 1) - x - x4;
        }
        edge |= halo_off[j] != 0;
    }
    const bool full_group = x4 + 3 < W || !P.fold_halo;
    auto emit = [&](int dl, int dd, const float (&m)[4][4]) {   // m[c][j]: matched pixel of plane c for pixel j at disparity dd
        float4 r;
        float* rp = reinterpret_cast<float*>(&r);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = x4 + j;
            const bool interior = (SIGN < 0) ? (x >= dd) : (x < W - dd);
            const float c = cost4(s[0][j], s[1][j], s[2][j], s[3][j], m[0][j], m[1][j], m[2][j], m[3][j]);
            rp[j] = (x < W) ? (interior ? c : bord[j]) : 0.f;
        }
        float* o = out + (size_t)dl * slice;
        if (full_group) __stcs(reinterpret_cast<float4*>(o), r);   // streaming store: 2 GB per frame, next touched by the CVF kernel from HBM
        else {  // last, partial group of a row: columns >= W belong to the mirrored halo
#pragma unroll
            for (int j = 0; j < 4; ++j) if (x4 + j < W) o[j] = rp[j];
        }
        if (edge) {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (halo_off[j] != 0) o[halo_off[j]] = rp[j];
        }
    };

    // Interior fast path: a warp whose 128 pixels are matched inside the image for EVERY owned disparity (left volume: x >= the
    // largest d; right volume: x + largest d < W), away from the mirrored-halo columns and the ragged last group -- 13 of the 15
    // warps of a 1920-pixel row at D = 128 -- runs the same arithmetic without the per-pixel interior / x < W selects, the address
    // clamps and the halo branches: ~50 instead of ~98 instructions per 128-bit store.  (The one load past the matched range in
    // the last iteration lands in the row's mirrored halo and is never used.)
    if (GROUPED != 2) {
        const int dmax = P.d_begin + P.d_count - 1;
        const bool all_in = (SIGN < 0) ? (x4 >= dmax) : (x4 + 3 < W - dmax);
        if (__all_sync(0xffffffffu, all_in && !edge && x4 + 3 < W)) {
            const float* q[4] = {o0 + x4, o1 + x4, o2 + x4, o3 + x4};
            float w[4][4];
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j) w[c][j] = __ldg(q[c] + j + SIGN * d);
            float* o = out;
#pragma unroll 4
            for (int dl = 0; dl < P.d_count; ++dl, ++d, o += slice) {
                float4 r;
                r.x = cost4(s[0][0], s[1][0], s[2][0], s[3][0], w[0][0], w[1][0], w[2][0], w[3][0]);
                r.y = cost4(s[0][1], s[1][1], s[2][1], s[3][1], w[0][1], w[1][1], w[2][1], w[3][1]);
                r.z = cost4(s[0][2], s[1][2], s[2][2], s[3][2], w[0][2], w[1][2], w[2][2], w[3][2]);
                r.w = cost4(s[0][3], s[1][3], s[2][3], s[3][3], w[0][3], w[1][3], w[2][3], w[3][3]);
                __stcs(reinterpret_cast<float4*>(o), r);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (SIGN < 0) { w[c][3] = w[c][2]; w[c][2] = w[c][1]; w[c][1] = w[c][0]; w[c][0] = __ldg(q[c] - (d + 1)); }
                    else { w[c][0] = w[c][1]; w[c][1] = w[c][2]; w[c][2] = w[c][3]; w[c][3] = __ldg(q[c] + 3 + (d + 1)); }
                }
            }
            return;
        }
    }
    if (GROUPED == 1 && (P.d_begin & 3) == 0) {
        // Aligned shard (every D/N split of the BASELINE configs): four disparities per block.  With d % 4 == 0 the matched
        // positions x4 + j -/+ (d + k), k = 0..3, lie in two ALIGNED 4-column groups of the other view, and the next block
        // needs exactly one new group per plane -- one 128-bit load per plane per four slices instead of four scalar loads.
        // A group that lies wholly outside [0, W) is only ever matched by border pixels (whose cost is bord[j]), so its
        // address is clamped to a valid group and its contents do not matter.
        const int glast = (W - 1) & ~3;
        auto group = [&](const float* base, int g) {
            g = g < 0 ? 0 : (g > glast ? glast : g);
            return __ldg(reinterpret_cast<const float4*>(base + g));
        };
        const float* ob[4] = {o0, o1, o2, o3};
        int mb = x4 + SIGN * d;                       // first matched column of pixel 0 at this block's first disparity
        float4 lo[4], hi[4];                          // columns [g, g+4) and [g+4, g+8) with g = mb - 4 (left) / mb (right)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            lo[c] = group(ob[c], SIGN < 0 ? mb - 4 : mb);
            hi[c] = group(ob[c], SIGN < 0 ? mb : mb + 4);
        }
        for (int dl = 0; dl < P.d_count; dl += 4, d += 4) {
            float4 nx[4];                             // the group the next block adds, loaded ahead of this block's arithmetic
#pragma unroll
            for (int c = 0; c < 4; ++c) nx[c] = group(ob[c], SIGN < 0 ? mb - 8 : mb + 8);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (dl + k < P.d_count) {
                    float m[4][4];
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            // element of the 8-column window [lo | hi]: left j - k + 4 (1..7), right j + k (0..6)
                            const int e = SIGN < 0 ? j - k + 4 : j + k;
                            m[c][j] = e < 4 ? comp(lo[c], e) : comp(hi[c], e - 4);
                        }
                    emit(dl + k, d + k, m);
                }
            }
            mb += 4 * SIGN;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (SIGN < 0) { hi[c] = lo[c]; lo[c] = nx[c]; }
                else { lo[c] = hi[c]; hi[c] = nx[c]; }
            }
        }
        return;
    }
    // unaligned shard: window w[c][j] = other[c][x4 + j + SIGN*d], one new scalar load per plane per disparity
    float w[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int xo = clampx(x4 + j + SIGN * d);
        w[0][j] = __ldg(o0 + xo); w[1][j] = __ldg(o1 + xo); w[2][j] = __ldg(o2 + xo); w[3][j] = __ldg(o3 + xo);
    }
#pragma unroll 4
    for (int dl = 0; dl < P.d_count; ++dl, ++d) {
        emit(dl, d, w);
        // slide the window by one disparity
        if (SIGN < 0) {
            const int xo = clampx(x4 - (d + 1));
#pragma unroll
            for (int c = 0; c < 4; ++c) { w[c][3] = w[c][2]; w[c][2] = w[c][1]; w[c][1] = w[c][0]; }
            w[0][0] = __ldg(o0 + xo); w[1][0] = __ldg(o1 + xo); w[2][0] = __ldg(o2 + xo); w[3][0] = __ldg(o3 + xo);
        } else {
            const int xo = clampx(x4 + 3 + (d + 1));
#pragma unroll
            for (int c = 0; c < 4; ++c) { w[c][0] = w[c][1]; w[c][1] = w[c][2]; w[c][2] = w[c][3]; }
            w[0][3] = __ldg(o0 + xo); w[1][3] = __ldg(o1 + xo); w[2][3] = __ldg(o2 + xo); w[3][3] = __ldg(o3 + xo);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Exact 8x8 box mean of a plane at one pixel: sum of 64 reflected taps in fp64, * 1/64, one
// rounding to float == cv::boxFilter(CV_32F, Size(8,8)) (anchor (4,4), BORDER_REFLECT_101,
// double accumulation).  Used by the guide precompute and the naive cross-check kernels.
// ------------------------------------------------------------------------------------------
template <typename F>
__device__ __forceinline__ float box8_direct(F tap, int x, int y, int W, int H)
{
    double s = 0.0;
#pragma unroll 1
    for (int dy = -kBoxAnchor; dy < kBoxK - kBoxAnchor; ++dy) {
        const int yy = reflect101(y + dy, H);
        double rs = 0.0;
#pragma unroll
        for (int dx = -kBoxAnchor; dx < kBoxK - kBoxAnchor; ++dx)
            rs = __dadd_rn(rs, (double)tap(reflect101(x + dx, W), yy));
        s = __dadd_rn(s, rs);
    }
    return (float)__dmul_rn(s, 1.0 / 64.0);
}

// ------------------------------------------------------------------------------------------
// K2 guide precompute (CVF::preprocess, CVF.cpp:44-70) + the d-independent part of the 3x3
// solve of GuidedFilter_cv (CVF.cpp:117-126): the symmetric adjugate of (Sigma + eps I) and
// 1/det.  One fused kernel: fp64 horizontal 8-sums of the 9 planes (I_c, I_c*I_c') per row,
// fp64 running sums down the rows of a segment, then mean/var and the solve terms.
// guide planes (each [H][Wp]): 0-2 I, 3-5 mean_I, 6-11 adj (M00,M01,M02,M11,M12,M22), 12 1/det,
// 13-18 var_I (rr,rg,rb,gg,gb,bb; kept for parity reads).
// ------------------------------------------------------------------------------------------
constexpr int kGuideI = 0, kGuideMean = 3, kGuideAdj = 6, kGuideIdet = 12, kGuideVar = 13, kGuidePlanes = 19;

// d-independent per-pixel terms from the nine box means m[0..8] = mean of I0,I1,I2, I0I0,I0I1,I0I2,I1I1,I1I2,I2I2:
// var_I (CVF.cpp:60-69), symmetric adjugate of (Sigma + eps I) and 1/det (CVF.cpp:108-120).
__device__ __forceinline__ void guide_solve(const float m[9], float v[6], float M[6], float& idet)
{
    v[0] = fsub(m[3], fmul(m[0], m[0]));
    v[1] = fsub(m[4], fmul(m[0], m[1]));
    v[2] = fsub(m[5], fmul(m[0], m[2]));
    v[3] = fsub(m[6], fmul(m[1], m[1]));
    v[4] = fsub(m[7], fmul(m[1], m[2]));
    v[5] = fsub(m[8], fmul(m[2], m[2]));
    const float a11 = fadd(v[0], kGifEps), a12 = v[1], a13 = v[2];
    const float a21 = v[1], a22 = fadd(v[3], kGifEps), a23 = v[4];
    const float a31 = v[2], a32 = v[4], a33 = fadd(v[5], kGifEps);
    // cofactors exactly as written at CVF.cpp:117-146 (the adjugate is symmetric bit-for-bit
    // because each mirrored entry is the same two products in commuted order)
    M[0] = fsub(fmul(a33, a22), fmul(a32, a23));  // M00
    M[1] = fsub(fmul(a31, a23), fmul(a33, a21));  // M01
    M[2] = fsub(fmul(a32, a21), fmul(a31, a22));  // M02
    M[3] = fsub(fmul(a33, a11), fmul(a31, a13));  // M11
    M[4] = fsub(fmul(a31, a12), fmul(a32, a11));  // M12
    M[5] = fsub(fmul(a22, a11), fmul(a21, a12));  // M22
    // DET = a11*(a33*a22-a32*a23) - a21*(a33*a12-a32*a13) + a31*(a23*a12-a22*a13)  (CVF.cpp:117-119)
    const float t1 = fsub(fmul(a33, a12), fmul(a32, a13));
    const float t2 = fsub(fmul(a23, a12), fmul(a22, a13));
    const float det = fadd(fsub(fmul(a11, M[0]), fmul(a21, t1)), fmul(a31, t2));
    idet = __fdiv_rn(1.0f, det);  // CVF.cpp:120
}

// 8-wide window sums from the 4 column sums c[0..3] a lane owns:
// h[j] = columns (4l+j) .. (4l+j+7) = suffix_l[j..3] + total_{l+1} + prefix_{l+2}[0..j-1]   (4 fp64 shuffles)
__device__ __forceinline__ void hsum8(const double c[4], double h[4])
{
    const double P1 = c[0], P2 = c[0] + c[1], P3 = P2 + c[2], Tt = P3 + c[3];
    const double S1 = c[3], S2 = c[2] + c[3], S3 = c[1] + S2;
#if defined(PSM_KNOCKOUT) && (PSM_KNOCKOUT & 1)   // diagnostics only, see psm_cvf_stream.cuh
    const double Tn = Tt, Q1 = P1, Q2 = P2, Q3 = P3;
#else
    const double Tn = __shfl_down_sync(0xffffffffu, Tt, 1);
    const double Q1 = __shfl_down_sync(0xffffffffu, P1, 2);
    const double Q2 = __shfl_down_sync(0xffffffffu, P2, 2);
    const double Q3 = __shfl_down_sync(0xffffffffu, P3, 2);
#endif
    h[0] = Tt + Tn;
    h[1] = (S3 + Tn) + Q1;
    h[2] = (S2 + Tn) + Q2;
    h[3] = (S1 + Tn) + Q3;
}

// K2: warp = strip of 128 input columns (120 output columns) x seg_rows rows of one view;
// lane = 4 columns.  fp64 running column sums of the nine planes (newest row added, oldest row
// removed, both re-read from the I planes whose column halo is mirrored), 8-wide row sums by
// hsum8, then guide_solve per pixel; every plane is written with 128-bit stores.
constexpr int kGuideSegRows = 32;
constexpr int kGuideStripOut = 120;

struct GuideParams {
    float* guide[2];
    int* guide_flags;   // [2]: bit 1 raised when a mean / adjugate / 1/det value is not finite
    int W, H, Wp, nstrips, nseg, seg_rows;
};

__global__ void __launch_bounds__(128) guide_kernel(const GuideParams P)
{
    const int lane = threadIdx.x & 31;
    int task = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int strip = task % P.nstrips; task /= P.nstrips;
    const int seg = task % P.nseg;
    const int view = task / P.nseg;
    if (view >= 2) return;
    const int W = P.W, H = P.H, Wp = P.Wp;
    const size_t plane = (size_t)H * Wp;
    float* __restrict__ G = P.guide[view];
    const int cin = strip * kGuideStripOut - 4 + 4 * lane;  // input columns; outputs are cin+4 .. cin+7
    const int co = cin + 4;
    const bool in_ok = cin <= W + 4;       // beyond: nothing this lane feeds is ever stored (and the halo ends at W+7)
    const bool out_ok = lane <= 29 && co < W;
    const int y0 = seg * P.seg_rows;
    const int y1 = min(H, y0 + P.seg_rows);

    double V[9][4];
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) V[k][j] = 0.0;

    // (widening I and I*I' with the integer trick of the streaming filter was measured here and is 15 % SLOWER, 119 vs 104 us:
    // this kernel has XU slack and the IMAD.WIDEs land on the FMA pipe next to its multiplies -- plain conversions stay)
    auto feed = [&](int r, const bool add) {
        const size_t ro = (size_t)reflect101(r, H) * Wp + cin;
        float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = a4, c4 = a4;
        if (in_ok) {
            a4 = __ldg(reinterpret_cast<const float4*>(G + ro));
            b4 = __ldg(reinterpret_cast<const float4*>(G + plane + ro));
            c4 = __ldg(reinterpret_cast<const float4*>(G + 2 * plane + ro));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = comp(a4, j), b = comp(b4, j), c = comp(c4, j);
            const float t[9] = {a, b, c, fmul(a, a), fmul(a, b), fmul(a, c), fmul(b, b), fmul(b, c), fmul(c, c)};  // CVF.cpp:62
#pragma unroll
            for (int k = 0; k < 9; ++k) V[k][j] = add ? __dadd_rn(V[k][j], (double)t[k]) : __dsub_rn(V[k][j], (double)t[k]);
        }
    };

    for (int r = y0 - kBoxAnchor; r < y0 + kBoxK - kBoxAnchor - 1; ++r) feed(r, true);
    for (int y = y0; y < y1; ++y) {
        feed(y + kBoxK - kBoxAnchor - 1, true);
        float m[9][4];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            double h[4];
            hsum8(V[k], h);
#pragma unroll
            for (int j = 0; j < 4; ++j) m[k][j] = (float)__dmul_rn(h[j], kMeanPlain);
        }
        feed(y - kBoxAnchor, false);
        float o[16][4];  // planes kGuideMean .. kGuidePlanes-1
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float mm[9] = {m[0][j], m[1][j], m[2][j], m[3][j], m[4][j], m[5][j], m[6][j], m[7][j], m[8][j]};
            float v[6], M[6], idet;
            guide_solve(mm, v, M, idet);
            const bool px_ok = co + j < W;  // columns W .. W4-1 of the last group hold zeros
            o[0][j] = px_ok ? mm[0] : 0.f; o[1][j] = px_ok ? mm[1] : 0.f; o[2][j] = px_ok ? mm[2] : 0.f;
#pragma unroll
            for (int k = 0; k < 6; ++k) { o[3 + k][j] = px_ok ? M[k] : 0.f; o[10 + k][j] = px_ok ? v[k] : 0.f; }
            o[9][j] = px_ok ? idet : 0.f;
        }
        if (out_ok) {
            const size_t oo = (size_t)y * Wp + co;
            unsigned worst = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                *reinterpret_cast<float4*>(G + (size_t)(kGuideMean + k) * plane + oo) = make_float4(o[k][0], o[k][1], o[k][2], o[k][3]);
                if (k < 10) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) worst = max(worst, __float_as_uint(o[k][j]) & 0x7fffffffu);
                }
            }
            if (worst >= 0x7f800000u) atomicOr(P.guide_flags + view, 2);
        }
    }
}

// Per-voxel coefficient math shared by every CVF kernel: CVF.cpp:92-95 (cov), :121-146 (a),
// :152-155 (b).  mp/mIp are the four stage-1 box means at this voxel.
struct GuidePix { float mI0, mI1, mI2, M00, M01, M02, M11, M12, M22, idet; };

__device__ __forceinline__ void gif_coeffs(float mp, float mIp0, float mIp1, float mIp2, const GuidePix& g,
                                           float& a0, float& a1, float& a2, float& b)
{
    const float c0 = fsub(mIp0, fmul(g.mI0, mp));
    const float c1 = fsub(mIp1, fmul(g.mI1, mp));
    const float c2 = fsub(mIp2, fmul(g.mI2, mp));
    a0 = fmul(g.idet, fadd(fadd(fmul(c0, g.M00), fmul(c1, g.M01)), fmul(c2, g.M02)));
    a1 = fmul(g.idet, fadd(fadd(fmul(c0, g.M01), fmul(c1, g.M11)), fmul(c2, g.M12)));
    a2 = fmul(g.idet, fadd(fadd(fmul(c0, g.M02), fmul(c1, g.M12)), fmul(c2, g.M22)));
    b = fsub(fsub(fsub(mp, fmul(a0, g.mI0)), fmul(a1, g.mI1)), fmul(a2, g.mI2));
}

// ------------------------------------------------------------------------------------------
// Naive CVF (PSM_CVF_NAIVE): two unfused passes with direct 64-tap fp64 sums per voxel.
// Slow by construction (256 loads per voxel per pass); exists as an independent device-side
// implementation to cross-check the fused streaming kernel and for tiny images.
// ------------------------------------------------------------------------------------------
__global__ void cvf_naive_ab_kernel(const float* __restrict__ vol, const float* __restrict__ guide, size_t plane,
                                    int W, int H, int Wp, int nslices,
                                    float* __restrict__ ab /* [4][nslices][H][Wp] */)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int dl = blockIdx.z;
    if (x >= W) return;
    const float* p = vol + (size_t)dl * plane;
    const float* I0 = guide;
    const float* I1 = guide + plane;
    const float* I2 = guide + 2 * plane;
    const float mp = box8_direct([&](int xx, int yy) { return p[(size_t)yy * Wp + xx]; }, x, y, W, H);
    const float m0 = box8_direct([&](int xx, int yy) { const size_t o = (size_t)yy * Wp + xx; return fmul(I0[o], p[o]); }, x, y, W, H);
    const float m1 = box8_direct([&](int xx, int yy) { const size_t o = (size_t)yy * Wp + xx; return fmul(I1[o], p[o]); }, x, y, W, H);
    const float m2 = box8_direct([&](int xx, int yy) { const size_t o = (size_t)yy * Wp + xx; return fmul(I2[o], p[o]); }, x, y, W, H);
    const size_t o = (size_t)y * Wp + x;
    GuidePix g;
    g.mI0 = guide[(kGuideMean + 0) * plane + o]; g.mI1 = guide[(kGuideMean + 1) * plane + o]; g.mI2 = guide[(kGuideMean + 2) * plane + o];
    g.M00 = guide[(kGuideAdj + 0) * plane + o]; g.M01 = guide[(kGuideAdj + 1) * plane + o]; g.M02 = guide[(kGuideAdj + 2) * plane + o];
    g.M11 = guide[(kGuideAdj + 3) * plane + o]; g.M12 = guide[(kGuideAdj + 4) * plane + o]; g.M22 = guide[(kGuideAdj + 5) * plane + o];
    g.idet = guide[kGuideIdet * plane + o];
    float a0, a1, a2, b;
    gif_coeffs(mp, m0, m1, m2, g, a0, a1, a2, b);
    const size_t vs = (size_t)nslices * plane;
    const size_t oo = (size_t)dl * plane + o;
    ab[oo] = a0; ab[vs + oo] = a1; ab[2 * vs + oo] = a2; ab[3 * vs + oo] = b;
}

__global__ void cvf_naive_q_kernel(const float* __restrict__ ab, const float* __restrict__ guide, size_t plane,
                                   int W, int H, int Wp, int nslices, float* __restrict__ vol_out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int dl = blockIdx.z;
    if (x >= W) return;
    const size_t vs = (size_t)nslices * plane;
    const float* a0 = ab + (size_t)dl * plane;
    const float* a1 = a0 + vs;
    const float* a2 = a1 + vs;
    const float* b = a2 + vs;
    auto tapper = [&](const float* pl) {
        return box8_direct([&](int xx, int yy) { return pl[(size_t)yy * Wp + xx]; }, x, y, W, H);
    };
    const size_t o = (size_t)y * Wp + x;
    // q = box(b); q += box(a_c) * I_c for c = 0,1,2   (CVF.cpp:157-163)
    float q = tapper(b);
    q = fadd(q, fmul(tapper(a0), guide[o]));
    q = fadd(q, fmul(tapper(a1), guide[plane + o]));
    q = fadd(q, fmul(tapper(a2), guide[2 * plane + o]));
    vol_out[(size_t)dl * plane + o] = q;
}

// ------------------------------------------------------------------------------------------
// K4 WTA (DispSel.cpp:83-109): per pixel argmin over global d in [1, D), strict <, ties -> lowest d,
// minCost starts at +inf, result 0 when nothing compares below +inf.  One thread = 4 pixels.
// Emits the u8 map and/or the packed (cost,d) key used by the disparity-sharded reduction.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wta_kernel(const float* __restrict__ vol, int W, int H, int Wp, int d_begin, int d_count,
                                                  uint8_t* __restrict__ dis /* [H][W] or null */,
                                                  unsigned long long* __restrict__ keys /* [H][W] or null */)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    if (x4 >= W) return;
    const size_t plane = (size_t)H * Wp;
    const float* p = vol + (size_t)y * Wp + x4;
    float mc[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
    int md[4] = {0, 0, 0, 0};
    int dl = (d_begin == 0) ? 1 : 0;  // global d = 0 is never a candidate (DispSel.cpp:96)
#pragma unroll 8
    for (; dl < d_count; ++dl) {
        const float4 c = __ldg(reinterpret_cast<const float4*>(p + (size_t)dl * plane));
        const int d = d_begin + dl;
        if (c.x < mc[0]) { mc[0] = c.x; md[0] = d; }
        if (c.y < mc[1]) { mc[1] = c.y; md[1] = d; }
        if (c.z < mc[2]) { mc[2] = c.z; md[2] = d; }
        if (c.w < mc[3]) { mc[3] = c.w; md[3] = d; }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x = x4 + j;
        if (x < W) {
            if (dis) dis[(size_t)y * W + x] = (uint8_t)md[j];
            if (keys) keys[(size_t)y * W + x] = ((unsigned long long)float_order_key(mc[j]) << 32) | (unsigned)md[j];
        }
    }
}

// Sharded WTA fused with its exchange (reduce-scatter + all-gather of a min-reduction, hand-rolled
// over NVLink peer memory).  Pixels are partitioned into `nranks` chunks; rank r is the reducer of
// chunk r.
//   wta_scatter_kernel : the WTA scan of wta_kernel, then the packed (cost,d) minimum of every pixel
//                        is stored into its reducer's exchange block, slot [this rank][pixel-in-chunk]
//                        (peer device memory, plain st.global over NVLink) -- compute + scatter in one kernel;
//                        its last CTA then raises this rank's ARRIVE flag in every peer's block;
//   chunk_reduce_kernel: waits for the ARRIVE flags of all ranks (acquire loads on its own block), takes
//                        the min over the nranks slots of its chunk, stores the winning disparity (u8)
//                        into EVERY rank's result map -- reduce + gather in one kernel -- and its last CTA
//                        raises this rank's DONE flag in every peer's block;
//   p2p_wait_done_kernel: spins until all ranks' DONE flags carry this frame's sequence number; after it the local result
//                        maps are complete.  Launched only in front of a consumer of the maps (D2H fetch, post-filter);
//                        the NEXT frame's scatter kernel does the same wait itself before it overwrites the key slots.
// Flags are monotonically increasing frame sequence numbers (never reset), written with release and read
// with acquire semantics at system scope, so the exchange needs NO library collective or host barrier:
// a waiting kernel only ever waits for kernels of other GPUs that do not wait for it.
constexpr int kMaxRanks = 8;
constexpr int kScatterRows = 4;   // image rows per CTA of wta_scatter_kernel

struct P2pPeers {
    unsigned long long* keys[kMaxRanks];  // per rank: exchange block of this view  [nranks][chunk]
    unsigned char* maps[kMaxRanks];       // per rank: result map of this view       [H*W]
    unsigned* flags[kMaxRanks];           // per rank: flag words of this view: arrive[kMaxRanks], done[kMaxRanks]
    unsigned* counter;                    // device-local CTA counter of the launching rank (zero between launches)
    int nranks, rank;
    unsigned chunk;                       // pixels per chunk (last chunk may be partly unused)
    unsigned seq;                         // frame sequence number (>= 1)
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// called by every thread of the CTA after its peer stores: the LAST CTA of the launch publishes `seq` in slot
// flag_base[rank] of every peer's flag words
__device__ __forceinline__ void p2p_publish(const P2pPeers& peers, int flag_offset)
{
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        __threadfence_system();                       // this CTA's peer stores are visible system-wide ...
        const unsigned total = gridDim.x * gridDim.y;
        const unsigned prev = atomicAdd(peers.counter, 1u);
        if (prev == total - 1) {                      // ... before the last CTA (which observed every increment) raises the flags
            __threadfence();
            *peers.counter = 0;
            for (int r = 0; r < peers.nranks; ++r) st_release_sys(peers.flags[r] + flag_offset + peers.rank, peers.seq);
        }
    }
}

__global__ void __launch_bounds__(256) wta_scatter_kernel(const float* __restrict__ vol, int W, int H, int Wp, int d_begin, int d_count,
                                                          P2pPeers peers)
{
    if (peers.seq > 1) {   // the key slots written below were read by the reducers of the PREVIOUS frame: wait for their DONE flags
        if (threadIdx.x < (unsigned)peers.nranks) {
            const unsigned* f = peers.flags[peers.rank] + kMaxRanks + threadIdx.x;
            while ((int)(ld_acquire_sys(f) - (peers.seq - 1)) < 0) __nanosleep(64);
        }
        __syncthreads();
    }
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const bool active = x4 < W;  // inactive lanes still take part in the staged store below
    const size_t plane = (size_t)H * Wp;
    // Stage the warp's 128 keys in shared memory and send them out lane-contiguously: every store
    // instruction then writes 32 consecutive keys (256 contiguous bytes, whole 32-byte sectors) to
    // one peer, instead of 32 lanes x 8 bytes at a 32-byte stride (partial sectors over NVLink).
    __shared__ unsigned long long stage[8][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // a CTA handles kScatterRows consecutive rows: four times fewer CTAs have to fence and count in p2p_publish
    for (int y = blockIdx.y * kScatterRows; y < min(H, (int)(blockIdx.y + 1) * kScatterRows); ++y) {
        const float* p = vol + (size_t)y * Wp + (active ? x4 : 0);
        float mc[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
        int md[4] = {0, 0, 0, 0};
        int dl = (d_begin == 0) ? 1 : 0;
#pragma unroll 8
        for (; active && dl < d_count; ++dl) {
            const float4 c = __ldg(reinterpret_cast<const float4*>(p + (size_t)dl * plane));
            const int d = d_begin + dl;
            if (c.x < mc[0]) { mc[0] = c.x; md[0] = d; }
            if (c.y < mc[1]) { mc[1] = c.y; md[1] = d; }
            if (c.z < mc[2]) { mc[2] = c.z; md[2] = d; }
            if (c.w < mc[3]) { mc[3] = c.w; md[3] = d; }
        }
        __syncwarp();   // the previous row's staged keys have been sent
#pragma unroll
        for (int j = 0; j < 4; ++j)
            stage[warp][4 * lane + j] = ((unsigned long long)float_order_key(mc[j]) << 32) | (unsigned)md[j];
        __syncwarp();
        const int xw = x4 - 4 * lane;  // first column of this warp
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            const int idx = 32 * s4 + lane;
            const int x = xw + idx;
            if (x < W) {
                const unsigned pix = (unsigned)y * (unsigned)W + (unsigned)x;
                const unsigned owner = pix / peers.chunk;
                peers.keys[owner][(size_t)peers.rank * peers.chunk + (pix - owner * peers.chunk)] = stage[warp][idx];  // peer (or own) memory
            }
        }
    }
    if (peers.seq) p2p_publish(peers, 0);   // ARRIVE
}

__global__ void __launch_bounds__(256) chunk_reduce_kernel(P2pPeers peers, unsigned npix)
{
    if (peers.seq) {   // wait until every rank's minima for this frame have landed in this rank's block
        if (threadIdx.x < (unsigned)peers.nranks) {
            const unsigned* f = peers.flags[peers.rank] + threadIdx.x;
            while ((int)(ld_acquire_sys(f) - peers.seq) < 0) __nanosleep(64);
        }
        __syncthreads();
    }
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;   // pixel inside this rank's chunk
    const unsigned pix = (unsigned)peers.rank * peers.chunk + i;
    if (i < peers.chunk && pix < npix) {
        const unsigned long long* mine = peers.keys[peers.rank];
        unsigned long long k = __ldcg(mine + i);   // L2: the slots were written by peer GPUs
        for (int r = 1; r < peers.nranks; ++r) {
            const unsigned long long v = __ldcg(mine + (size_t)r * peers.chunk + i);
            k = v < k ? v : k;
        }
        const unsigned char d = (unsigned char)(k & 0xffu);
        for (int r = 0; r < peers.nranks; ++r) peers.maps[r][pix] = d;  // every rank receives the final map
    }
    if (peers.seq) p2p_publish(peers, kMaxRanks);   // DONE
}

__global__ void p2p_wait_done_kernel(const unsigned* flags /* this rank's flag words: [2 views][2 * kMaxRanks] */, int nranks, unsigned seq)
{
    if (threadIdx.x < 2u * (unsigned)nranks) {
        const unsigned view = threadIdx.x / nranks, r = threadIdx.x - view * nranks;
        const unsigned* f = flags + view * 2 * kMaxRanks + kMaxRanks + r;
        while ((int)(ld_acquire_sys(f) - seq) < 0) __nanosleep(64);
    }
}

// Final step of the sharded WTA: min over ranks of the packed keys, low 8 bits -> u8 map.
__global__ void keys_reduce_kernel(const unsigned long long* __restrict__ gathered, int nranks, size_t npix,
                                   uint8_t* __restrict__ dis)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    unsigned long long k = gathered[i];
    for (int r = 1; r < nranks; ++r) {
        const unsigned long long v = gathered[(size_t)r * npix + i];
        k = v < k ? v : k;
    }
    dis[i] = (uint8_t)(k & 0xffu);
}

}  // namespace psm
