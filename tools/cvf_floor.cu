// cvf_floor.cu -- compute floor of the exact guided-filter row: one warp-row of the fused CVF kernel
// reduced to its mandatory arithmetic (per lane and row: 64 f32->f64 and 32 f64->f32 conversions,
// 160 DADD + 32 DMUL, 64 SHFL, ~170 fp32 ops) with NO memory traffic, run at the same occupancy
// (4 CTAs x 96 threads per SM).  Prints cycles per warp-row per SM; the real kernel needs
// (kernel time x SM clock) / (warp-rows per SM) -- see DESIGN.md section 6.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void hsum8(const double c[4], double h[4])
{
    const double P1 = c[0], P2 = c[0] + c[1], P3 = P2 + c[2], Tt = P3 + c[3];
    const double S1 = c[3], S2 = c[2] + c[3], S3 = c[1] + S2;
    const double Tn = __shfl_down_sync(0xffffffffu, Tt, 1);
    const double Q1 = __shfl_down_sync(0xffffffffu, P1, 2), Q2 = __shfl_down_sync(0xffffffffu, P2, 2), Q3 = __shfl_down_sync(0xffffffffu, P3, 2);
    h[0] = Tt + Tn; h[1] = (S3 + Tn) + Q1; h[2] = (S2 + Tn) + Q2; h[3] = (S1 + Tn) + Q3;
}

__device__ __forceinline__ double widen_pos(float f)
{
    const unsigned u = __float_as_uint(f);
    double d = __hiloint2double((int)((u >> 3) + 0x38000000u), (int)(u << 29));
    if (u - 0x00800000u >= 0x7f000000u) d = (double)f;
    return d;
}
template <int IW> __device__ __forceinline__ double wid(float f) { return IW ? widen_pos(f) : (double)f; }

template <int MAXT, int MINB, int IW>
__global__ void __launch_bounds__(MAXT, MINB) floor_kernel(float* out, const float* in, int rows, long long* cyc)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float x[4][4], g[10][4];
    double S1[4][4], S2[4][4];
    for (int q = 0; q < 4; ++q) for (int j = 0; j < 4; ++j) { x[q][j] = in[(tid + 4 * q + j) & 1023]; S1[q][j] = 0; S2[q][j] = 0; }
    for (int q = 0; q < 10; ++q) for (int j = 0; j < 4; ++j) g[q][j] = in[(tid * 3 + q * 4 + j) & 1023];
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < rows; ++r) {
        float av[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 4; ++j) S1[q][j] = __dadd_rn(S1[q][j], wid<IW>(__fmul_rn(x[q][j], g[q][j])));      // widen newest
        float m[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { double h[4]; hsum8(S1[q], h);
#pragma unroll
            for (int j = 0; j < 4; ++j) m[q][j] = (float)__dmul_rn(h[j], 1.0 / 64.0); }                          // narrow
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 4; ++j) S1[q][j] = __dsub_rn(S1[q][j], wid<IW>(__fmul_rn(x[q][j], g[q + 4][j])));  // widen oldest
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // cov, a, b: 33 fp32 ops per column
            const float c0 = __fsub_rn(m[1][j], __fmul_rn(g[0][j], m[0][j])), c1 = __fsub_rn(m[2][j], __fmul_rn(g[1][j], m[0][j])), c2 = __fsub_rn(m[3][j], __fmul_rn(g[2][j], m[0][j]));
            av[0][j] = __fmul_rn(g[9][j], __fadd_rn(__fadd_rn(__fmul_rn(c0, g[3][j]), __fmul_rn(c1, g[4][j])), __fmul_rn(c2, g[5][j])));
            av[1][j] = __fmul_rn(g[9][j], __fadd_rn(__fadd_rn(__fmul_rn(c0, g[4][j]), __fmul_rn(c1, g[6][j])), __fmul_rn(c2, g[7][j])));
            av[2][j] = __fmul_rn(g[9][j], __fadd_rn(__fadd_rn(__fmul_rn(c0, g[5][j]), __fmul_rn(c1, g[7][j])), __fmul_rn(c2, g[8][j])));
            av[3][j] = __fsub_rn(__fsub_rn(__fsub_rn(m[0][j], __fmul_rn(av[0][j], g[0][j])), __fmul_rn(av[1][j], g[1][j])), __fmul_rn(av[2][j], g[2][j]));
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 4; ++j) S2[q][j] = __dsub_rn(__dadd_rn(S2[q][j], (double)av[q][j]), (double)x[q][j]);  // widen newest + oldest
        float mb[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { double h[4]; hsum8(S2[q], h);
#pragma unroll
            for (int j = 0; j < 4; ++j) mb[q][j] = (float)__dmul_rn(h[j], 1.0 / 64.0); }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float qv = __fadd_rn(mb[3][j], __fmul_rn(mb[0][j], g[0][j]));
            qv = __fadd_rn(qv, __fmul_rn(mb[1][j], g[1][j]));
            qv = __fadd_rn(qv, __fmul_rn(mb[2][j], g[2][j]));
            x[j][0] = __fadd_rn(x[j][0], qv * 1e-30f);  // keep the result live, keep the data finite
        }
    }
    const long long t1 = clock64();
    float acc = 0;
    for (int q = 0; q < 4; ++q) for (int j = 0; j < 4; ++j) acc += x[q][j] + (float)S1[q][j] + (float)S2[q][j];
    out[tid] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MAXT, int MINB, int IW>
void run(const char* name, int threads, int ctas_per_sm, size_t smem)
{
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int rows = 2000, ctas = nsm * ctas_per_sm;
    float *out, *in; long long* cyc;
    cudaMalloc(&out, (size_t)ctas * threads * 4); cudaMalloc(&in, 4096); cudaMalloc(&cyc, ctas * 8);
    float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = 0.25f + 0.001f * i;
    cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
    auto k = floor_kernel<MAXT, MINB, IW>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int it = 0; it < 2; ++it) k<<<ctas, threads, smem>>>(out, in, rows, cyc);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<<<ctas, threads, smem>>>(out, in, rows, cyc); cudaEventRecord(b); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double warps = ctas_per_sm * threads / 32.0, clk = ms * 1e-3 * 1.93e9, per = clk / (warps * rows);
    printf("%-34s warps/SM=%4.0f  %7.1f SM-cycles per warp-row -> C4 floor %.2f ms   (%s)\n", name, warps, per,
           2.0 * 18 * 128 * 1124 / nsm * per / 1.93e9 * 1e3, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(in); cudaFree(cyc);
}

int main()
{
    run<128, 3, 0>("shipped mix, 4x96 thr", 96, 4, 49152);
    run<128, 3, 0>("shipped mix, 3x128 thr", 128, 3, 65536);
    run<128, 3, 0>("shipped mix, 2x96 thr (6 warps)", 96, 2, 98304);
    run<128, 4, 0>("<=128 regs, 4x128 thr (16 warps)", 128, 4, 32768);
    run<128, 4, 0>("<=128 regs, 5x96 thr (15 warps)", 96, 5, 40000);
    run<256, 2, 0>("<=128 regs, 2x256 thr (16 warps)", 256, 2, 65536);
    run<128, 3, 1>("int widening stage 1, 4x96 thr", 96, 4, 49152);
    run<128, 4, 1>("int widening, <=128 regs 16 warps", 128, 4, 32768);
    return 0;
}
