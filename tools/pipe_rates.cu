// pipe_rates.cu -- B200 micro-benchmark of the pipes the exact box filter leans on:
// DADD, F2F.F64.F32, F2F.F32.F64, DMUL, integer widening, f64 warp shuffles, FADD2.
// Prints ops/clk/SM so the CVF kernel design can be budgeted against measured rates.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int OP>
__global__ void k(float* out, const float* in, long long* cyc)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float f[ILP]; double d[ILP]; unsigned u[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { f[i] = in[(tid + i) & 1023]; d[i] = (double)f[i] + 1.0; u[i] = __float_as_uint(f[i]); }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) d[i] = __dadd_rn(d[i], 1.25);                         // DADD
            if (OP == 1) { d[i] = (double)f[i]; f[i] = __uint_as_float(__float_as_uint(f[i]) ^ (unsigned)(__double2hiint(d[i]) & 1)); }  // F2F.F64.F32 (+ALU dep)
            if (OP == 2) { f[i] = (float)d[i]; d[i] = __hiloint2double(__double2hiint(d[i]) ^ (__float_as_int(f[i]) & 1), __double2loint(d[i])); } // F2F.F32.F64
            if (OP == 3) d[i] = __dmul_rn(d[i], 1.0000001);                     // DMUL
            if (OP == 4) {                                                      // integer widening f32->f64 (non-negative normal or zero)
                unsigned v = u[i];
                unsigned hi = (v >> 3) + (v ? 0x38000000u : 0u);
                unsigned lo = v << 29;
                d[i] = __dadd_rn(d[i], __hiloint2double((int)hi, (int)lo));
                u[i] = v + 8;
            }
            if (OP == 5) d[i] = __shfl_down_sync(0xffffffffu, d[i], 1);          // f64 shuffle (2x SHFL)
            if (OP == 6) { float2 a = make_float2(f[i], f[i]); a = __fadd2_rn(a, make_float2(1.5f, 2.5f)); f[i] = a.x + a.y; } // FADD2 + FADD
            if (OP == 7) f[i] = __fadd_rn(f[i], 1.5f);                          // FADD
            if (OP == 8) { d[i] = __dadd_rn(d[i], (double)f[i]); f[i] = __fadd_rn(f[i], 1.0f); }  // cvt + DADD + FADD mix
            if (OP == 9) f[i] = __shfl_down_sync(0xffffffffu, f[i], 1);          // f32 shuffle
        }
    }
    long long t1 = clock64();
    float acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc += f[i] + (float)d[i] + (float)u[i];
    out[tid] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP> void run(const char* name, int threads, int ops_per_iter)
{
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    float *out, *in; long long* cyc;
    cudaMalloc(&out, (size_t)nsm * threads * 4); cudaMalloc(&in, 4096); cudaMemset(in, 0x3f, 4096);
    cudaMalloc(&cyc, nsm * 8);
    k<OP><<<nsm, threads>>>(out, in, cyc);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<OP><<<nsm, threads>>>(out, in, cyc); cudaEventRecord(b); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h[256]; cudaMemcpy(h, cyc, nsm * 8, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < nsm; ++i) c += h[i]; c /= nsm;
    double ops = (double)ITERS * ILP * threads * ops_per_iter;
    printf("%-28s threads/SM=%4d  %8.2f ops/clk/SM  (%.0f cyc, %.3f ms, %.2f GHz eff)\n", name, threads, ops / c, c, ms, c / (ms * 1e6));
    cudaFree(out); cudaFree(in); cudaFree(cyc);
}

int main()
{
    for (int threads : {256, 1024}) {
        run<0>("DADD", threads, 1);
        run<1>("F2F.F64.F32 (+LOP)", threads, 1);
        run<2>("F2F.F32.F64 (+LOP)", threads, 1);
        run<3>("DMUL", threads, 1);
        run<4>("int-widen + DADD", threads, 1);
        run<5>("SHFL f64", threads, 1);
        run<6>("FADD2+FADD", threads, 1);
        run<7>("FADD", threads, 1);
        run<8>("cvt+DADD+FADD", threads, 1);
        run<9>("SHFL f32", threads, 1);
    }
    return 0;
}
