// tools/cvf_knockout.cu -- diagnostic: time the shipped CVF kernels on a C4-size problem with single operation classes
// REMOVED (compile with -DPSM_KNOCKOUT=<mask>, see psm_cvf_stream.cuh).  The outputs are wrong by construction; only the
// time is of interest: what the kernel's duration is sensitive to, since no pipe and no stall reason dominates in ncu.
// Build + run (on the GPU box): see tools/gpu_knockout.sh
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../primestereomatch_b200/csrc/psm_cvf_stream.cuh"

using namespace psm;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void fill_kernel(float* p, size_t n, unsigned seed)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned h = (unsigned)i * 2654435761u + seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        p[i] = (h & 0xffffff) * (1.0f / 16777216.0f) + 0.01f;
    }
}

template <int S2M>
float run(const CvfParams& P, unsigned grid, int nthreads, int reps)
{
    auto kern = cvf_stream_kernel<3, 1, S2M, 1, (S2M == kS2Exact ? 1 : 0), 0>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) kern<<<grid, nthreads>>>(P);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) kern<<<grid, nthreads>>>(P);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main()
{
    const int W = 1920, H = 1080, D = 128;
    const int Wp = pitch_for_width(W);
    const size_t plane = (size_t)H * Wp;
    float *vin[2], *vout[2], *guide[2];
    int* flags;
    CK(cudaMalloc(&flags, 8)); CK(cudaMemset(flags, 0, 8));
    for (int v = 0; v < 2; ++v) {
        CK(cudaMalloc(&vin[v], (D * plane + 64) * sizeof(float)));
        CK(cudaMalloc(&vout[v], (D * plane + 64) * sizeof(float)));
        CK(cudaMalloc(&guide[v], (kGuidePlanes * plane + 64) * sizeof(float)));
        fill_kernel<<<1024, 256>>>(vin[v], D * plane + 64, 17u + v);
        fill_kernel<<<1024, 256>>>(guide[v], kGuidePlanes * plane + 64, 99u + v);
    }
    CK(cudaDeviceSynchronize());
    CvfParams P;
    for (int v = 0; v < 2; ++v) { P.vol_in[v] = vin[v] + kPadLeft; P.vol_out[v] = vout[v] + kPadLeft; P.guide[v] = guide[v] + kPadLeft; }
    P.guide_flags = flags;
    P.W = W; P.H = H; P.Wp = Wp; P.Dloc = D;
    P.nstrips = (W + kStripOut - 1) / kStripOut;
    const int nthreads = 96, wpc = 3;
    P.ndgroups = (D + wpc - 1) / wpc;
    P.nseg = 4; P.seg_rows = 272;
    P.remap_sms = 0; P.remap_ctas = 0;
    P.pack_gl = 0; P.pack_first = 0; P.pack_ndg = 0; P.pack_x0 = 0;
    P.one = 1.f; P.mone = -1.f;
    const unsigned grid = 2u * P.nseg * P.nstrips * P.ndgroups;
    const float ex = run<kS2Exact>(P, grid, nthreads, 5);
    const float mx = run<kS2Mixed>(P, grid, nthreads, 5);
    printf("PSM_KNOCKOUT=%2d  exact %.3f ms   mixed %.3f ms\n", PSM_KNOCKOUT, ex, mx);
    return 0;
}
