/*
 * prime_stereo_b200.h -- C-ABI of the B200-native STEREO_GIF hot path
 * (cost-volume construction -> guided-image-filter cost aggregation -> WTA).
 *
 * This is the drop-in boundary that sits under the reference's DispEst "_GPU" stage
 * methods (the `m`-toggle compute mode).  Plain C linkage, plain pointers and sizes,
 * no C++/torch/OpenCV types.  Every entry point returns 0 on success and a non-zero
 * PSM_E* code on failure (the reference's stage methods are `int`, 0 = success:
 * /root/reference/src/DispEst.cpp:275,307,327); psm_last_error() gives the text.
 *
 * Reference interfaces replaced (paths relative to /root/reference):
 *   psm_device_count      <-> openCLdevicepoll()                 src/main.cpp:29, src/oclUtil.cpp:18
 *   psm_create / destroy  <-> DispEst::DispEst(..., ocl=true) OpenCL part / ~DispEst
 *                                                                 src/DispEst.cpp:57-140, :145-162
 *   psm_set_images        <-> DispEst::setInputImages + H2D in CVC_cl::buildCV
 *                                                                 src/DispEst.cpp:164-170, src/CVC_cl.cpp:93-160
 *   psm_cost_const        <-> DispEst::CostConst_GPU             src/DispEst.cpp:272-276
 *   psm_cost_filter       <-> DispEst::CostFilter_GPU            src/DispEst.cpp:299-308
 *   psm_cost_filter_fgf   <-> DispEst::CostFilter_FGF -> FastGuidedFilter   src/DispEst.cpp:281-296, src/fastguidedfilter.cpp
 *   psm_disp_select       <-> DispEst::DispSelect_GPU + D2H in DispSel_cl::CVSelect
 *                                                                 src/DispEst.cpp:323-328, src/DispSel_cl.cpp:123-134
 *   psm_post_process      <-> DispEst::PostProcess_GPU -> PP::processDM -> JointWMF::filter
 *                                                                 src/DispEst.cpp:338-344, src/PP.cpp:402-425
 *   psm_read_cost_slice   <-> DispEst::printCV (cost-slice dump) src/DispEst.cpp:181-194
 *   psm_stage_ms          <-> cvc_time/cvf_time/dispsel_time     src/StereoMatch.cpp:209-241
 *
 * Numerics contract: results equal the reference's CPU building blocks src/CVC.cpp, src/CVF.cpp (GuidedFilter_cv, the
 * full-resolution guided filter) and src/DispSel.cpp bit for bit -- NOT the divergent OpenCL kernels (SURVEY.md section
 * 2.1), and NOT what the reference binary's CPU mode runs today: StereoMatch::compute calls CostFilter_FGF (the
 * sub-sampled Fast Guided Filter) because the CostFilter_CPU that would call GuidedFilter_cv is declared but never
 * defined; that branch is psm_cost_filter_fgf.  "Bit for bit" for the box filters means: fp64 sums of 64 floats are
 * exact -- and therefore independent of summation order -- unless a window spans more than 2^23 in magnitude; beyond that
 * OpenCV's order and this library's round differently in fp64 and the final float can differ in its last bit with
 * probability ~2^-28 per mean (never observed).  PSM_CVF_MIXED trades the bit-exact q for speed inside the tolerance
 * the task states (a, b bit-exact, |dq| <= 4e-6 on O(1) costs).  See DESIGN.md.
 *
 * Threading: one psm_ctx must not be used concurrently; it may be used from different
 * host threads over its life (every call binds its CUDA device first), matching how the
 * reference constructs DispEst on the HCI thread and runs stages on the worker thread.
 */
#ifndef PRIME_STEREO_B200_H
#define PRIME_STEREO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct psm_ctx psm_ctx;

enum {
    PSM_OK = 0,
    PSM_EINVAL = 1,   /* bad argument */
    PSM_ECUDA = 2,    /* CUDA runtime error (text in psm_last_error) */
    PSM_ESTATE = 3,   /* stage called out of order (images not set, ...) */
    PSM_ENOMEM = 4
};

/* Views */
enum { PSM_LEFT = 0, PSM_RIGHT = 1 };

/* psm_set_option keys */
enum {
    PSM_OPT_CVF_MODE = 1,  /* PSM_CVF_* below; default PSM_CVF_EXACT */
    PSM_OPT_GRAY_MODE = 2, /* 0: fma(c2,.114,fma(c0,.299,c1*.587)) (OpenCV SIMD/IPP builds, default)
                              1: (c0*.299+c1*.587)+c2*.114 (non-FMA OpenCV builds) */
    PSM_OPT_TIMING = 3,    /* 1: record cudaEvents around each stage (default 1) */
    PSM_OPT_P2P_SYNC = 4   /* fused multi-GPU exchange: 1 (default) = the kernels synchronise across ranks by
                              device-side flags in the exchange blocks (no library collective, no host barrier);
                              0 = the caller separates select / reduce / fetch by its own cross-rank barriers */
    /* keys 100..106 are kernel tuning knobs used by bench.py experiments (variant, rows per segment, extra
       shared memory, threads per CTA, block remap, packed remainder strips off, CVC build); they never
       change results */
};
enum {
    PSM_CVF_EXACT = 0, /* streaming fused kernel, every box sum accumulated in fp64: bit-exact q */
    PSM_CVF_MIXED = 1, /* the tolerance mode north_star allows for fp32: first box stage fp64 (a, b bit-exact), second
                          box stage fp32 with a fixed summation tree: q within ~3e-6 of EXACT, deterministic
                          (tests/mixed_model.py restates it on the CPU); same maps as EXACT on every tested frame */
    PSM_CVF_NAIVE = 2  /* unfused two-pass direct 64-tap fp64 kernels: slow device-side cross-check */
};

/* Diagnostic, needs no device: the work decomposition the streaming guided-filter kernel would use for a W x H x d_count
 * problem on a GPU with sm_count SMs (no reference counterpart; the reference's unit of parallel work is the slice,
 * src/DispEst.cpp:235-268).  out[0..9] = threads per CTA, full 112-column strips, slice groups, row segments, rows per
 * segment, lanes per slice of the packed remainder (0 = none), first output-aligned column of the packed remainder,
 * packed slice groups, first packed CTA, grid size.  The CPU tests check its invariants (every column and row of every
 * slice is produced exactly once). */
int psm_cvf_plan(int width, int height, int d_count, int sm_count, int no_pack, int* out, int n);

/* Number of usable CUDA devices (0 if none / no driver): gates the `m` toggle. */
int psm_device_count(void);

/* Create a context for W x H images and disparities [0, max_disp) on CUDA device `device`.
 * Allocates all device state (planar images, gradients, guide planes, both cost volumes,
 * disparity maps).  max_disp <= 256 (u8 disparity maps). */
int psm_create(psm_ctx** out, int width, int height, int max_disp, int device);

/* Disparity-sharded context (multi-GPU, one context per GPU): this context owns the global
 * slices [d_begin, d_begin + d_count) of a max_disp-deep problem.  psm_create == shard [0,max_disp). */
int psm_create_sharded(psm_ctx** out, int width, int height, int max_disp,
                       int d_begin, int d_count, int device);

int psm_destroy(psm_ctx* ctx);

int psm_set_option(psm_ctx* ctx, int key, int value);

/* Use an existing CUDA stream (cudaStream_t passed as void*) for all work; NULL restores the
 * context's own stream.  Lets a caller time with its own events on that stream. */
int psm_set_stream(psm_ctx* ctx, void* cuda_stream);

/* Inputs: interleaved 3-channel float images (cv::Mat CV_32FC3, BGR, values in [0,1]) in HOST
 * memory, row steps in BYTES.  Copies H2D and builds planar channels + x-gradients
 * (CVC::preprocess, src/CVC.cpp:41-46) on the device. */
int psm_set_images(psm_ctx* ctx, const float* left, size_t left_step,
                   const float* right, size_t right_step);

/* Same, 8-bit interleaved BGR input; the convertTo(CV_32F, 1/255.0f) of
 * src/StereoMatch.cpp:193-197 is done on the device (4x less H2D traffic). */
int psm_set_images_u8(psm_ctx* ctx, const uint8_t* left, size_t left_step,
                      const uint8_t* right, size_t right_step);

/* Same as psm_set_images but the interleaved float images already live in DEVICE memory. */
int psm_set_images_device(psm_ctx* ctx, const float* d_left, size_t left_step,
                          const float* d_right, size_t right_step);

/* Pipelined upload (extension; the reference's DispEst is synchronous per frame): start the H2D copy
 * of the NEXT frame on the context's copy stream into a second staging set while the current frame
 * is still being computed, then make it the current frame.  Host buffers must be page-locked for the
 * copy to overlap and must stay valid until psm_set_images_commit has returned.
 *   psm_set_images_async(k+1) ... stages of frame k ... psm_set_images_commit()  ->  frame k+1 is current
 * commit = compute stream waits for the upload, then the same ingest kernels as psm_set_images. */
int psm_set_images_async(psm_ctx* ctx, const float* left, size_t left_step,
                         const float* right, size_t right_step);
int psm_set_images_u8_async(psm_ctx* ctx, const uint8_t* left, size_t left_step,
                            const uint8_t* right, size_t right_step);
int psm_set_images_commit(psm_ctx* ctx);

/* Stage 1: both raw cost volumes (CostConst_GPU). Asynchronous on the context stream. */
int psm_cost_const(psm_ctx* ctx);

/* Stage 2: guide precompute + guided filter of every slice of both volumes, in place
 * (CostFilter_GPU). Asynchronous on the context stream. */
int psm_cost_filter(psm_ctx* ctx);

/* Stage 2, Fast-Guided-Filter variant (DispEst::CostFilter_FGF, what the reference's CPU branch runs today): every
 * slice is sub-sampled by `sub_sample_rate` (1, 2, 4 or 8; the reference's `s` key cycles 2, 4, 8, default 4),
 * guided-filtered there with a (2*(8/s)+1)^2 box and the coefficient means are bilinearly up-sampled.  In place,
 * asynchronous.  Results equal the CPU restatement orc_cost_filter_fgf (test infrastructure, oracle/) bit for bit, which is pinned against
 * OpenCV's own (non-IPP) cv::blur / cv::resize arithmetic (tests/golden/make_golden_fgf.py). */
int psm_cost_filter_fgf(psm_ctx* ctx, int sub_sample_rate);

/* Stage 3: WTA over d in [1, max_disp) for both views; copies the u8 maps to HOST memory
 * (row steps in bytes) and synchronises (DispSelect_GPU).  Only valid on an unsharded context. */
int psm_disp_select(psm_ctx* ctx, uint8_t* left, size_t left_step,
                    uint8_t* right, size_t right_step);

/* Stage 3 with the D2H copies enqueued but NOT synchronised: the host maps are valid after psm_sync
 * (or any later synchronising call).  Page-locked host memory keeps the call asynchronous. */
int psm_disp_select_async(psm_ctx* ctx, uint8_t* left, size_t left_step,
                          uint8_t* right, size_t right_step);

/* Stage 3 without the D2H copy: maps stay on the device (see psm_device_ptr). Asynchronous. */
int psm_disp_select_device(psm_ctx* ctx);

/* Stage 4: post-processing of both maps -- the joint weighted-median filter of PP::processDM (window 19x19,
 * "exp" colour weights on the 6-bit-quantised left / right image, JointWMF defaults), on the device; the
 * reference runs this stage on the CPU even in GPU mode (src/DispEst.cpp:338-344).  Input: the maps of the last
 * selection stage (for a sharded context with the fused exchange: the complete maps of psm_disp_reduce_p2p);
 * output: filtered u8 maps copied to HOST memory (row steps in bytes), synchronises.
 * Parity: equals the reference's JointWMF whenever the feature image has <= 256 distinct 6-bit colours; beyond
 * that the reference clusters colours with an RNG-seeded cv::kmeans (an approximation by its own documentation,
 * JointWMF.h:70-72) while this stage evaluates the un-clustered weights -- see DESIGN.md. */
int psm_post_process(psm_ctx* ctx, uint8_t* left, size_t left_step, uint8_t* right, size_t right_step);
/* Same without the D2H copy (maps stay on the device: psm_device_ptr selectors 4, 5). Asynchronous. */
int psm_post_process_device(psm_ctx* ctx);

/* Sharded stage 3a: per-pixel packed minima over this context's slices, written to DEVICE
 * buffers of H*W uint64 each:  key = (order_preserving_u32(cost) << 32) | global_d.
 * min() over ranks' keys reproduces the reference's strict-< / lowest-d tie-break. Asynchronous. */
int psm_disp_select_keys(psm_ctx* ctx, uint64_t* d_keys_left, uint64_t* d_keys_right);

/* Sharded stage 3b: reduce `nranks` gathered key planes (DEVICE, [nranks][H*W] per view) to the
 * final u8 maps in HOST memory and synchronise. */
int psm_disp_reduce_keys(psm_ctx* ctx, const uint64_t* d_gathered_left,
                         const uint64_t* d_gathered_right, int nranks,
                         uint8_t* left, size_t left_step, uint8_t* right, size_t right_step);

/* ---- Sharded stage 3 fused with its exchange over NVLink peer memory ------------------------------
 * Instead of "local WTA kernel, then library collectives", the min-reduction over ranks is done by two
 * kernels that read/write PEER device memory directly (reduce-scatter + all-gather, hand-rolled):
 *   psm_disp_select_keys_p2p : WTA over this rank's slices; the packed minimum of every pixel is stored
 *                              straight into the exchange block of the rank that reduces that pixel;
 *   psm_disp_reduce_p2p      : this rank reduces its pixel chunk over all ranks' minima and stores the
 *                              winning disparity into EVERY rank's result map;
 *   psm_disp_fetch_p2p       : copy this rank's (complete) result maps to host memory and synchronise.
 * The three kernels order themselves across ranks through ARRIVE / DONE flag words in the exchange blocks
 * (release stores over NVLink, acquire loads; PSM_OPT_P2P_SYNC = 1, default): select -> reduce -> fetch can be
 * enqueued back to back on every rank without any library collective or host barrier, provided every rank runs
 * the same sequence of frames and each rank's GPU can run its kernels independently of the others (one context
 * per GPU).  With PSM_OPT_P2P_SYNC = 0 a cross-rank barrier must separate select from reduce and reduce from
 * fetch.  All launches of one context must stay on ONE stream (psm_set_stream); collectives the caller issues
 * around them must be ordered against that stream.  Every rank owns one exchange block
 * (psm_p2p_create_buffer), shared with the other per-GPU processes through CUDA IPC (psm_ipc_export /
 * psm_ipc_import).  At most 8 ranks. */
int psm_p2p_create_buffer(psm_ctx* ctx, int nranks, void** d_buffer);
int psm_ipc_export(psm_ctx* ctx, void* d_ptr, unsigned char handle_out[64]);
int psm_ipc_import(psm_ctx* ctx, const unsigned char handle[64], void** d_ptr);
/* d_buffers[r] = exchange block of rank r as seen from this process (own pointer for r == rank). */
int psm_p2p_set_peers(psm_ctx* ctx, void* const* d_buffers, int nranks, int rank);
int psm_disp_select_keys_p2p(psm_ctx* ctx);
int psm_disp_reduce_p2p(psm_ctx* ctx);
int psm_disp_fetch_p2p(psm_ctx* ctx, uint8_t* left, size_t left_step, uint8_t* right, size_t right_step);

/* Debug / parity reads (synchronising). `d` is a GLOBAL disparity index owned by this context. */
int psm_read_cost_slice(psm_ctx* ctx, int view, int d, float* dst, size_t dst_step);
int psm_write_cost_slice(psm_ctx* ctx, int view, int d, const float* src, size_t src_step);
/* plane ids: 0-2 I (split channels), 3-5 mean_I, 6-11 var_I (rr,rg,rb,gg,gb,bb), 12 x-gradient */
int psm_read_guide_plane(psm_ctx* ctx, int view, int plane, float* dst, size_t dst_step);
/* guided-filter coefficients of one raw slice (a0,a1,a2,b; each H*W floats, dense rows), computed
 * from the CURRENT content of slice d (call between psm_cost_const and psm_cost_filter). */
int psm_read_ab_slice(psm_ctx* ctx, int view, int d, float* a3, float* b);

/* Device pointers for zero-copy callers / benchmarks. what: 0 left volume, 1 right volume,
 * 2 left u8 map, 3 right u8 map, 4 / 5 post-processed left / right u8 map.  pitch_elems receives the row pitch in elements. */
int psm_device_ptr(psm_ctx* ctx, int what, void** ptr, size_t* pitch_elems);

/* Last measured stage durations in milliseconds (cudaEvent pairs on the context stream):
 * stage 0 ingest (H2D + planar/gradient), 1 CVC, 2 CVF (guide + filter), 3 WTA, 4 CVF filter kernel only, 5 post-process.
 * Synchronises on the stage's end event. */
int psm_stage_ms(psm_ctx* ctx, int stage, float* ms);

/* Number of kernels this context launched since creation (bench.py "gpu_launches"). */
int psm_launch_count(psm_ctx* ctx, uint64_t* n);

int psm_sync(psm_ctx* ctx);

/* Human-readable text of the last error on this context (or the last creation error if ctx==NULL). */
const char* psm_last_error(psm_ctx* ctx);

/* Library build info string: version, arch, compile flags. */
const char* psm_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* PRIME_STEREO_B200_H */
