/* cadl -- camera-aware depth-loss kernels for NVIDIA B200 (sm_100a): the C ABI.
 *
 * This is the drop-in boundary of the hot path.  The reference has no FFI for this path: its
 * loss/metric classes are header-only LibTorch op chains (src/loss/depth_loss.h,
 * src/evaluation/depth_metrics.h).  The replacement headers under <pkg>/host/ keep those class
 * signatures and call ONLY the functions declared here; each entry point names the reference
 * interface it replaces (file:line relative to the reference root).
 *
 * Conventions
 *   - plain pointers and sizes; no torch types.  All data pointers are DEVICE pointers on the
 *     current CUDA device unless marked "host".  fp32, NCHW contiguous:
 *       pred (B,1,H,W)  gt (B,1,H,W)  rgb (B,3,H,W)  K (B,3,3) or (3,3) row-major
 *       mask (B,1,H,W) bytes (torch::kBool storage), may be NULL
 *   - the caller owns every buffer (results, gradient, workspace); kernels never allocate.
 *   - every call is asynchronous on `stream` (a cudaStream_t) and re-entrant; two calls may run
 *     concurrently iff they use different workspaces.
 *   - return value: 0 = ok, CADL_ERR_* (<1000) = argument error detected before launch,
 *     1000 + cudaError_t = launch/runtime error.  Nothing throws.  cadl_error_string() names it.
 *   - there is NO CPU fallback: without a CUDA device the calls return 1000+cudaErrorNoDevice.
 */
#ifndef CADL_H_
#define CADL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CADL_VERSION 100

typedef void* cadl_stream_t; /* cudaStream_t */

/* loss terms (bitmask) */
#define CADL_TERM_SI 1u      /* ScaleInvariantLoss      depth_loss.h:20-69   */
#define CADL_TERM_GRAD 2u    /* GradientMatchingLoss    depth_loss.h:82-167  */
#define CADL_TERM_SMOOTH 4u  /* SmoothnessLoss          depth_loss.h:178-238 */
#define CADL_TERM_REPROJ 8u  /* ReprojectionLoss        depth_loss.h:255-355 */
#define CADL_TERM_ALL 15u

/* metric variants (bitmask) */
#define CADL_METRICS_EVAL 1u  /* DepthMetrics::compute            depth_metrics.h:40-88 */
#define CADL_METRICS_TRAIN 2u /* trainers' computeDepthMetrics    tensorboard_trainer_enhanced.h:400-439 */

#define CADL_MAX_SCALES 4

enum {
    CADL_OK = 0,
    CADL_ERR_NULL = 1,        /* a required pointer is NULL */
    CADL_ERR_SHAPE = 2,       /* B,H,W out of range for the requested terms */
    CADL_ERR_WORKSPACE = 3,   /* workspace too small or misaligned */
    CADL_ERR_UNSUPPORTED = 4, /* term combination / num_scales not provided by this build */
    CADL_ERR_ALIGN = 5,       /* a data pointer is not 4-byte aligned */
    CADL_ERR_CUDA = 1000      /* 1000 + cudaError_t */
};

/* The constructor arguments of the reference classes, as one POD. */
typedef struct cadl_params {
    uint32_t terms;   /* CADL_TERM_* to evaluate */
    uint32_t metrics; /* CADL_METRICS_* to evaluate in the same pass over pred/gt (0 = none) */
    float w_si, w_grad, w_smooth, w_reproj; /* CombinedDepthLoss ctor, depth_loss.h:368-371 */
    float si_lambda;                        /* ScaleInvariantLoss ctor, depth_loss.h:22 (0.5) */
    float eps_si, eps_grad, eps_smooth, eps_reproj; /* depth_loss.h:22,84,180,257 (1e-6 each) */
    int32_t num_scales;                     /* GradientMatchingLoss ctor, depth_loss.h:84 (4) */
    int32_t k_batched;                      /* 1: K is (B,3,3); 0: (3,3) broadcast, depth_loss.h:278-280 */
    float min_depth, max_depth;             /* DepthMetrics::compute, depth_metrics.h:44-45 (0.1, 10) */
    float upstream;                         /* dL/dloss folded into the gradient (1.0 = loss.backward()) */
    /* exact-global-batch mode (SURVEY 8e mode B): this rank holds B of global_B images */
    int32_t global_B;                       /* 0 or B: single process */
    /* split API only: cadl_stack_prepare was called for this step (same pred/gt/shape/params/workspace/stream, same
     * host thread) and returned CADL_OK -- cadl_stack_reduce then shares the SMs with it and cadl_stack_grad joins it */
    int32_t pyramid_prepared;
} cadl_params;

/* Everything the hot path returns, written on the device by the last block of the last kernel.
 * float members are what the reference returns as fp32 tensors / .item<float>(); the fp64 and
 * integer members are the exact accumulators they were rounded from. */
typedef struct cadl_results {
    /* losses (un-weighted terms, and the weighted total of the requested terms) */
    float loss_total, loss_si, loss_grad, loss_smooth, loss_reproj;
    float _pad0[3];
    double d_total, d_si, d_grad, d_smooth, d_reproj;
    int64_t n_si, n_reproj; /* valid-pixel counts (depth_loss.h:52, :323-325) */
    /* DepthMetrics::compute: abs_rel sq_rel rmse rmse_log mae log10 d1 d2 d3 n_valid mean_pred mean_gt */
    float eval[12];
    int64_t eval_counts[4]; /* n_valid, #(ratio<1.25), #(<1.25^2), #(<1.25^3) */
    /* computeDepthMetrics: abs_rel sq_rel rmse rmse_log a1 a2 a3 (+ pad) */
    float train[8];
    int64_t train_counts[4];
} cadl_results;

void cadl_default_params(cadl_params* p);
int cadl_version(void);
/* sizeof(cadl_params) / sizeof(cadl_results) of this build: lets an FFI binding verify its mirror */
size_t cadl_sizeof_params(void);
size_t cadl_sizeof_results(void);
const char* cadl_error_string(int code);
#ifdef CADL_DEBUG
/* Debug build only (libcadl_dbg.so; the product library exports neither).
 * cadl_debug_force_generic, bit mask: 1 = always take the generic phase-B kernel (any shape/alignment) instead of the
 * aligned fast path; 8 = fast path as ONE tile kernel (cadl_phase_b_fast.cuh) instead of the pyramid + streaming
 * kernels; together with 8: 2 = stage tiles with cp.async instead of TMA; 16 = no programmatic dependent launch;
 * 32 = the pooled-pyramid kernels in line on the caller's stream instead of beside phase A on the auxiliary stream;
 * 64 = reprojection alone with the separate count kernel instead of the single cooperative launch; 128 = the loss
 * statistics from phase A (the kernels of the split API) instead of the pooled-sum pass.  All variants must
 * produce the same values (tests/test_math_gpu.py).  Process-global.
 * cadl_debug_kernel_times: with enable != 0 every following cadl_stack_fwd_bwd call records a CUDA event after each of
 * its launches and SYNCHRONISES at the end.  Returns the number of intervals of the LAST timed call and copies up to
 * cap of them: ms_out[i] = milliseconds up to the end of the launch names_out[i] (static strings). */
void cadl_debug_force_generic(int on);
int cadl_debug_kernel_times(int enable, float* ms_out, const char** names_out, int cap);
#endif
/* Test hook: counts (into *mismatches_dev, a device uint64 the caller zeroed) the inputs with bit pattern in
 * [lo_bits, hi_bits] for which a device-math replica differs from the CUDA library form.
 *   which 0: log replica (scalar and packed fp32x2) vs logf;  which 1: Markstein a/param vs IEEE division;
 *   which 2: lg2.approx.ftz vs log2 in fp64: counts the inputs whose absolute error exceeds `param`;
 *   which 3: a/param through double precision (the gradient pass's path for divisors Markstein's scheme does not
 *   cover) vs IEEE division, any divisor. */
int cadl_selftest(int which, uint32_t lo_bits, uint32_t hi_bits, float param, unsigned long long* mismatches_dev,
                  cadl_stream_t stream);

/* Bytes of workspace for a (B,H,W) problem; 256-byte aligned pointer required.  The workspace must
 * be zeroed ONCE (cadl_workspace_init) and is left clean by every call. */
size_t cadl_workspace_bytes(int B, int H, int W);
int cadl_workspace_init(void* workspace, size_t bytes, cadl_stream_t stream);

/* Fused forward+backward of the requested loss terms (+ optional metrics):
 *   replaces CombinedDepthLoss::forwardWithIntrinsics + autograd backward
 *   (depth_loss.h:416-433; call site src/training/production_trainer.h:203-206), and, with
 *   terms = SI|GRAD|SMOOTH, CombinedDepthLoss::forward (depth_loss.h:390-404).
 * grad_pred (B,1,H,W) receives upstream * dL_total/dpred; may be NULL (forward only, e.g.
 * getComponents*, depth_loss.h:438-467).  rgb may be NULL without TERM_SMOOTH, K without
 * TERM_REPROJ. */
int cadl_stack_fwd_bwd(const float* pred, const float* gt, const float* rgb, const float* K,
                       const uint8_t* mask, int B, int H, int W, const cadl_params* params,
                       float* grad_pred, cadl_results* results, void* workspace,
                       size_t workspace_bytes, cadl_stream_t stream);

/* Split API, optional first call of a step: start the pooled-pyramid kernels of the gradient-matching term (they
 * need pred/gt and shape constants only) on an internal auxiliary stream forked from `stream`, so that they run
 * beside cadl_stack_reduce and the caller's all-reduce instead of in front of the gradient pass.  Returns
 * CADL_ERR_UNSUPPORTED when the streaming fast path does not apply (shape not a multiple of 8, unaligned pointers,
 * num_scales != 4, no gradient-matching term ...): then leave params->pyramid_prepared at 0.  On CADL_OK set
 * params->pyramid_prepared = 1 for the cadl_stack_reduce and cadl_stack_grad calls of THIS step. */
int cadl_stack_prepare(const float* pred, const float* gt, int B, int H, int W, const cadl_params* params,
                       void* workspace, size_t workspace_bytes, cadl_stream_t stream);

/* The same computation split at its one global dependency: phase A reduces the
 * batch-global scalars into a vector of `cadl_stats_count()` doubles at
 * (char*)workspace + cadl_stats_offset(); a multi-GPU caller all-reduces (sum) that vector between
 * the two calls and sets params->global_B; phase B writes gradients and results. */
int cadl_stack_reduce(const float* pred, const float* gt, const uint8_t* mask, int B, int H, int W,
                      const cadl_params* params, void* workspace, size_t workspace_bytes,
                      cadl_stream_t stream);
int cadl_stack_grad(const float* pred, const float* gt, const float* rgb, const float* K,
                    const uint8_t* mask, int B, int H, int W, const cadl_params* params,
                    float* grad_pred, cadl_results* results, void* workspace,
                    size_t workspace_bytes, cadl_stream_t stream);
size_t cadl_stats_offset(void);
int cadl_stats_count(void);

/* Single-term entry points (each = the class's forward + its autograd backward). */
/* ScaleInvariantLoss::forward, depth_loss.h:33-64 */
int cadl_si_fwd_bwd(const float* pred, const float* gt, const uint8_t* mask, int B, int H, int W,
                    float lambda, float eps, float upstream, float* grad_pred,
                    cadl_results* results, void* workspace, size_t workspace_bytes,
                    cadl_stream_t stream);
/* GradientMatchingLoss::forward, depth_loss.h:95-166 */
int cadl_gradmatch_fwd_bwd(const float* pred, const float* gt, int B, int H, int W, int num_scales,
                           float eps, float upstream, float* grad_pred, cadl_results* results,
                           void* workspace, size_t workspace_bytes, cadl_stream_t stream);
/* SmoothnessLoss::forward, depth_loss.h:189-234 */
int cadl_smooth_fwd_bwd(const float* pred, const float* rgb, int B, int H, int W, float eps,
                        float upstream, float* grad_pred, cadl_results* results, void* workspace,
                        size_t workspace_bytes, cadl_stream_t stream);
/* ReprojectionLoss::forward, depth_loss.h:268-331 */
int cadl_reproj_fwd_bwd(const float* pred, const float* gt, const float* K, int k_batched,
                        const uint8_t* mask, int B, int H, int W, float eps, float upstream,
                        float* grad_pred, cadl_results* results, void* workspace,
                        size_t workspace_bytes, cadl_stream_t stream);

/* autograd backward of the shim: grad_out[i] = grad_in[i] * (*upstream_dev).  A device scalar so no
 * host sync is needed; when *upstream_dev == 1.0f and grad_out == grad_in no memory is touched. */
int cadl_scale_grad(const float* grad_in, const float* upstream_dev, float* grad_out, size_t n,
                    cadl_stream_t stream);

/* Metrics only (no loss): DepthMetrics::compute (depth_metrics.h:40-88) and/or the trainers'
 * computeDepthMetrics (tensorboard_trainer_enhanced.h:400-439) over n = B*H*W values. */
int cadl_metrics(const float* pred, const float* gt, const uint8_t* mask, size_t n,
                 uint32_t which, float min_depth, float max_depth, cadl_results* results,
                 void* workspace, size_t workspace_bytes, cadl_stream_t stream);

/* RayDirectionComputer (src/preprocessing/ray_direction_computer.cpp):
 *   layout 0: out (B, H*W, 3) row-major -- computeRayDirections, :17-62
 *   layout 1: out (B, 3, H, W) planar   -- computeRayDirectionsMaps, :64-101 (the loader's layout,
 *             src/data/sunrgbd_loader.cpp:345-347)
 *   pose (B,4,4) row-major or NULL: rotate by R and re-normalise -- transformRaysToWorld, :103-127 */
int cadl_rays_from_K(const float* K, int k_batched, const float* pose, int B, int H, int W,
                     int layout, float* out, cadl_stream_t stream);

/* ---- "next" rows of the scope table: the steps either side of the loss path ---- */

/* SunRGBDLoader::resizeSample on the device (src/data/sunrgbd_loader.cpp:445-489): rgb (B,3,h,w) -> (B,3,H,W)
 * bilinear align_corners=false, depth (B,1,h,w) -> (B,1,H,W) nearest, K (B,3,3) rescaled (fx,cx by W/w; fy,cy by
 * H/h).  One launch after the H2D copy instead of per-sample host interpolate calls. */
int cadl_batch_prep(const float* rgb_in, const float* depth_in, const float* K_in, int B, int h, int w, int H, int W,
                    float* rgb_out, float* depth_out, float* K_out, cadl_stream_t stream);

/* The loader's augmentSample + the resize that follows it (src/data/sunrgbd_loader.cpp:352-384, :161-166) for a whole
 * batch in the same single launch.  aug_dev: (B, CADL_AUG_STRIDE) floats per image, drawn by the host RNG exactly as
 * the loader draws them:
 *   [0..3] crop_x, crop_y, crop_w, crop_h in input pixels (applyCrop :388-415; crop_w == 0: no crop)
 *   [4]    != 0: horizontal flip (applyHorizontalFlip :417-432)
 *   [5]    != 0: colour jitter clamp(rgb * [6] + [7] - 1, 0, 1) applied before the resize (applyColorJitter :434-443)
 * K follows: cx -= crop_x, cy -= crop_y; flip: cx = crop_w - cx - 1; then the resize scaling.  (The per-sample ray
 * maps the loader also crops/flips/resizes are regenerated from the final K with cadl_rays_from_K instead.) */
#define CADL_AUG_STRIDE 8
int cadl_batch_augment(const float* rgb_in, const float* depth_in, const float* K_in, const float* aug_dev, int B, int h,
                       int w, int H, int W, float* rgb_out, float* depth_out, float* K_out, cadl_stream_t stream);

/* ---- exact global-batch mode without a collective-library call (SURVEY.md 8e mode B) ----
 * The 32-double statistics vector is exchanged by one small kernel per rank over NVLink peer memory: each rank owns
 * an inbox (cadl_p2p_alloc: cudaMalloc + CUDA IPC handle), maps every peer's (cadl_p2p_open on the 64-byte handles
 * the ranks all-gather once, with any transport), and calls cadl_stats_exchange between cadl_stack_reduce and
 * cadl_stack_grad with an epoch counter that is the same on every rank and grows by one per call (start at 1).
 * The kernel pushes this rank's vector into every inbox, waits for the others (at most timeout_s seconds of device
 * time; <= 0 selects 600 s: rank skew of seconds is routine -- checkpointing, validation, loader stalls), and sums in rank
 * order -- the result is bit-identical on all ranks.  inboxes_host[r] = rank r's inbox as mapped in this process
 * (inboxes_host[rank] = the pointer cadl_p2p_alloc returned).  world <= 16, one node.  On a timeout the statistics
 * become NaN (the step's loss and gradient are then visibly invalid, never silently wrong) and a sticky flag is set:
 * cadl_p2p_error reads it (synchronises) and, with clear != 0, resets it. */
size_t cadl_p2p_inbox_bytes(int world);
int cadl_p2p_alloc(int world, void** inbox_dev, unsigned char handle_out[64]);
int cadl_p2p_open(const unsigned char handle[64], void** inbox_dev);
int cadl_p2p_close(void* inbox_dev, int own);
int cadl_stats_exchange(void* workspace, void* const* inboxes_host, int rank, int world, unsigned long long epoch,
                        double timeout_s, cadl_stream_t stream);
int cadl_p2p_error(void* own_inbox_dev, int world, int* error_host, int clear);

/* Device-resident running sums of per-batch scalars: acc_dev[i] += weight * values_dev[i] for i < n and
 * acc_dev[n] += weight (acc_dev: n + 1 doubles the caller zeroed).  Replaces the trainers'
 * `metrics.loss += loss.item<float>() * batch_size` (src/training/production_trainer.h:213-216,
 * tensorboard_trainer_enhanced.h:307,363) and its host sync per batch: read the sums once per epoch. */
int cadl_accumulate(const float* values_dev, int n, double weight, double* acc_dev, cadl_stream_t stream);

/* torch::nn::utils::clip_grad_norm_(params, max_norm) and the trainers' computeGradientNorm
 * (src/training/tensorboard_trainer_enhanced.h:300-302, :560-571) over `count` gradient tensors without a host
 * sync.  grad_ptrs_dev / sizes_dev: device arrays of pointers and element counts; chunk_prefix_dev: device
 * array of count+1 exclusive prefix sums of ceil(size/4096) (total_chunks = its last entry).  out2_dev receives
 * {total_norm, clip_coef}; with do_clip != 0 every gradient is multiplied by clip_coef = min(max_norm /
 * (total_norm + 1e-6), 1).  Workspace: cadl_clip_workspace_bytes(), zeroed once. */
size_t cadl_clip_workspace_bytes(void);
int cadl_clip_grad_norm(float* const* grad_ptrs_dev, const long long* sizes_dev, const long long* chunk_prefix_dev,
                        int count, long long total_chunks, float max_norm, float* out2_dev, void* workspace,
                        size_t workspace_bytes, int do_clip, cadl_stream_t stream);

/* Builder extension (NOT in the reference, whose forwardPhotometric is a stub returning zeros,
 * depth_loss.h:343-351): photometric reprojection  back-project -> [R|t] -> project -> bilinear
 * sample -> L1 residual, with explicit backward to depth.  T (B,4,4) row-major target->source.
 * results->loss_reproj receives the loss; parity is defined against oracle/oracle_torch.py only. */
int cadl_photometric_fwd_bwd(const float* pred, const float* K, int k_batched, const float* T,
                             const float* source, const float* target, int B, int H, int W,
                             float eps, float upstream, float* grad_pred, cadl_results* results,
                             void* workspace, size_t workspace_bytes, cadl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CADL_H_ */
