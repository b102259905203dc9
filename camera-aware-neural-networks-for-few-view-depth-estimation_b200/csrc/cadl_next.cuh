// cadl "next" rows of the scope table, the steps either side of the loss path:
//   * batch_prep_kernel  -- the loader's resizeSample on the device: rgb bilinear (align_corners = false), depth
//     nearest, K rescaled (reference src/data/sunrgbd_loader.cpp:445-489), fused in one launch after H2D;
//   * gradnorm / clip    -- torch::nn::utils::clip_grad_norm_ + the trainers' per-parameter .item() loop
//     (src/training/tensorboard_trainer_enhanced.h:300-302, :560-571) as two multi-tensor kernels with the
//     total norm left on the device (no host sync).
#pragma once
#include "cadl_common.cuh"
#include "cadl_phase_a.cuh"

namespace cadl {

// ------------------------------------------------------------------------------------------------
// resizeSample: at::upsample_bilinear2d (align_corners=false) for rgb, at::upsample_nearest2d for depth.
// Index arithmetic follows ATen (aten/src/ATen/native/UpSample.h): scale = in / out as float;
//   bilinear: src = scale * (dst + 0.5) - 0.5, clamped at 0; i0 = (int)src, i1 = min(i0 + 1, in - 1), l1 = src - i0
//   nearest : src = min((int)floorf(dst * scale), in - 1)
// ------------------------------------------------------------------------------------------------
struct PrepArgs {
    const float* rgb_in;    // (B,3,h,w)
    const float* depth_in;  // (B,1,h,w)
    const float* K_in;      // (B,3,3)
    float* rgb_out;         // (B,3,H,W)
    float* depth_out;       // (B,1,H,W)
    float* K_out;           // (B,3,3)
    int B, h, w, H, W;
    // optional per-image augmentation (B, CADL_AUG_STRIDE) floats, the loader's augmentSample
    // (src/data/sunrgbd_loader.cpp:352-384) composed with the resize that follows it (:161-166):
    //   [0..3] crop_x, crop_y, crop_w, crop_h  (applyCrop :388-415; crop_w == 0: no crop)
    //   [4]    horizontal flip != 0             (applyHorizontalFlip :417-432)
    //   [5]    colour jitter != 0, [6] contrast factor, [7] brightness factor   (applyColorJitter :434-443)
    const float* aug;
};

template <bool AUG>      // AUG = false: plain resizeSample, every augmentation branch compiled out
__global__ void __launch_bounds__(256) batch_prep_kernel(const PrepArgs a) {
    const int b = blockIdx.z;
    const int y = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    // the region of the input that is resized: the whole image, or the crop window
    int cx0 = 0, cy0 = 0, cw = a.w, ch = a.h;
    bool flip = false, jitter = false;
    float contrast = 1.f, brightness = 1.f;
    if constexpr (AUG) {
        const float* g = a.aug + (size_t)b * 8;
        if (__ldg(g + 2) > 0.f) {
            cx0 = (int)__ldg(g + 0); cy0 = (int)__ldg(g + 1); cw = (int)__ldg(g + 2); ch = (int)__ldg(g + 3);
            // the reference's Slice clamps the window to the tensor (sunrgbd_loader.cpp:395-397): an overshooting
            // window (crop_x can be 1 with crop_w == w at scale 1.0) must not read out of bounds
            cx0 = cx0 < 0 ? 0 : (cx0 > a.w - 1 ? a.w - 1 : cx0);
            cy0 = cy0 < 0 ? 0 : (cy0 > a.h - 1 ? a.h - 1 : cy0);
            cw = cw > a.w - cx0 ? a.w - cx0 : cw;
            ch = ch < 1 ? 1 : (ch > a.h - cy0 ? a.h - cy0 : ch);
        }
        flip = __ldg(g + 4) != 0.f;
        jitter = __ldg(g + 5) != 0.f;
        contrast = __ldg(g + 6);
        brightness = __ldg(g + 7);
    }
    if (x == 0 && y == 0) {
        const float* Ki = a.K_in + (size_t)b * 9;
        float* Ko = a.K_out + (size_t)b * 9;
        float k[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) k[i] = Ki[i];
        if (AUG && __ldg(a.aug + (size_t)b * 8 + 2) > 0.f) {   // :411-414  principal point follows the crop
            k[2] = k[2] - (float)cx0;
            k[5] = k[5] - (float)cy0;
        }
        if (flip) k[2] = ((float)cw - k[2]) - 1.0f;               // :428-431  W - cx - 1
        // sunrgbd_loader.cpp:480-488: fx, cx scale with W; fy, cy with H
        const float sx = (float)a.W / (float)cw, sy = (float)a.H / (float)ch;
        k[0] = k[0] * sx; k[4] = k[4] * sy; k[2] = k[2] * sx; k[5] = k[5] * sy;
#pragma unroll
        for (int i = 0; i < 9; ++i) Ko[i] = k[i];
    }
    if (x >= a.W) return;
    const float scale_h = (float)ch / (float)a.H, scale_w = (float)cw / (float)a.W;
    // column of the (cropped, flipped) image -> column of the input
    auto col = [&](int xc) { return AUG ? cx0 + (flip ? cw - 1 - xc : xc) : xc; };
    // nearest (depth): sunrgbd_loader.cpp:461-467
    {
        const int sy = min((int)floorf((float)y * scale_h), ch - 1);
        const int sx = min((int)floorf((float)x * scale_w), cw - 1);
        a.depth_out[((size_t)b * a.H + y) * a.W + x] = __ldg(a.depth_in + ((size_t)b * a.h + cy0 + sy) * a.w + col(sx));
    }
    // bilinear (rgb): sunrgbd_loader.cpp:453-459
    float fy = scale_h * ((float)y + 0.5f) - 0.5f;
    float fx = scale_w * ((float)x + 0.5f) - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < ch - 1 ? 1 : 0), x1 = x0 + (x0 < cw - 1 ? 1 : 0);
    const float ly1 = fy - (float)y0, lx1 = fx - (float)x0;
    const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    const int c0 = col(x0), c1 = col(x1);
    // colour jitter acts on the pixels before the resize: clamp(rgb * contrast + brightness - 1, 0, 1), one rounding per op
    auto tap = [&](const float* src, int yy, int xx) {
        float v = __ldg(src + (size_t)(cy0 + yy) * a.w + xx);
        if (AUG && jitter) {
            v = __fadd_rn(__fadd_rn(__fmul_rn(v, contrast), brightness), -1.0f);
            v = clamp_nan(v, 0.f, 1.f);
        }
        return v;
    };
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float* src = a.rgb_in + ((size_t)b * 3 + c) * a.h * a.w;
        // explicit roundings (no FMA contraction): both instantiations and the CPU restatement round alike
        const float top = __fadd_rn(__fmul_rn(lx0, tap(src, y0, c0)), __fmul_rn(lx1, tap(src, y0, c1)));
        const float bot = __fadd_rn(__fmul_rn(lx0, tap(src, y1, c0)), __fmul_rn(lx1, tap(src, y1, c1)));
        const float v = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
        a.rgb_out[(((size_t)b * 3 + c) * a.H + y) * a.W + x] = v;
    }
}

// acc[i] += weight * values[i] (i < n), acc[n] += weight: device-resident running sums of per-batch scalars
// (the trainers' `metrics.loss += loss.item<float>() * batch_size`, production_trainer.h:213-216, without the sync)
__global__ void accumulate_kernel(const float* values, int n, double weight, double* acc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) acc[i] += weight * (double)values[i];
    else if (i == n) acc[n] += weight;
}

inline cudaError_t launch_batch_prep(const PrepArgs& a, cudaStream_t st) {
    dim3 grid((a.W + 255) / 256, a.H, a.B);
    if (a.aug) batch_prep_kernel<true><<<grid, 256, 0, st>>>(a);
    else batch_prep_kernel<false><<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// clip_grad_norm_ over a list of tensors.  ptrs / sizes live on the device (built once per model).
//   total = sqrt(sum_t sum_i g_t[i]^2);  coef = min(max_norm / (total + 1e-6), 1);  g *= coef
// Kernel 1: grid-stride over (tensor, chunk) pairs, fp32 per thread -> fp64 per block -> last block.
// Kernel 2: scales every tensor by the device-resident coefficient (skipped per element when coef == 1).
// ------------------------------------------------------------------------------------------------
constexpr int kClipChunk = 4096;   // elements per (tensor, chunk) work item

struct ClipArgs {
    float* const* ptrs;
    const long long* sizes;
    const long long* chunk_prefix;   // exclusive prefix sum of ceil(size / kClipChunk), length count + 1
    int count;
    float max_norm;
    float* out;                       // [0] total norm, [1] clip coefficient
    WsHeader* hdr;
    double* part;
    int part_rows;
};

__device__ __forceinline__ int find_tensor(const long long* prefix, int count, long long item) {
    int lo = 0, hi = count;      // prefix[lo] <= item < prefix[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (prefix[mid] <= item) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) gradnorm_kernel(const ClipArgs a) {
    __shared__ double s_d[8];
    __shared__ int s_last;
    const long long items = a.chunk_prefix[a.count];
    float acc = 0.f;
    for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        const int t = find_tensor(a.chunk_prefix, a.count, it);
        const long long c0 = (it - a.chunk_prefix[t]) * kClipChunk;
        const long long n = a.sizes[t];
        const float* g = a.ptrs[t];
        const long long c1 = c0 + kClipChunk < n ? c0 + kClipChunk : n;
        for (long long i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
            const float v = __ldg(g + i);
            acc = fmaf(v, v, acc);
        }
    }
    const double r = block_sum_double((double)acc, s_d);
    if (threadIdx.x == 0) a.part[blockIdx.x] = r;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&a.hdr->ticket_b, 1u);
        s_last = (t == gridDim.x - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const volatile double* part = a.part;
    double s = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += part[i];
    s = block_sum_double(s, s_d);
    if (threadIdx.x == 0) {
        const float total = (float)sqrt(s);
        float coef = a.max_norm / (total + 1e-6f);     // torch/nn/utils/clip_grad.h: clip_coef = max_norm / (total_norm + 1e-6)
        coef = coef > 1.0f ? 1.0f : coef;              // clamped to 1.0
        a.out[0] = total;
        a.out[1] = coef;
        a.hdr->ticket_b = 0u;
    }
}

__global__ void __launch_bounds__(256) gradscale_kernel(const ClipArgs a) {
    const float coef = a.out[1];
    if (coef == 1.0f) return;      // x * 1 == x: nothing to write
    const long long items = a.chunk_prefix[a.count];
    for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        const int t = find_tensor(a.chunk_prefix, a.count, it);
        const long long c0 = (it - a.chunk_prefix[t]) * kClipChunk;
        const long long n = a.sizes[t];
        float* g = a.ptrs[t];
        const long long c1 = c0 + kClipChunk < n ? c0 + kClipChunk : n;
        for (long long i = c0 + threadIdx.x; i < c1; i += blockDim.x) g[i] *= coef;
    }
}

// ------------------------------------------------------------------------------------------------
// Exact global-batch mode (SURVEY 8e mode B) without a collective library call: the 32-double statistics vector is
// exchanged by ONE small kernel per rank over NVLink peer memory.  Every rank owns an "inbox" (cudaMalloc + CUDA IPC,
// mapped by all peers): 2 parities x world slots of {32 doubles, epoch flag}.  The kernel
//   1. pushes this rank's vector into slot [epoch & 1][rank] of every peer's inbox (plain stores over NVLink),
//      fences (system scope) and then publishes the epoch flag of each slot;
//   2. waits until all `world` flags of its OWN inbox carry this epoch;
//   3. sums the slots in rank order (fixed order: every rank computes bit-identical sums) into its statistics vector.
// Two parities suffice: nobody can start epoch e+1 before it has read epoch e, and epoch e+2 needs everybody's e+1.
// The wait is bounded (timeout_ns of globaltimer, chosen by the caller: rank skew of seconds is routine in training --
// checkpointing, validation, loader stalls): on timeout the kernel does not hang the GPU; it sets a sticky error flag
// and writes NaN into the statistics, so the step's loss and gradient are visibly invalid instead of silently wrong.
// ------------------------------------------------------------------------------------------------
constexpr int kP2PSlotDoubles = ST_COUNT + 2;     // 32 values, flag, pad
constexpr int kP2PMaxWorld = 16;

struct P2PArgs {
    double* stats;                      // this rank's statistics vector (in/out)
    double* inbox[kP2PMaxWorld];        // inbox[r]: rank r's inbox as mapped in THIS process (inbox[rank] = own)
    int rank, world;
    unsigned long long epoch;           // > 0, the same on every rank, incremented per exchange
    int* error;                         // set to 1 on timeout (sticky until cadl_p2p_error clears it)
    unsigned long long timeout_ns;
};

__global__ void __launch_bounds__(64) stats_exchange_kernel(const P2PArgs a) {
    __shared__ int s_timeout;
    const int t = threadIdx.x;
    if (t == 0) s_timeout = 0;
    pdl_wait();        // launched with programmatic stream serialization behind phase A: its statistics
    const size_t slot = ((size_t)(a.epoch & 1ull) * a.world + a.rank) * kP2PSlotDoubles;
    if (t < ST_COUNT) {
        const double v = a.stats[t];
        for (int p = 0; p < a.world; ++p) reinterpret_cast<volatile double*>(a.inbox[p] + slot)[t] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (t < a.world) {
        unsigned long long* flag = reinterpret_cast<unsigned long long*>(a.inbox[t] + slot + ST_COUNT);
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(flag), "l"(a.epoch) : "memory");
        // wait for rank t's contribution in the own inbox
        const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(
            a.inbox[a.rank] + ((size_t)(a.epoch & 1ull) * a.world + t) * kP2PSlotDoubles + ST_COUNT);
        unsigned long long t0, now, seen = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
            if (seen == a.epoch) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > a.timeout_ns) { *a.error = 1; s_timeout = 1; break; }
            __nanosleep(200);
        }
    }
    __syncthreads();
    if (t < ST_COUNT) {
        double sum = 0.0;
        for (int r = 0; r < a.world; ++r)
            sum += reinterpret_cast<const volatile double*>(a.inbox[a.rank] + ((size_t)(a.epoch & 1ull) * a.world + r) * kP2PSlotDoubles)[t];
        a.stats[t] = s_timeout ? __longlong_as_double(0x7ff8000000000000ll) : sum;      // a stale slot must not pass as a statistic
    }
}

}  // namespace cadl
