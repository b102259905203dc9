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
};

__global__ void __launch_bounds__(256) batch_prep_kernel(const PrepArgs a) {
    const int b = blockIdx.z;
    const int y = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x == 0 && y == 0) {
        // sunrgbd_loader.cpp:480-488: fx, cx scale with W; fy, cy with H
        const float sx = (float)a.W / (float)a.w, sy = (float)a.H / (float)a.h;
        const float* Ki = a.K_in + (size_t)b * 9;
        float* Ko = a.K_out + (size_t)b * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) Ko[i] = Ki[i];
        Ko[0] = Ki[0] * sx; Ko[4] = Ki[4] * sy; Ko[2] = Ki[2] * sx; Ko[5] = Ki[5] * sy;
    }
    if (x >= a.W) return;
    const float scale_h = (float)a.h / (float)a.H, scale_w = (float)a.w / (float)a.W;
    // nearest (depth): sunrgbd_loader.cpp:461-467
    {
        const int sy = min((int)floorf((float)y * scale_h), a.h - 1);
        const int sx = min((int)floorf((float)x * scale_w), a.w - 1);
        a.depth_out[((size_t)b * a.H + y) * a.W + x] = __ldg(a.depth_in + ((size_t)b * a.h + sy) * a.w + sx);
    }
    // bilinear (rgb): sunrgbd_loader.cpp:453-459
    float fy = scale_h * ((float)y + 0.5f) - 0.5f;
    float fx = scale_w * ((float)x + 0.5f) - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < a.h - 1 ? 1 : 0), x1 = x0 + (x0 < a.w - 1 ? 1 : 0);
    const float ly1 = fy - (float)y0, lx1 = fx - (float)x0;
    const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float* src = a.rgb_in + ((size_t)b * 3 + c) * a.h * a.w;
        const float v = ly0 * (lx0 * __ldg(src + (size_t)y0 * a.w + x0) + lx1 * __ldg(src + (size_t)y0 * a.w + x1)) +
                        ly1 * (lx0 * __ldg(src + (size_t)y1 * a.w + x0) + lx1 * __ldg(src + (size_t)y1 * a.w + x1));
        a.rgb_out[(((size_t)b * 3 + c) * a.H + y) * a.W + x] = v;
    }
}

inline cudaError_t launch_batch_prep(const PrepArgs& a, cudaStream_t st) {
    dim3 grid((a.W + 255) / 256, a.H, a.B);
    batch_prep_kernel<<<grid, 256, 0, st>>>(a);
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

}  // namespace cadl
