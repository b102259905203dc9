// cadl rays: K -> per-pixel unit ray directions on the device
// (reference: src/preprocessing/ray_direction_computer.cpp:17-62, :64-101, :103-127).
// Store-bound: 12 B/px written, nothing read but 9 (+16) floats per image.
#pragma once
#include "cadl_common.cuh"

namespace cadl {

// One thread = four adjacent pixels of a row: the planar layout gets three 128-bit stores (one per component), the
// (H*W, 3) layout three 128-bit stores of the thread's twelve consecutive floats (48 contiguous bytes per thread, the
// warp's 1536 bytes contiguous: every 32-byte sector is written in full) -- round 1 stored 4 bytes at stride 12.
// Arithmetic follows the reference statement by statement; __f*_rn keeps nvcc from contracting x*x + y*y + z*z into
// FMAs so the result is the un-contracted C value.  SCALAR = 1: any width / unaligned output, one pixel per thread.
template <int SCALAR>
__global__ void __launch_bounds__(256) rays_kernel(const float* __restrict__ K, int k_batched,
                                                   const float* __restrict__ pose, int H, int W, int layout,
                                                   float* __restrict__ out) {
    constexpr int PX = SCALAR ? 1 : 4;
    const int b = blockIdx.z;
    const int v = blockIdx.y;
    const int u0 = (blockIdx.x * blockDim.x + threadIdx.x) * PX;
    if (u0 >= W) return;
    const float* Kb = K + (k_batched ? (size_t)b * 9 : 0);
    const float fx = __ldg(Kb + 0), cx = __ldg(Kb + 2), fy = __ldg(Kb + 4), cy = __ldg(Kb + 5);   // :29-32
    const float fx_inv = __fdiv_rn(1.0f, fx);                                                     // :35
    const float fy_inv = __fdiv_rn(1.0f, fy);                                                     // :36
    const float y = __fmul_rn(__fsub_rn((float)v, cy), fy_inv);                                   // :48
    float P[12];
    if (pose) {
#pragma unroll
        for (int i = 0; i < 12; ++i) P[i] = __ldg(pose + (size_t)b * 16 + i);   // row-major 4x4; R = top-left 3x3 (:108)
    }
    float r[PX][3];
#pragma unroll
    for (int k = 0; k < PX; ++k) {
        const float x = __fmul_rn(__fsub_rn((float)(u0 + k), cx), fx_inv);                        // :47
        const float z = 1.0f;                                                                     // :49
        const float norm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));  // :52
        float r0 = __fdiv_rn(x, norm), r1 = __fdiv_rn(y, norm), r2 = __fdiv_rn(z, norm);           // :53-55
        if (pose) {
            float w0 = __fadd_rn(__fadd_rn(__fmul_rn(P[0], r0), __fmul_rn(P[1], r1)), __fmul_rn(P[2], r2));
            float w1 = __fadd_rn(__fadd_rn(__fmul_rn(P[4], r0), __fmul_rn(P[5], r1)), __fmul_rn(P[6], r2));
            float w2 = __fadd_rn(__fadd_rn(__fmul_rn(P[8], r0), __fmul_rn(P[9], r1)), __fmul_rn(P[10], r2));
            float n2 = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(w0, w0), __fmul_rn(w1, w1)), __fmul_rn(w2, w2)));
            if (n2 > 0.0f) { w0 = __fdiv_rn(w0, n2); w1 = __fdiv_rn(w1, n2); w2 = __fdiv_rn(w2, n2); }   // :119
            r0 = w0; r1 = w1; r2 = w2;
        }
        r[k][0] = r0; r[k][1] = r1; r[k][2] = r2;
    }
    const size_t hw = (size_t)H * W;
    const size_t i = (size_t)v * W + u0;
    if constexpr (SCALAR) {
        if (layout == 0) {
            float* o = out + ((size_t)b * hw + i) * 3;   // (B, H*W, 3)
            o[0] = r[0][0]; o[1] = r[0][1]; o[2] = r[0][2];
        } else {
            float* o = out + (size_t)b * 3 * hw + i;     // (B, 3, H, W)
            o[0] = r[0][0]; o[hw] = r[0][1]; o[2 * hw] = r[0][2];
        }
    } else {
        if (layout == 0) {
            float4* o = reinterpret_cast<float4*>(out + ((size_t)b * hw + i) * 3);   // 12 consecutive floats
            o[0] = make_float4(r[0][0], r[0][1], r[0][2], r[1][0]);
            o[1] = make_float4(r[1][1], r[1][2], r[2][0], r[2][1]);
            o[2] = make_float4(r[2][2], r[3][0], r[3][1], r[3][2]);
        } else {
            float* o = out + (size_t)b * 3 * hw + i;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                *reinterpret_cast<float4*>(o + (size_t)c * hw) = make_float4(r[0][c], r[1][c], r[2][c], r[3][c]);
        }
    }
}

inline cudaError_t launch_rays(const float* K, int k_batched, const float* pose, int B, int H, int W, int layout,
                               float* out, cudaStream_t st) {
    const bool vec = (W % 4 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
    if (vec) {
        dim3 grid((W / 4 + 255) / 256, H, B);
        rays_kernel<0><<<grid, 256, 0, st>>>(K, k_batched, pose, H, W, layout, out);
    } else {
        dim3 grid((W + 255) / 256, H, B);
        rays_kernel<1><<<grid, 256, 0, st>>>(K, k_batched, pose, H, W, layout, out);
    }
    return cudaGetLastError();
}

}  // namespace cadl
