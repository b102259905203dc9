// cadl phase B, streaming form of the fast path (aligned shapes): the two kernels that turn the three pooled scales of
// the gradient-matching term into ONE field before the full-resolution pass runs.
//
// The gradient-matching normalisers are shape constants (depth_loss.h:162-163), so what the three pooled scales add
// to dL/dpred depends on pred/gt only:
//
//   pyr_pool_kernel   one thread per 8x8 block, marching its 8 rows once: avg-pool sums of scales 1..3 in ATen's
//                     row-major order (running sums, no halo, no shared memory), log(clamp(.)) and 1/q per cell
//                     -> workspace arrays LP_s, LG_s, RQ_s                                 (8 B/px read, 3.9 written)
//   pyr_coef_kernel   one thread per 8x8 block: the four signed edge residuals of every pooled cell at scales 3, 2, 1,
//                     the coarse coefficients gathered down to scale 1 -> C1 (B, H/2, W/2): what each pixel of a
//                     2x2 cell adds to its gradient; loss sums of the three coarse scales
//
// The full-resolution pass that consumes C1 is stream3_kernel (cadl_stream3.cuh).
#pragma once
#include "cadl_common.cuh"
#include "cadl_math.cuh"
#include "cadl_phase_b.cuh"
#include "cadl_phase_b_fast.cuh"

namespace cadl {

struct PyrArrays {
    float* lp[3];   // log(clamp(avg_pool_s(pred)))   s = 1..3   (B, H>>s, W>>s)
    float* lg[3];   // log(clamp(avg_pool_s(gt)))
    float* rq[3];   // 1/avg_pool_s(pred) inside the clamp range, else 0 (clamp backward)
    float* c1;      // (B, H/2, W/2): coarse-scale gradient each full-resolution pixel of the cell receives
};

// ================================================================================================
// pyr_pool_kernel
// ================================================================================================
#ifndef CADL_PYR_MINB
#define CADL_PYR_MINB 4
#endif
__global__ void __launch_bounds__(256, CADL_PYR_MINB) pyr_pool_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                       int B, int H, int W, float eps, PyrArrays py,
                                                       unsigned int* img_rec_words) {
    // Launched with programmatic stream serialization behind phase A, of which it needs nothing: it starts as
    // phase A's CTAs drain.  It only has to END after phase A (pdl_wait below), so that the kernels behind it,
    // which wait for THIS grid, also see phase A's statistics.
    pdl_trigger();
    const int W8 = W >> 3, H8 = H >> 3;
    // the streaming kernel's per-image records (stream-ordered before it; their offset depends on the shape, and one
    // workspace serves calls of different shapes: what lies there may be another shape's partial sums)
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < B * 32; i += blockDim.x) img_rec_words[i] = 0u;      // 128 bytes per image
    // grid-stride over the 8x8 blocks: the grid is either one thread per block or, when the kernel runs beside
    // phase A on a second stream, one CTA per SM
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < B * H8 * W8; idx += gridDim.x * blockDim.x) {
    const int bx = idx % W8, by = (idx / W8) % H8, b = idx / (W8 * H8);
    const float* pp = pred + (size_t)b * H * W + (size_t)(by * 8) * W + bx * 8;
    const float* gp = gt + (size_t)b * H * W + (size_t)(by * 8) * W + bx * 8;
    const int W1 = W >> 1, W2 = W >> 2, W3 = W >> 3, H1 = H >> 1, H2 = H >> 2, H3 = H >> 3;

    float s1p[4], s1g[4], s2p[2], s2g[2], s3p = 0.f, s3g = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const float4 pa = __ldg(reinterpret_cast<const float4*>(pp + (size_t)r * W));
        const float4 pb = __ldg(reinterpret_cast<const float4*>(pp + (size_t)r * W + 4));
        const float4 ga = __ldg(reinterpret_cast<const float4*>(gp + (size_t)r * W));
        const float4 gb = __ldg(reinterpret_cast<const float4*>(gp + (size_t)r * W + 4));
        const float vp[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
        const float vg[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
        // running window sums, row-major sequential inside each window (ATen avg_pool2d order)
        if ((r & 1) == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) { s1p[c] = 0.f; s1g[c] = 0.f; }
        }
        if ((r & 3) == 0) { s2p[0] = s2p[1] = 0.f; s2g[0] = s2g[1] = 0.f; }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            s1p[c >> 1] += vp[c]; s1g[c >> 1] += vg[c];
            s2p[c >> 2] += vp[c]; s2g[c >> 2] += vg[c];
            s3p += vp[c]; s3g += vg[c];
        }
        if (r & 1) {   // a row of 4 scale-1 cells is complete
            float qp[4], qg[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) { qp[c] = s1p[c] * 0.25f; qg[c] = s1g[c] * 0.25f; }
            const float2 a0 = log_exact2(make_float2(clamp_nan(qp[0], eps, 1000.0f), clamp_nan(qp[1], eps, 1000.0f)));
            const float2 a1 = log_exact2(make_float2(clamp_nan(qp[2], eps, 1000.0f), clamp_nan(qp[3], eps, 1000.0f)));
            const float2 b0 = log_exact2(make_float2(clamp_nan(qg[0], eps, 1000.0f), clamp_nan(qg[1], eps, 1000.0f)));
            const float2 b1 = log_exact2(make_float2(clamp_nan(qg[2], eps, 1000.0f), clamp_nan(qg[3], eps, 1000.0f)));
            float rq[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) rq[c] = in_range_pos(qp[c], eps, 1000.0f) ? rcp_approx(qp[c]) : 0.f;
            const size_t o = ((size_t)b * H1 + (by * 4 + (r >> 1))) * W1 + bx * 4;
            *reinterpret_cast<float4*>(py.lp[0] + o) = make_float4(a0.x, a0.y, a1.x, a1.y);
            *reinterpret_cast<float4*>(py.lg[0] + o) = make_float4(b0.x, b0.y, b1.x, b1.y);
            *reinterpret_cast<float4*>(py.rq[0] + o) = make_float4(rq[0], rq[1], rq[2], rq[3]);
        }
        if ((r & 3) == 3) {   // a row of 2 scale-2 cells
            const float q0 = s2p[0] * 0.0625f, q1 = s2p[1] * 0.0625f, g0 = s2g[0] * 0.0625f, g1 = s2g[1] * 0.0625f;
            const float2 a = log_exact2(make_float2(clamp_nan(q0, eps, 1000.0f), clamp_nan(q1, eps, 1000.0f)));
            const float2 c = log_exact2(make_float2(clamp_nan(g0, eps, 1000.0f), clamp_nan(g1, eps, 1000.0f)));
            const size_t o = ((size_t)b * H2 + (by * 2 + (r >> 2))) * W2 + bx * 2;
            *reinterpret_cast<float2*>(py.lp[1] + o) = a;
            *reinterpret_cast<float2*>(py.lg[1] + o) = c;
            *reinterpret_cast<float2*>(py.rq[1] + o) = make_float2(in_range_pos(q0, eps, 1000.0f) ? rcp_approx(q0) : 0.f,
                                                                   in_range_pos(q1, eps, 1000.0f) ? rcp_approx(q1) : 0.f);
        }
    }
    {
        const float q = s3p * 0.015625f, g = s3g * 0.015625f;
        const float2 l = log_exact2(make_float2(clamp_nan(q, eps, 1000.0f), clamp_nan(g, eps, 1000.0f)));
        const size_t o = ((size_t)b * H3 + by) * W3 + bx;
        py.lp[2][o] = l.x;
        py.lg[2][o] = l.y;
        py.rq[2][o] = in_range_pos(q, eps, 1000.0f) ? rcp_approx(q) : 0.f;
    }
    }   // blocks
    pdl_wait();
}

// ================================================================================================
// pyr_coef_kernel
// ================================================================================================
// One scale, the N x N cells of this thread's block: coefficient = (d loss_s / d log q) * (1/q) * spread.
// Neighbours outside the image are read with a clamped index, so an edge across the border has residual exactly 0.
template <int N>
__device__ __forceinline__ void coef_level(const float* __restrict__ LP, const float* __restrict__ LG,
                                           const float* __restrict__ RQ, int Hs, int Ws, int cy0, int cx0,
                                           float inv_nx, float inv_ny, float (&coef)[N][N], float& acc_x, float& acc_y) {
    float lp[N + 2][N + 2], lg[N + 2][N + 2];
    const int xl = cx0 > 0 ? cx0 - 1 : 0, xr = cx0 + N < Ws ? cx0 + N : Ws - 1;
#pragma unroll
    for (int i = 0; i < N + 2; ++i) {
        int cy = cy0 + i - 1;
        cy = cy < 0 ? 0 : (cy >= Hs ? Hs - 1 : cy);
        const float* rp = LP + (size_t)cy * Ws;
        const float* rg = LG + (size_t)cy * Ws;
        if constexpr (N == 4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(rp + cx0)), c = __ldg(reinterpret_cast<const float4*>(rg + cx0));
            lp[i][1] = a.x; lp[i][2] = a.y; lp[i][3] = a.z; lp[i][4] = a.w;
            lg[i][1] = c.x; lg[i][2] = c.y; lg[i][3] = c.z; lg[i][4] = c.w;
        } else if constexpr (N == 2) {
            const float2 a = __ldg(reinterpret_cast<const float2*>(rp + cx0)), c = __ldg(reinterpret_cast<const float2*>(rg + cx0));
            lp[i][1] = a.x; lp[i][2] = a.y;
            lg[i][1] = c.x; lg[i][2] = c.y;
        } else {
            lp[i][1] = __ldg(rp + cx0);
            lg[i][1] = __ldg(rg + cx0);
        }
        if (i >= 1 && i <= N) {   // the corners are never used
            lp[i][0] = __ldg(rp + xl); lg[i][0] = __ldg(rg + xl);
            lp[i][N + 1] = __ldg(rp + xr); lg[i][N + 1] = __ldg(rg + xr);
        } else {
            lp[i][0] = lp[i][N + 1] = 0.f; lg[i][0] = lg[i][N + 1] = 0.f;
        }
    }
    // signed x-edges ex[i][j]: between columns j-1 and j of row i (j = 0..N); y-edges ey[i][j]: rows i-1 and i
    float sx[N][N + 1], sy[N + 1][N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j <= N; ++j) {
            const float e = (lp[i + 1][j + 1] - lp[i + 1][j]) - (lg[i + 1][j + 1] - lg[i + 1][j]);   // depth_loss.h:140-148,162
            sx[i][j] = sgn3(e);
            if (j >= 1) acc_x += fabsf(e);          // the edge to the right of an own cell
        }
#pragma unroll
    for (int i = 0; i <= N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const float e = (lp[i + 1][j + 1] - lp[i][j + 1]) - (lg[i + 1][j + 1] - lg[i][j + 1]);   // :151-159,163
            sy[i][j] = sgn3(e);
            if (i >= 1) acc_y += fabsf(e);          // the edge below an own cell
        }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        float rq[N];
        const float* rr = RQ + (size_t)(cy0 + i) * Ws + cx0;
        if constexpr (N == 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(rr));
            rq[0] = v.x; rq[1] = v.y; rq[2] = v.z; rq[3] = v.w;
        } else if constexpr (N == 2) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(rr));
            rq[0] = v.x; rq[1] = v.y;
        } else {
            rq[0] = __ldg(rr);
        }
#pragma unroll
        for (int j = 0; j < N; ++j)
            coef[i][j] = ((sx[i][j] - sx[i][j + 1]) * inv_nx + (sy[i][j] - sy[i + 1][j]) * inv_ny) * rq[j];
    }
}

struct PyrCoefArgs {
    PyrArrays py;
    int B, H, W;
    float inv_nx[4], inv_ny[4];
    float wg;           // w_grad * upstream / num_scales
    double* b_part;     // partial rows (BF_COUNT doubles each)
    int row0;           // first row this kernel writes
};

__global__ void __launch_bounds__(256, CADL_PYR_MINB) pyr_coef_kernel(const PyrCoefArgs a) {
    __shared__ float s_f[8][6];
    pdl_wait();        // the pooled arrays of pyr_pool_kernel
    pdl_trigger();
    const int W8 = a.W >> 3, H8 = a.H >> 3;
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // GX1, GY1, GX2, GY2, GX3, GY3
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < a.B * H8 * W8; idx += gridDim.x * blockDim.x) {
        const int bx = idx % W8, by = (idx / W8) % H8, b = idx / (W8 * H8);
        const int H1 = a.H >> 1, W1 = a.W >> 1, H2 = a.H >> 2, W2 = a.W >> 2, H3 = a.H >> 3, W3 = a.W >> 3;
        float c3[1][1], c2[2][2], c1[4][4];
        {
            const size_t o = (size_t)b * H3 * W3;
            const float sp = a.wg * (1.0f / 64.0f);
            coef_level<1>(a.py.lp[2] + o, a.py.lg[2] + o, a.py.rq[2] + o, H3, W3, by, bx, a.inv_nx[3] * sp, a.inv_ny[3] * sp,
                          c3, acc[4], acc[5]);
        }
        {
            const size_t o = (size_t)b * H2 * W2;
            const float sp = a.wg * (1.0f / 16.0f);
            coef_level<2>(a.py.lp[1] + o, a.py.lg[1] + o, a.py.rq[1] + o, H2, W2, by * 2, bx * 2, a.inv_nx[2] * sp,
                          a.inv_ny[2] * sp, c2, acc[2], acc[3]);
        }
        {
            const size_t o = (size_t)b * H1 * W1;
            const float sp = a.wg * 0.25f;
            coef_level<4>(a.py.lp[0] + o, a.py.lg[0] + o, a.py.rq[0] + o, H1, W1, by * 4, bx * 4, a.inv_nx[1] * sp,
                          a.inv_ny[1] * sp, c1, acc[0], acc[1]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 v;
            v.x = c1[i][0] + (c2[i >> 1][0] + c3[0][0]);
            v.y = c1[i][1] + (c2[i >> 1][0] + c3[0][0]);
            v.z = c1[i][2] + (c2[i >> 1][1] + c3[0][0]);
            v.w = c1[i][3] + (c2[i >> 1][1] + c3[0][0]);
            *reinterpret_cast<float4*>(a.py.c1 + ((size_t)b * H1 + by * 4 + i) * W1 + bx * 4) = v;
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const float v = warp_sum(acc[q]);
        if (lane == 0) s_f[warp][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < BF_COUNT) {
        const int q = threadIdx.x;
        double t = 0.0;
        if (q >= BF_GX1 && q <= BF_GY3)
            for (int w = 0; w < 8; ++w) t += (double)s_f[w][q - BF_GX1];
        a.b_part[(size_t)(a.row0 + blockIdx.x) * BF_COUNT + q] = t;
    }
}

}  // namespace cadl
