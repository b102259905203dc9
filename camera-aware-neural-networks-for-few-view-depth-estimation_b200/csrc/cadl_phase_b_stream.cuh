// cadl phase B, streaming form of the fast path (aligned shapes, same dispatch condition as cadl_phase_b_fast.cuh).
//
// The tile kernel in cadl_phase_b_fast.cuh does everything for one 48x128 tile inside one CTA: it stages pred/gt
// with an 8-pixel halo, builds the avg-pool pyramid in shared memory, and only then runs the full-resolution
// pass.  ncu shows what that costs: ~36 % of its instructions are the prelude, its phases are separated by six
// CTA-wide barriers (top stall reason), and 101 KB of shared memory caps the SM at 16 warps.  The coarse scales
// do not need any of that to be fused: the gradient-matching normalisers are shape constants (depth_loss.h:162-163),
// so what the three pooled scales add to dL/dpred is a field that depends on pred/gt only.  This file splits the
// work accordingly:
//
//   pyr_pool_kernel   one thread per 8x8 block, marching its 8 rows once: avg-pool sums of scales 1..3 in ATen's
//                     row-major order (running sums, no halo, no shared memory), log(clamp(.)) and 1/q per cell
//                     -> workspace arrays LP_s, LG_s, RQ_s                                 (8 B/px read, 3.9 written)
//   pyr_coef_kernel   one thread per 8x8 block: the four signed edge residuals of every pooled cell at scales 3, 2, 1,
//                     the coarse coefficients gathered down to scale 1 -> C1 (B, H/2, W/2): what each pixel of a
//                     2x2 cell adds to its gradient; loss sums of the three coarse scales
//   phase_b_stream_kernel   the full-resolution pass alone: one warp = 128 columns marching down ~32 rows, logs
//                     evaluated in registers as each row arrives, every edge once, no shared memory, no barriers,
//                     no halo except one pixel per warp end.  Each image's strip-rows are divided evenly over the
//                     warps of ONE resident wave (2 CTAs x 6 warps per SM), so there is no tail of partial waves
//                     (claiming 8-row chunks dynamically was measured slower: +10 % instructions for the extra
//                     prologues and a tail of up to one chunk out of four).  The warp that finishes an image folds
//                     that image's partial rows into one, so the kernel-final reduction reads B rows, not thousands
//                     (a single CTA reading one row per warp was a 30 us serial tail).
//
// Values are the same as the tile kernel's (same operations on the sign-critical paths); tests compare the two.
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
__global__ void __launch_bounds__(256) pyr_pool_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                       int B, int H, int W, float eps, PyrArrays py,
                                                       unsigned int* img_cnt) {
    // Launched with programmatic stream serialization behind phase A, of which it needs nothing: it starts as
    // phase A's CTAs drain.  It only has to END after phase A (pdl_wait below), so that the kernels behind it,
    // which wait for THIS grid, also see phase A's statistics.
    pdl_trigger();
    const int W8 = W >> 3, H8 = H >> 3;
    // the streaming kernel's per-image completion counters (stream-ordered before it; their offset depends on the
    // shape, and one workspace serves calls of different shapes)
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < B; i += blockDim.x) img_cnt[i] = 0u;
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

__global__ void __launch_bounds__(256) pyr_coef_kernel(const PyrCoefArgs a) {
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

// ================================================================================================
// phase_b_stream_kernel: the full-resolution pass
// ================================================================================================
struct StreamArgs {
    const float* c1;     // PyrArrays::c1
    int nstrip;          // 128-column strips per image row
    int cpi;             // shares per image (partial rows of an image are contiguous)
    double* chunk_part;  // one partial row per share; a.b_part holds [B per-image rows][pyr rows]
    unsigned int* img_cnt;   // shares done per image (zeroed by pyr_pool_kernel)
    int finalize_inline; // 1: the last CTA reduces and writes cadl_results; 0: stream_finish_kernel does (smoothness + gradient)
    unsigned long long* trace;   // cadl_debug_set_trace: per warp {smid, start ns, end ns, items}; null = off
    int trace_cap;
};
// L2 prefetch distance in rows, ahead of the one-row register prefetch (us/step at config 3: 1 -> 196, 2 -> 191.5,
// 3 -> 193.1, 5 -> 192.8, 8 -> 194.2)
#ifndef CADL_STREAM_PF
#define CADL_STREAM_PF 2
#endif
constexpr int kStreamPrefetchRows = CADL_STREAM_PF;     // L2 prefetch distance inside a chunk

// Sum of one image's chunk rows in a fixed order (lane-strided, fixed shuffle tree), by the warp that finished the image.
__device__ __noinline__ void fold_image_rows(const double* rows, int n, double* out, unsigned int* cnt, int lane) {
    __threadfence();
    double t[BF_COUNT];
#pragma unroll
    for (int q = 0; q < BF_COUNT; ++q) t[q] = 0.0;
#pragma unroll 2
    for (int i = lane; i < n; i += 32) {
#pragma unroll
        for (int q = 0; q < BF_COUNT; ++q) t[q] += __ldcg(rows + (size_t)i * BF_COUNT + q);
    }
    double mine = 0.0;
#pragma unroll
    for (int q = 0; q < BF_COUNT; ++q) {
        const double r = warp_sum(t[q]);
        mine = (lane == q) ? r : mine;
    }
    if (lane < BF_COUNT) out[lane] = mine;
    if (lane == 0) *cnt = 0u;
}

// One image row as a lane holds it: its own 4 pixels, the right neighbour, and the end lanes' halo pixel.
struct StreamRow {
    float p[5], g[4], I[3][5];     // pred (+ right neighbour), gt, rgb (+ right neighbour)
    float lp[4], lg[4];            // log(clamp(pred)), log(clamp(gt))          depth_loss.h:115-116
    float hp, hg, hI[3];           // halo pixel: lane 0 its left neighbour, the last lane its right neighbour
    float hlp, hlg;                // logs of the halo pixel
};

template <int F, bool HAS_MASK>
__global__ void __launch_bounds__(kStreamThreads, kStreamCtasPerSm) phase_b_stream_kernel(const PhaseBArgs a, const StreamArgs sa) {
    __shared__ double s_d[8];
    __shared__ int s_last;
    constexpr bool SMOOTH = (F & FB_SMOOTH) != 0;
    constexpr bool SI = (F & FB_SI) != 0;
    constexpr bool RP = (F & FB_RP) != 0;
    static_assert((F & FB_GRAD) != 0, "the streaming kernel is the gradient-matching path");
    const int tid = threadIdx.x, lane = tid & 31;
    const int H = a.H, W = a.W;
    const int plane = H * W, W1 = W >> 1;                    // 3*H*W < 2^31 (checked on the host)
    const float up = a.upstream;
    const float inx0 = a.inv_nx[0] * 0.25f * a.w_grad * up, iny0 = a.inv_ny[0] * 0.25f * a.w_grad * up;
    const float eps_g = a.eps_grad, eps_r = a.eps_rp;
    constexpr float kExpScale = -1.4426950408889634f / 3.0f;    // exp(-mean_c|dI|) = 2^(kExpScale * sum_c|dI|)

    pdl_trigger();
    pdl_wait();        // C1 of pyr_coef_kernel, and through it the statistics of phase A
    // scalars derived from the phase-A statistics, weights and upstream folded in
    float c1 = 0.f, c2 = 0.f, rpn = 0.f;
    {
        const double n = a.stats[ST_SI_N], S = a.stats[ST_SI_S], nr = a.stats[ST_RP_N];
        if (SI && n > 0.0) {
            c1 = (float)(2.0 / n) * a.w_si * up;
            c2 = (float)(-2.0 * (double)a.lambda * S / (n * n)) * a.w_si * up;
        }
        if (RP && nr > 0.0) rpn = (float)(1.0 / nr) * a.w_rp * up;
    }

    // L2 prefetch kStreamPrefetchRows rows ahead of the register loads: lanes 0..19 each own one 128-byte line of a
    // 128-column row segment (pred, gt, 3 x rgb  x 4 lines); the pointer is set up per segment and advanced by a row
    const int pf_t = lane >> 2, pf_seg = lane & 3;
    const bool pf_lane = lane < (SMOOTH ? 20 : 8);
    const size_t rowbytes = (size_t)W * sizeof(float);
    // Work items: every image's strip-rows (strip-major) cut into cpi equal shares, one per warp of the single
    // resident wave (more than one per warp only for huge batches).  cadl_debug_set_trace shows warps finishing
    // within +-12 % of each other; handing the last quarter of each image out dynamically in 4-row chunks was
    // measured SLOWER (114 vs 104 us: every chunk restarts the two-row prologue with its loads exposed).
    const int nwarps = gridDim.x * (kStreamThreads / 32);
    const int gwarp = blockIdx.x * (kStreamThreads / 32) + (tid >> 5);
    const int nitems = a.B * sa.cpi;
    const int SR = sa.nstrip * H;                      // strip-rows per image  (< 2^31: nstrip * H <= H * W / 4)
    unsigned long long t_start = 0;
    int items_done = 0;
    if (sa.trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    for (int item = gwarp; item < nitems; item += nwarps) {
        ++items_done;
        const int b = item / sa.cpi, wi = item - b * sa.cpi;
        int cur = (int)((long long)SR * wi / sa.cpi);
        const int end = (int)((long long)SR * (wi + 1) / sa.cpi);
        const int row = item;
        {
            float acc[BF_COUNT];
#pragma unroll
            for (int q = 0; q < BF_COUNT; ++q) acc[q] = 0.f;

            // pred / gt / grad / C1 are addressed as (kernel-parameter base) + 32-bit element offset, so no per-image
            // 64-bit pointers stay live across the row loop (B*H*W < 2^31, checked on the host); rgb needs 64 bits
            const int img = b * plane;
            const int c1img = b * (H >> 1) * W1;
            const float* __restrict__ rgbb = SMOOTH ? a.rgb + (size_t)b * 3 * plane : nullptr;
            float abw = 0.f;
            if (SMOOTH) abw = (1.0f / ((float)(a.img_psum[b] / ((double)H * W)) + a.eps_smooth)) * a.w_smooth * up;   // a_b (:192-193)
            const float snx = a.sm_nx * abw, sny = a.sm_ny * abw;
            float fxe = 1.f, fye = 1.f, rfx = 1.f, rfy = 1.f, cxv = 0.f, cyv = 0.f;
            bool mk_ok = true;
            if constexpr (RP) {
                float fx, fy;
                load_K(a, b, fx, fy, cxv, cyv);
                fxe = fx + a.eps_rp;
                fye = fy + a.eps_rp;
                rfx = __frcp_rn(fxe);
                rfy = __frcp_rn(fye);
                mk_ok = markstein_safe(fxe) && markstein_safe(fye);
            }

            while (cur < end) {
            const int strip = cur / H, ys = cur - strip * H;
            const int ye = (end - cur < H - ys) ? ys + (end - cur) : H;     // rows [ys, ye) of this strip
            cur += ye - ys;

            // Lanes right of the image (partial last strip) load the last in-image float4 instead: they neither
            // store nor count, and the last in-image lane takes its right neighbour from its halo pixel.
            const int gx0 = strip * 128 + 4 * lane;
            const bool lane_in = gx0 < W;                       // W % 4 == 0: a lane is fully inside or outside
            const int gxc = lane_in ? gx0 : W - 4;
            const bool lastlane = (lane == 31) || (gx0 + 4 >= W);
            // halo pixel: right neighbour for the last lane (clamped to the image: the pixel itself at the right
            // border, so that edge vanishes), left neighbour for lane 0 (itself at the left border)
            const int hx = lastlane ? (gx0 + 4 < W ? gx0 + 4 : W - 1) : (gx0 >= 1 ? gx0 - 1 : 0);
            const bool left_edge = (lane == 0) && (gx0 >= 1);    // lane 0 evaluates the edge to its left neighbour strip
            const bool pf_on = pf_lane && (strip * 128 + pf_seg * 32 < W);
            const int pf_end = ye + 1 < H ? ye + 1 : H;              // rows [.., ye] are read by this segment
            const char* pf_ptr = reinterpret_cast<const char*>(
                (pf_t == 0 ? a.pred + img : pf_t == 1 ? a.gt + img : a.rgb + ((size_t)b * 3 + (pf_t >= 2 ? pf_t - 2 : 0)) * plane) +
                (size_t)(ys + kStreamPrefetchRows) * W + strip * 128 + pf_seg * 32);
            const float ufx0 = (float)gx0;                       // u of the lane's first pixel; u + k is exact

            // issue the global loads of one image row; rows outside the image are the border row again (every
            // vertical edge across the border then has residual exactly 0).  No use of the values here.
            auto fetch = [&](int off, int offh, StreamRow& R) {
                const float4 p4 = __ldg(reinterpret_cast<const float4*>(a.pred + (img + off)));
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.gt + (img + off)));
                R.p[0] = p4.x; R.p[1] = p4.y; R.p[2] = p4.z; R.p[3] = p4.w;
                R.g[0] = g4.x; R.g[1] = g4.y; R.g[2] = g4.z; R.g[3] = g4.w;
                R.hp = __ldg(a.pred + (img + offh));
                R.hg = __ldg(a.gt + (img + offh));
                if constexpr (SMOOTH) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float4 v = ldg_stream(reinterpret_cast<const float4*>(rgbb + c * plane + off));
                        R.I[c][0] = v.x; R.I[c][1] = v.y; R.I[c][2] = v.z; R.I[c][3] = v.w;
                        R.hI[c] = __ldg(rgbb + c * plane + offh);
                    }
                }
            };
            // logs of a row and its right neighbours across lanes -- called when the row's loads have landed
            auto finish_row = [&](StreamRow& R) {
                float2 v[5];
                v[0] = make_float2(clamp_nan(R.p[0], eps_g, 1000.0f), clamp_nan(R.p[1], eps_g, 1000.0f));
                v[1] = make_float2(clamp_nan(R.p[2], eps_g, 1000.0f), clamp_nan(R.p[3], eps_g, 1000.0f));
                v[2] = make_float2(clamp_nan(R.g[0], eps_g, 1000.0f), clamp_nan(R.g[1], eps_g, 1000.0f));
                v[3] = make_float2(clamp_nan(R.g[2], eps_g, 1000.0f), clamp_nan(R.g[3], eps_g, 1000.0f));
                v[4] = make_float2(clamp_nan(R.hp, eps_g, 1000.0f), clamp_nan(R.hg, eps_g, 1000.0f));
                log_exact2_n<5>(v);
                R.lp[0] = v[0].x; R.lp[1] = v[0].y; R.lp[2] = v[1].x; R.lp[3] = v[1].y;
                R.lg[0] = v[2].x; R.lg[1] = v[2].y; R.lg[2] = v[3].x; R.lg[3] = v[3].y;
                R.hlp = v[4].x; R.hlg = v[4].y;
                if constexpr (SMOOTH) {
                    const float pr = __shfl_down_sync(0xffffffffu, R.p[0], 1);
                    R.p[4] = lastlane ? R.hp : pr;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float ir = __shfl_down_sync(0xffffffffu, R.I[c][0], 1);
                        R.I[c][4] = lastlane ? R.hI[c] : ir;
                    }
                }
            };
            // terms of the vertical edges (row C -> row N)
            auto yterms = [&](bool count, const StreamRow& C, const StreamRow& N, float (&sy)[4], float (&ty)[4]) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float e = (N.lp[k] - C.lp[k]) - (N.lg[k] - C.lg[k]);      // depth_loss.h:151-163
                    sy[k] = sgn3(e);
                    if (count) acc[BF_GY0] += fabsf(e);
                }
                if constexpr (SMOOTH) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float s = fabsf(N.I[0][k] - C.I[0][k]) + fabsf(N.I[1][k] - C.I[1][k]) + fabsf(N.I[2][k] - C.I[2][k]);
                        const float wy = ex2_approx(s * kExpScale);                 // depth_loss.h:218-227
                        const float d = N.p[k] - C.p[k];
                        ty[k] = wy * sgn3(d);
                        if (count) acc[BF_SMY] = fmaf(wy, fabsf(d), acc[BF_SMY]);
                    }
                }
            };

            // One row: C is the current row (complete), N receives the next one.  up: signed terms of the edges
            // to the row above (from the previous step); dn: those to the row below (for the next step).
            auto step = [&](int gy, StreamRow& C, StreamRow& N, const float (&sy_up)[4], const float (&ty_up)[4],
                            float (&sy_dn)[4], float (&ty_dn)[4]) {
                // 1. issue next row's loads; they are consumed at step 4, after ~2/3 of this row's arithmetic
                const int rn = (gy + 1 < H ? gy + 1 : gy) * W;
                fetch(rn + gxc, rn + hx, N);
                if (pf_on && gy + kStreamPrefetchRows < pf_end) prefetch_l2(pf_ptr);
                pf_ptr += rowbytes;
                uchar4 mk4 = make_uchar4(0, 0, 0, 0);
                if constexpr (HAS_MASK) mk4 = __ldg(reinterpret_cast<const uchar4*>(a.mask + img + gy * W + gxc));
                const float2 ccv = __ldg(reinterpret_cast<const float2*>(sa.c1 + (c1img + (gy >> 1) * W1 + (gxc >> 1))));

                // 2. horizontal edges of the current row: each lane evaluates the four edges to the right of its
                //    pixels; the sign of the edge to its left comes from the left lane
                float gm[4], smg[4] = {0.f, 0.f, 0.f, 0.f};
                {
                    const float lr_p = __shfl_down_sync(0xffffffffu, C.lp[0], 1), lr_g = __shfl_down_sync(0xffffffffu, C.lg[0], 1);
                    const float lpx[5] = {C.lp[0], C.lp[1], C.lp[2], C.lp[3], lastlane ? C.hlp : lr_p};
                    const float lgx[5] = {C.lg[0], C.lg[1], C.lg[2], C.lg[3], lastlane ? C.hlg : lr_g};
                    float sx[5];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float e = (lpx[j + 1] - lpx[j]) - (lgx[j + 1] - lgx[j]);   // depth_loss.h:140-148,162
                        sx[j + 1] = sgn3(e);
                        if (lane_in) acc[BF_GX0] += fabsf(e);
                    }
                    float sl = __shfl_up_sync(0xffffffffu, sx[4], 1);
                    if (lane == 0) sl = left_edge ? sgn3((C.lp[0] - C.hlp) - (C.lg[0] - C.hlg)) : 0.f;
                    sx[0] = sl;
#pragma unroll
                    for (int k = 0; k < 4; ++k) gm[k] = (sx[k] - sx[k + 1]) * inx0;
                }
                if constexpr (SMOOTH) {
                    float tx[5];                                  // tx[j]: edge (x_{j-1} -> x_j); j = 0 belongs to the left lane
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float s = fabsf(C.I[0][k + 1] - C.I[0][k]) + fabsf(C.I[1][k + 1] - C.I[1][k]) + fabsf(C.I[2][k + 1] - C.I[2][k]);
                        const float wx = ex2_approx(s * kExpScale);                 // depth_loss.h:211-226
                        const float d = C.p[k + 1] - C.p[k];
                        tx[k + 1] = wx * sgn3(d);
                        if (lane_in) acc[BF_SMX] = fmaf(wx, fabsf(d), acc[BF_SMX]);
                    }
                    float tl = __shfl_up_sync(0xffffffffu, tx[4], 1);
                    if (lane == 0) {
                        tl = 0.f;
                        if (left_edge) {
                            const float s = fabsf(C.I[0][0] - C.hI[0]) + fabsf(C.I[1][0] - C.hI[1]) + fabsf(C.I[2][0] - C.hI[2]);
                            tl = ex2_approx(s * kExpScale) * sgn3(C.p[0] - C.hp);
                        }
                    }
                    tx[0] = tl;
#pragma unroll
                    for (int k = 0; k < 4; ++k) smg[k] = (tx[k] - tx[k + 1]) * snx;
                }

                // 3. pointwise terms
                const bool um[4] = {mk4.x != 0, mk4.y != 0, mk4.z != 0, mk4.w != 0};
                float ayv = 0.f, yh = 0.f;
                if constexpr (RP) {
                    ayv = (float)gy - cyv;
                    yh = ayv * rfy;                               // d pY / d p (tolerance path)
                }
                float rpk[4], pw[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float p = C.p[k];
                    rpk[k] = rcp_approx(p);
                    float gsum = (k < 2 ? ccv.x : ccv.y);
                    if constexpr (SI) {
                        const float g = C.g[k];
                        const bool m = HAS_MASK ? um[k] : (g > eps_g);        // eps_si == eps_grad on this path
                        const float d = C.lp[k] - C.lg[k];
                        if (m && in_range_pos(p, eps_g, 1000.0f)) gsum = fmaf(fmaf(c1, d, c2), rpk[k], gsum);
                    }
                    if constexpr (RP) {
                        const float g = C.g[k];
                        const bool m = HAS_MASK ? um[k] : (g > eps_r);
                        if (m && lane_in) {
                            // same operations, same order as depth_loss.h:299-315 (see cadl_phase_b.cuh)
                            const float axk = __fadd_rn(__fadd_rn(ufx0, (float)k), -cxv);      // (u - cx), u = float(column)
                            const float xhk = axk * rfx;                                       // d pX / d p: tolerance path
                            float pX, gX, pY, gY;
                            if (mk_ok) {
                                pX = div_by_const(__fmul_rn(axk, p), fxe, rfx);
                                gX = div_by_const(__fmul_rn(axk, g), fxe, rfx);
                                pY = div_by_const(__fmul_rn(ayv, p), fye, rfy);
                                gY = div_by_const(__fmul_rn(ayv, g), fye, rfy);
                            } else {
                                pX = __fdiv_rn(__fmul_rn(axk, p), fxe);
                                gX = __fdiv_rn(__fmul_rn(axk, g), fxe);
                                pY = __fdiv_rn(__fmul_rn(ayv, p), fye);
                                gY = __fdiv_rn(__fmul_rn(ayv, g), fye);
                            }
                            const float dX = pX - gX, dY = pY - gY, dZ = p - g;
                            const float ss = fmaf(dZ, dZ, fmaf(dY, dY, dX * dX)) + eps_r;
                            const float re = rsqrt_approx(ss);
                            acc[BF_RP_E] = fmaf(ss, re, acc[BF_RP_E]);              // e = sqrt(ss)
                            gsum = fmaf(fmaf(dX, xhk, fmaf(dY, yh, dZ)) * re, rpn, gsum);
                        }
                    }
                    pw[k] = gsum;
                }

                // 4. the next row has landed: its logs, the vertical edges, assembly and the 128-bit store
                finish_row(N);
#pragma unroll
                for (int k = 0; k < 4; ++k) ty_dn[k] = 0.f;
                yterms(lane_in, C, N, sy_dn, ty_dn);
                float out[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float gsum = pw[k];
                    if constexpr (SMOOTH) gsum += fmaf(ty_up[k] - ty_dn[k], sny, smg[k]);
                    const float gmk = fmaf(sy_up[k] - sy_dn[k], iny0, gm[k]);
                    gsum = in_range_pos(C.p[k], eps_g, 1000.0f) ? fmaf(gmk, rpk[k], gsum) : gsum;   // clamp backward
                    out[k] = gsum;
                }
                if (a.grad && lane_in)
                    *reinterpret_cast<float4*>(a.grad + (img + gy * W + gx0)) = make_float4(out[0], out[1], out[2], out[3]);
            };

            // prologue: the row above this segment only contributes its lower edges
            StreamRow RA, RB;
            float u0s[4], u0t[4] = {0.f, 0.f, 0.f, 0.f}, u1s[4], u1t[4];
            {
                const int r0 = (ys > 0 ? ys - 1 : 0) * W, r1 = ys * W;
                fetch(r0 + gxc, r0 + hx, RB);
                fetch(r1 + gxc, r1 + hx, RA);
                finish_row(RB);
                finish_row(RA);
                yterms(false, RB, RA, u0s, u0t);
            }
            // two rows per trip, the two row buffers and the two edge buffers trading places: no register copies
            for (int gy = ys; gy < ye; gy += 2) {
                step(gy, RA, RB, u0s, u0t, u1s, u1t);
                if (gy + 1 >= ye) break;
                step(gy + 1, RB, RA, u1s, u1t, u0s, u0t);
            }
            }   // segments

            // this share's partial row (fixed content whichever warp ran it: deterministic)
            float v[BF_COUNT];
#pragma unroll
            for (int q = 0; q < BF_COUNT; ++q) v[q] = 0.f;
            v[BF_GX0] = warp_sum(acc[BF_GX0]);
            v[BF_GY0] = warp_sum(acc[BF_GY0]);
            if constexpr (SMOOTH) { v[BF_SMX] = warp_sum(acc[BF_SMX]); v[BF_SMY] = warp_sum(acc[BF_SMY]); }
            if constexpr (RP) v[BF_RP_E] = warp_sum(acc[BF_RP_E]);
            float mine = 0.f;
#pragma unroll
            for (int q = 0; q < BF_COUNT; ++q) mine = (lane == q) ? v[q] : mine;
            if (lane < BF_COUNT) sa.chunk_part[(size_t)row * BF_COUNT + lane] = (double)mine;
            // The warp that completes an image folds that image's rows into ONE row, in a fixed order, while the
            // other warps keep streaming: the kernel-final reduction then reads B rows instead of thousands.
            __threadfence();
            int last = 0;
            if (lane == 0) last = atomicAdd(&sa.img_cnt[b], 1u) == (unsigned)sa.cpi - 1u;
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) fold_image_rows(sa.chunk_part + (size_t)b * sa.cpi * BF_COUNT, sa.cpi, a.b_part + (size_t)b * BF_COUNT,
                                      sa.img_cnt + b, lane);
        }
    }
    if (sa.trace && lane == 0 && gwarp < sa.trace_cap) {
        unsigned long long t_end;
        unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        sa.trace[4 * gwarp + 0] = smid;
        sa.trace[4 * gwarp + 1] = t_start;
        sa.trace[4 * gwarp + 2] = t_end;
        sa.trace[4 * gwarp + 3] = (unsigned long long)items_done;
    }
    if (!sa.finalize_inline) return;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned t = atomicAdd(&a.hdr->ticket_b, 1u);
        s_last = (t == gridDim.x - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        finalize_results(a, s_d);
        if (a.metrics) write_metric_results(a.stats, a.metrics, *a.results, tid);
    }
}

// The pass after the streaming kernel when the smoothness term is on: grad[b, :] -= off[b] (every CTA of image b
// derives off[b] itself from the image's folded partial row -- the same arithmetic as finalize_results), while ONE
// extra CTA does the kernel-final reduction and writes cadl_results.  That reduction is a chain of dependent L2
// round trips (~13 us as the streaming kernel's last CTA, with 147 SMs idle); here it runs beside the offset pass.
__global__ void __launch_bounds__(256) stream_finish_kernel(const PhaseBArgs a, int vec_ok) {
    __shared__ double s_d[8];
    __shared__ float s_off;
    // grid (bx, B + 1): row 0 is dispatched first and holds the reduction CTA, rows 1..B are the images
    const int tid = threadIdx.x, b = (int)blockIdx.y - 1;
    pdl_wait();        // the gradient and the partial rows of phase_b_stream_kernel
    if (b < 0) {
        if (blockIdx.x == 0) {
            finalize_results(a, s_d);
            if (a.metrics) write_metric_results(a.stats, a.metrics, *a.results, tid);
        }
        return;
    }
    if (tid == 0) {
        double Lb;
        float off;
        smooth_image_share(a, b, __ldcg(a.b_part + (size_t)b * BF_COUNT + BF_SMX), __ldcg(a.b_part + (size_t)b * BF_COUNT + BF_SMY), Lb, off);
        s_off = off;
    }
    __syncthreads();
    const float o = s_off;
    const int HW = a.H * a.W;
    float* g = a.grad + (size_t)b * HW;
    if (vec_ok) {
        float4* g4 = reinterpret_cast<float4*>(g);
        const int n4 = HW >> 2;
        for (int i = blockIdx.x * blockDim.x + tid; i < n4; i += gridDim.x * blockDim.x) {
            float4 v = g4[i];
            v.x -= o; v.y -= o; v.z -= o; v.w -= o;
            g4[i] = v;
        }
    } else {
        for (int i = blockIdx.x * blockDim.x + tid; i < HW; i += gridDim.x * blockDim.x) g[i] -= o;
    }
}

}  // namespace cadl
