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
//                     no halo except one pixel per warp end.  The image's strip-rows are divided evenly over the
//                     warps of ONE wave (148 SMs x 16 warps), so there is no tail.
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
                                                       int B, int H, int W, float eps, PyrArrays py) {
    const int W8 = W >> 3, H8 = H >> 3;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H8 * W8) return;
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
    const int W8 = a.W >> 3, H8 = a.H >> 3;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // GX1, GY1, GX2, GY2, GX3, GY3
    if (idx < a.B * H8 * W8) {
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
    int wpi;             // warps per image
    int nstrip;          // 128-column strips per image row
    int rows_total;      // partial rows finalize_results sums (B * wpi + rows of pyr_coef_kernel)
};

template <int F, bool HAS_MASK>
__global__ void __launch_bounds__(kThreadsB, 2) phase_b_stream_kernel(const PhaseBArgs a, const StreamArgs sa) {
    __shared__ double s_d[8];
    __shared__ int s_last;
    constexpr bool SMOOTH = (F & FB_SMOOTH) != 0;
    constexpr bool SI = (F & FB_SI) != 0;
    constexpr bool RP = (F & FB_RP) != 0;
    static_assert((F & FB_GRAD) != 0, "the streaming kernel is the gradient-matching path");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = a.H, W = a.W;
    const int gw = blockIdx.x * (kThreadsB / 32) + warp;
    const int b = gw / sa.wpi, wi = gw - b * sa.wpi;

    if (b < a.B) {
        float acc[BF_COUNT];
#pragma unroll
        for (int q = 0; q < BF_COUNT; ++q) acc[q] = 0.f;

        const int img = b * H * W;                               // B*H*W < 2^31 (checked on the host)
        const float* __restrict__ predb = a.pred + img;
        const float* __restrict__ gtb = a.gt + img;
        const float* __restrict__ rgbb = SMOOTH ? a.rgb + (size_t)b * 3 * H * W : nullptr;
        const float* __restrict__ c1b = sa.c1 + (size_t)b * (H >> 1) * (W >> 1);
        const int plane = H * W, W1 = W >> 1;
        const float up = a.upstream;

        // scalars derived from the phase-A statistics, weights and upstream folded in
        float c1 = 0.f, c2 = 0.f, rpn = 0.f, abw = 0.f;
        {
            const double n = a.stats[ST_SI_N], S = a.stats[ST_SI_S], nr = a.stats[ST_RP_N];
            if (SI && n > 0.0) {
                c1 = (float)(2.0 / n) * a.w_si * up;
                c2 = (float)(-2.0 * (double)a.lambda * S / (n * n)) * a.w_si * up;
            }
            if (RP && nr > 0.0) rpn = (float)(1.0 / nr) * a.w_rp * up;
            if (SMOOTH) abw = (1.0f / ((float)(a.img_psum[b] / ((double)H * W)) + a.eps_smooth)) * a.w_smooth * up;   // a_b (:192-193)
        }
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
        const float inx0 = a.inv_nx[0] * 0.25f * a.w_grad * up, iny0 = a.inv_ny[0] * 0.25f * a.w_grad * up;
        const float snx = a.sm_nx * abw, sny = a.sm_ny * abw;
        const float eps_g = a.eps_grad, eps_r = a.eps_rp;
        constexpr float kExpScale = -1.4426950408889634f / 3.0f;    // exp(-mean_c|dI|) = 2^(kExpScale * sum_c|dI|)

        // this warp's share of the image's strip-rows (strip-major): [cur, end)
        const long long SR = (long long)sa.nstrip * H;
        int cur = (int)(SR * wi / sa.wpi);
        const int end = (int)(SR * (wi + 1) / sa.wpi);

        while (cur < end) {
            const int strip = cur / H, ys = cur - strip * H;
            const int ye = (end - cur < H - ys) ? ys + (end - cur) : H;     // rows [ys, ye) of this strip
            cur += ye - ys;

            const int gx0 = strip * 128 + 4 * lane;
            const bool lane_in = gx0 < W;                       // W % 4 == 0: a lane is fully inside or outside
            const int gxr = clampi(gx0 + 4, 0, W - 1);                  // right neighbour column of the last lane
            const bool right_in = gx0 + 4 < W;
            const bool endlane = (lane == 31) || (lane == 0 && gx0 >= 1);
            const int hx = (lane == 31) ? gxr : (gx0 >= 1 ? gx0 - 1 : 0);     // column of the end lanes' halo pixel
            const bool h_rgb_ok = (lane == 31) ? right_in : true;
            float axk[4] = {0.f, 0.f, 0.f, 0.f}, xhk[4] = {0.f, 0.f, 0.f, 0.f};
            if constexpr (RP) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    axk[k] = (float)(gx0 + k) - cxv;
                    xhk[k] = axk[k] * rfx;                       // d pX / d p: tolerance path
                }
            }

            // ---- row state ----
            float pc[5], gc[4], Ic[3][5];       // current row: own 4 (+ right neighbour)
            float pn[5], gn[4], In[3][5];       // next row
            float hn_p = 0.f, hn_g = 0.f, hn_I[3] = {0.f, 0.f, 0.f};   // end lanes' halo pixel of the next row
            float hc_p = 0.f, hc_I[3] = {0.f, 0.f, 0.f};               // ... of the current row
            float hl_p = 0.f, hl_g = 0.f;                              // logs of the current row's halo pixel
            float lpc[4], lgc[4];                                      // logs of the current row
            float sy_up[4] = {0.f, 0.f, 0.f, 0.f}, ty_up[4] = {0.f, 0.f, 0.f, 0.f};

            // issue the global loads of one image row (clamped at the borders); no use of the values here
            auto fetch = [&](int gy_raw, float (&p)[5], float (&g)[4], float (&I)[3][5], float& hp, float& hg, float (&hI)[3]) {
                const bool in_img = (gy_raw >= 0) && (gy_raw < H);
                const int ro = clampi(gy_raw, 0, H - 1) * W;
                if (lane_in) {
                    const float4 p4 = __ldg(reinterpret_cast<const float4*>(predb + ro + gx0));
                    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gtb + ro + gx0));
                    p[0] = p4.x; p[1] = p4.y; p[2] = p4.z; p[3] = p4.w;
                    g[0] = g4.x; g[1] = g4.y; g[2] = g4.z; g[3] = g4.w;
                } else {   // lanes right of the image hold the replicated border pixel (their edges vanish)
                    const float ps = __ldg(predb + ro + W - 1), gs = __ldg(gtb + ro + W - 1);
                    p[0] = p[1] = p[2] = p[3] = ps;
                    g[0] = g[1] = g[2] = g[3] = gs;
                }
                if constexpr (SMOOTH) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (in_img && lane_in) v = ldg_stream(reinterpret_cast<const float4*>(rgbb + c * plane + ro + gx0));
                        I[c][0] = v.x; I[c][1] = v.y; I[c][2] = v.z; I[c][3] = v.w;
                    }
                }
                // the two lanes at the warp's ends also fetch the pixel beyond their end: lane 31 its right
                // neighbour, lane 0 its left neighbour (same registers, same instructions, different lanes)
                if (endlane) {
                    hp = __ldg(predb + ro + hx);
                    hg = __ldg(gtb + ro + hx);
                    if constexpr (SMOOTH) {
                        const bool ok = in_img && h_rgb_ok;
#pragma unroll
                        for (int c = 0; c < 3; ++c) hI[c] = ok ? __ldg(rgbb + c * plane + ro + hx) : 0.f;
                    }
                }
            };
            // logs of a row (depth_loss.h:115-116) and its right neighbours across lanes
            auto finish_row = [&](float (&p)[5], const float (&g)[4], float (&I)[3][5], float hp, float hg, const float (&hI)[3],
                                  float (&lp)[4], float (&lg)[4], float& hlp, float& hlg) {
                const float2 a0 = log_exact2(make_float2(clamp_nan(p[0], eps_g, 1000.0f), clamp_nan(p[1], eps_g, 1000.0f)));
                const float2 a1 = log_exact2(make_float2(clamp_nan(p[2], eps_g, 1000.0f), clamp_nan(p[3], eps_g, 1000.0f)));
                const float2 b0 = log_exact2(make_float2(clamp_nan(g[0], eps_g, 1000.0f), clamp_nan(g[1], eps_g, 1000.0f)));
                const float2 b1 = log_exact2(make_float2(clamp_nan(g[2], eps_g, 1000.0f), clamp_nan(g[3], eps_g, 1000.0f)));
                const float2 hh = log_exact2(make_float2(clamp_nan(hp, eps_g, 1000.0f), clamp_nan(hg, eps_g, 1000.0f)));
                lp[0] = a0.x; lp[1] = a0.y; lp[2] = a1.x; lp[3] = a1.y;
                lg[0] = b0.x; lg[1] = b0.y; lg[2] = b1.x; lg[3] = b1.y;
                hlp = hh.x; hlg = hh.y;
                if constexpr (SMOOTH) {
                    const float pr = __shfl_down_sync(0xffffffffu, p[0], 1);
                    p[4] = (lane == 31) ? hp : pr;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float ir = __shfl_down_sync(0xffffffffu, I[c][0], 1);
                        I[c][4] = (lane == 31) ? hI[c] : ir;
                    }
                }
            };
            // terms of the vertical edges (current row -> next row)
            auto yterms = [&](bool count, float (&sy)[4], float (&ty)[4], const float (&lpn)[4], const float (&lgn)[4]) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float e = (lpn[k] - lpc[k]) - (lgn[k] - lgc[k]);      // depth_loss.h:151-163
                    sy[k] = sgn3(e);
                    if (count) acc[BF_GY0] += fabsf(e);
                }
                if constexpr (SMOOTH) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float s = fabsf(In[0][k] - Ic[0][k]) + fabsf(In[1][k] - Ic[1][k]) + fabsf(In[2][k] - Ic[2][k]);
                        const float wy = ex2_approx(s * kExpScale);                 // depth_loss.h:218-227
                        const float d = pn[k] - pc[k];
                        ty[k] = wy * sgn3(d);
                        if (count) acc[BF_SMY] = fmaf(wy, fabsf(d), acc[BF_SMY]);
                    }
                }
            };

            // prologue: the row above this segment only contributes its lower edges
            float lpn[4], lgn[4], hln_p, hln_g;
            {
                float hp0 = 0.f, hg0 = 0.f, hI0[3] = {0.f, 0.f, 0.f}, t0, t1;
                fetch(ys - 1, pc, gc, Ic, hp0, hg0, hI0);
                fetch(ys, pn, gn, In, hn_p, hn_g, hn_I);
                finish_row(pc, gc, Ic, hp0, hg0, hI0, lpc, lgc, t0, t1);
                finish_row(pn, gn, In, hn_p, hn_g, hn_I, lpn, lgn, hln_p, hln_g);
                yterms(false, sy_up, ty_up, lpn, lgn);
            }
            auto roll = [&]() {
                hc_p = hn_p; hl_p = hln_p; hl_g = hln_g;
#pragma unroll
                for (int c = 0; c < 3; ++c) hc_I[c] = hn_I[c];
#pragma unroll
                for (int k = 0; k < 4; ++k) { lpc[k] = lpn[k]; lgc[k] = lgn[k]; gc[k] = gn[k]; }
#pragma unroll
                for (int k = 0; k < 5; ++k) pc[k] = pn[k];
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int k = 0; k < 5; ++k) Ic[c][k] = In[c][k];
            };
            roll();

            for (int gy = ys; gy < ye; ++gy) {
                // 1. issue next row's loads; they are consumed at step 4, after ~2/3 of this row's arithmetic
                fetch(gy + 1, pn, gn, In, hn_p, hn_g, hn_I);
                uchar4 mk4 = make_uchar4(0, 0, 0, 0);
                if constexpr (HAS_MASK) {
                    if (lane_in) mk4 = __ldg(reinterpret_cast<const uchar4*>(a.mask + img + gy * W + gx0));
                }
                float2 ccv = make_float2(0.f, 0.f);
                if (lane_in) ccv = __ldg(reinterpret_cast<const float2*>(c1b + (gy >> 1) * W1 + (gx0 >> 1)));

                // 2. horizontal edges of the current row: each lane evaluates the four edges to the right of its
                //    pixels; the sign of the edge to its left comes from the left lane
                float gm[4], smg[4] = {0.f, 0.f, 0.f, 0.f};
                {
                    const float lr_p = __shfl_down_sync(0xffffffffu, lpc[0], 1), lr_g = __shfl_down_sync(0xffffffffu, lgc[0], 1);
                    const float lpx[5] = {lpc[0], lpc[1], lpc[2], lpc[3], (lane == 31) ? hl_p : lr_p};
                    const float lgx[5] = {lgc[0], lgc[1], lgc[2], lgc[3], (lane == 31) ? hl_g : lr_g};
                    float sx[5];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float e = (lpx[j + 1] - lpx[j]) - (lgx[j + 1] - lgx[j]);   // depth_loss.h:140-148,162
                        sx[j + 1] = sgn3(e);
                        if (lane_in) acc[BF_GX0] += fabsf(e);
                    }
                    float sl = __shfl_up_sync(0xffffffffu, sx[4], 1);
                    if (lane == 0) {
                        sl = 0.f;                                 // left neighbour lives in another strip: evaluate that edge here
                        if (gx0 >= 1) sl = sgn3((lpc[0] - hl_p) - (lgc[0] - hl_g));
                    }
                    sx[0] = sl;
#pragma unroll
                    for (int k = 0; k < 4; ++k) gm[k] = (sx[k] - sx[k + 1]) * inx0;
                }
                if constexpr (SMOOTH) {
                    float tx[5];                                  // tx[j]: edge (x_{j-1} -> x_j); j = 0 belongs to the left lane
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float s = fabsf(Ic[0][k + 1] - Ic[0][k]) + fabsf(Ic[1][k + 1] - Ic[1][k]) + fabsf(Ic[2][k + 1] - Ic[2][k]);
                        const float wx = ex2_approx(s * kExpScale);                 // depth_loss.h:211-226
                        const float d = pc[k + 1] - pc[k];
                        tx[k + 1] = wx * sgn3(d);
                        if (lane_in) acc[BF_SMX] = fmaf(wx, fabsf(d), acc[BF_SMX]);
                    }
                    float tl = __shfl_up_sync(0xffffffffu, tx[4], 1);
                    if (lane == 0) {
                        tl = 0.f;
                        if (gx0 >= 1) {
                            const float s = fabsf(Ic[0][0] - hc_I[0]) + fabsf(Ic[1][0] - hc_I[1]) + fabsf(Ic[2][0] - hc_I[2]);
                            tl = ex2_approx(s * kExpScale) * sgn3(pc[0] - hc_p);
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
                    const float p = pc[k];
                    rpk[k] = rcp_approx(p);
                    float gsum = (k < 2 ? ccv.x : ccv.y);
                    if constexpr (SI) {
                        const float g = gc[k];
                        const bool m = HAS_MASK ? um[k] : (g > eps_g);        // eps_si == eps_grad on this path
                        const float d = lpc[k] - lgc[k];
                        if (m && in_range_pos(p, eps_g, 1000.0f)) gsum = fmaf(fmaf(c1, d, c2), rpk[k], gsum);
                    }
                    if constexpr (RP) {
                        const float g = gc[k];
                        const bool m = HAS_MASK ? um[k] : (g > eps_r);
                        if (m && lane_in) {
                            // same operations, same order as depth_loss.h:299-315 (see cadl_phase_b.cuh)
                            float pX, gX, pY, gY;
                            if (mk_ok) {
                                pX = div_by_const(__fmul_rn(axk[k], p), fxe, rfx);
                                gX = div_by_const(__fmul_rn(axk[k], g), fxe, rfx);
                                pY = div_by_const(__fmul_rn(ayv, p), fye, rfy);
                                gY = div_by_const(__fmul_rn(ayv, g), fye, rfy);
                            } else {
                                pX = __fdiv_rn(__fmul_rn(axk[k], p), fxe);
                                gX = __fdiv_rn(__fmul_rn(axk[k], g), fxe);
                                pY = __fdiv_rn(__fmul_rn(ayv, p), fye);
                                gY = __fdiv_rn(__fmul_rn(ayv, g), fye);
                            }
                            const float dX = pX - gX, dY = pY - gY, dZ = p - g;
                            const float ss = fmaf(dZ, dZ, fmaf(dY, dY, dX * dX)) + eps_r;
                            const float re = rsqrt_approx(ss);
                            acc[BF_RP_E] = fmaf(ss, re, acc[BF_RP_E]);              // e = sqrt(ss)
                            gsum = fmaf(fmaf(dX, xhk[k], fmaf(dY, yh, dZ)) * re, rpn, gsum);
                        }
                    }
                    pw[k] = gsum;
                }

                // 4. the next row has landed: its logs, the vertical edges, assembly and the 128-bit store
                finish_row(pn, gn, In, hn_p, hn_g, hn_I, lpn, lgn, hln_p, hln_g);
                float sy_dn[4], ty_dn[4] = {0.f, 0.f, 0.f, 0.f};
                yterms(lane_in, sy_dn, ty_dn, lpn, lgn);
                float out[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float gsum = pw[k];
                    if constexpr (SMOOTH) gsum += fmaf(ty_up[k] - ty_dn[k], sny, smg[k]);
                    const float gmk = fmaf(sy_up[k] - sy_dn[k], iny0, gm[k]);
                    gsum = in_range_pos(pc[k], eps_g, 1000.0f) ? fmaf(gmk, rpk[k], gsum) : gsum;   // clamp backward
                    out[k] = gsum;
                }
                if (a.grad && lane_in)
                    *reinterpret_cast<float4*>(a.grad + img + gy * W + gx0) = make_float4(out[0], out[1], out[2], out[3]);

                // roll the row state
#pragma unroll
                for (int k = 0; k < 4; ++k) { sy_up[k] = sy_dn[k]; ty_up[k] = ty_dn[k]; }
                roll();
            }
        }

        // this warp's partial row
#pragma unroll
        for (int q = 0; q < BF_COUNT; ++q) {
            const float v = warp_sum(acc[q]);
            if (lane == q) a.b_part[(size_t)gw * BF_COUNT + q] = (double)v;
        }
    }
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
        if (tid == 0 && a.metrics) write_metric_results(a.stats, a.metrics, *a.results);
    }
}

}  // namespace cadl
