// cadl phase B, fast path: the fused gradient pass for the aligned case
//   W % 4 == 0, 16-byte aligned tensors, H % 8 == 0, W % 8 == 0, num_scales == 4, eps_si == eps_grad
// (every BASELINE configuration: 240x320, 480x640, 960x1280).  Anything else takes the generic kernel in
// cadl_phase_b.cuh, which computes the same values.
//
// These kernels are ISSUE-bound, not bandwidth-bound (ncu: ~700 lane-instructions per pixel in the first
// version), so the design minimises instructions per pixel:
//   * every stencil edge (gradient matching at 4 scales, smoothness) is evaluated ONCE by the pixel that
//     owns it; the neighbour reuses the signed result through registers (vertical: the warp marches down
//     its rows) or a shuffle (horizontal);
//   * logs are evaluated once per pixel, in place over the staged raw tile, with the packed fp32x2 replica
//     of logf (cadl_math.cuh) -- bit-identical to ATen's CUDA log, which keeps sign(residual) identical;
//   * image borders need no per-pixel predicates: staging replicates the edge pixel into the halo, so a
//     non-existent edge has residual exactly 0, sign 0 and |.| 0;
//   * divisions by (fx+eps), (fy+eps) are Markstein-corrected reciprocal multiplies (correctly rounded);
//     1/p, 1/e, exp(-x) on tolerance-only paths use the SFU approximations.
//
// CTA = 256 threads, tile 48 x 128; shared memory ~101 KB so two CTAs share an SM and one CTA's staging
// overlaps the other's arithmetic.
#pragma once
#include "cadl_common.cuh"
#include "cadl_math.cuh"
#include "cadl_phase_a.cuh"
#include "cadl_phase_b.cuh"

namespace cadl {

constexpr int FTH = 48;                 // tile rows
constexpr int FTW = 128;                // tile cols
constexpr int FRH = FTH + 2 * HALO;     // 64 staged rows
constexpr int FRW = FTW + 2 * HALO;     // 144 staged cols
constexpr int FRPW = FTH / (kThreadsB / 32);   // 6 rows per warp

__host__ __device__ constexpr int fpool_h(int s) { return FTH / (1 << s) + 2; }
__host__ __device__ constexpr int fpool_w(int s) { return FTW / (1 << s) + 2; }
__host__ __device__ constexpr int fcc_h(int s) { return FTH / (1 << s); }
__host__ __device__ constexpr int fcc_w(int s) { return FTW / (1 << s); }
__host__ __device__ constexpr int fpool_off(int s) {
    return s == 1 ? 0 : (s == 2 ? fpool_h(1) * fpool_w(1) : fpool_h(1) * fpool_w(1) + fpool_h(2) * fpool_w(2));
}
constexpr int kFPoolCells = fpool_h(1) * fpool_w(1) + fpool_h(2) * fpool_w(2) + fpool_h(3) * fpool_w(3);
__host__ __device__ constexpr int fcc_off(int s) {
    return s == 1 ? 0 : (s == 2 ? fcc_h(1) * fcc_w(1) : fcc_h(1) * fcc_w(1) + fcc_h(2) * fcc_w(2));
}
constexpr int kFCCCells = fcc_h(1) * fcc_w(1) + fcc_h(2) * fcc_w(2) + fcc_h(3) * fcc_w(3);
constexpr size_t kFastSmemFloats = 2 * FRH * FRW + 2 * kFPoolCells + kFCCCells;
constexpr size_t kFastSmemBytes = kFastSmemFloats * sizeof(float);

struct FastSmem {
    float* sp;   // [FRH][FRW] raw pred, later log pred (in place) on the 1-pixel ring + interior
    float* sg;   // [FRH][FRW] raw gt,   later log gt
    float* pl;   // pooled log pred (scales 1..3, 1-cell halo)
    float* pg;   // pooled log gt
    float* cc;   // per-cell coefficients; scale 1 finally holds c1 + c2 + c3 (what a pixel adds)
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Sums of one 8x8 block for scale S (cells of f x f, f = 2^S), row-major sequential inside each window --
// the loop order of ATen's avg_pool2d on CPU and CUDA -- then log(clamp(.)) on the packed pipes.
template <int S>
__device__ __forceinline__ void pool_scale(const float (&v)[8][8], const PhaseBArgs& a, float* cc, float* dst, int t,
                                           int by, int bx, int y0, int x0, int H, int W) {
    constexpr int f = 1 << S, nc = 8 >> S;
    constexpr float inv_area = 1.0f / (float)(f * f);
    const int Hs = H >> S, Ws = W >> S;
    float q[nc][nc];
#pragma unroll
    for (int ci = 0; ci < nc; ++ci)
#pragma unroll
        for (int cj = 0; cj < nc; ++cj) {
            float sum = 0.f;
#pragma unroll
            for (int r = 0; r < f; ++r)
#pragma unroll
                for (int c = 0; c < f; ++c) sum += v[ci * f + r][cj * f + c];
            q[ci][cj] = sum * inv_area;      // sum / (f*f): exact power-of-two scaling
        }
    float ql[nc][nc];
    if constexpr (nc == 1) {
        ql[0][0] = log_exact(clamp_nan(q[0][0], a.eps_grad, 1000.0f));
    } else {
#pragma unroll
        for (int ci = 0; ci < nc; ++ci)
#pragma unroll
            for (int cj = 0; cj < nc; cj += 2) {
                const float2 l2 = log_exact2(make_float2(clamp_nan(q[ci][cj], a.eps_grad, 1000.0f),
                                                         clamp_nan(q[ci][cj + 1], a.eps_grad, 1000.0f)));
                ql[ci][cj] = l2.x;
                ql[ci][cj + 1] = l2.y;
            }
    }
#pragma unroll
    for (int ci = 0; ci < nc; ++ci)
#pragma unroll
        for (int cj = 0; cj < nc; ++cj) {
            const int cy = by * nc + ci - (HALO >> S);
            const int cx = bx * nc + cj - (HALO >> S);
            if (cy < -1 || cy > (FTH >> S) || cx < -1 || cx > (FTW >> S)) continue;
            const int gyc = (y0 >> S) + cy, gxc = (x0 >> S) + cx;
            const bool valid = gyc >= 0 && gyc < Hs && gxc >= 0 && gxc < Ws;
            dst[fpool_off(S) + (cy + 1) * fpool_w(S) + (cx + 1)] = valid ? ql[ci][cj] : 0.f;
            if (t == 0 && cy >= 0 && cy < (FTH >> S) && cx >= 0 && cx < (FTW >> S)) {
                const float qq = q[ci][cj];
                const bool cm = (qq >= a.eps_grad) && (qq <= 1000.0f);
                cc[fcc_off(S) + cy * fcc_w(S) + cx] = (valid && cm) ? (1.0f / qq) : 0.f;
            }
        }
}

template <int F>
__global__ void __launch_bounds__(kThreadsB, 2) phase_b_fast_kernel(const PhaseBArgs a) {
    extern __shared__ __align__(16) float smem_raw[];
    __shared__ float s_f[kThreadsB / 32][BF_COUNT];
    __shared__ double s_d[8];
    __shared__ int s_last;
    FastSmem sm;
    sm.sp = smem_raw;
    sm.sg = sm.sp + FRH * FRW;
    sm.pl = sm.sg + FRH * FRW;
    sm.pg = sm.pl + kFPoolCells;
    sm.cc = sm.pg + kFPoolCells;

    constexpr bool GRAD = (F & FB_GRAD) != 0;
    constexpr bool SMOOTH = (F & FB_SMOOTH) != 0;
    constexpr bool SI = (F & FB_SI) != 0;
    constexpr bool RP = (F & FB_RP) != 0;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x;
    const int tx = tile % a.tiles_x;
    const int ty = (tile / a.tiles_x) % a.tiles_y;
    const int b = tile / (a.tiles_x * a.tiles_y);
    const int x0 = tx * FTW, y0 = ty * FTH;
    const int H = a.H, W = a.W;
    const size_t img = (size_t)b * H * W;
    const float* __restrict__ predb = a.pred + img;
    const float* __restrict__ gtb = a.gt ? a.gt + img : nullptr;
    const bool has_mask = a.mask != nullptr;
    const Derived dv = derive(a);

    float acc[BF_COUNT];
#pragma unroll
    for (int q = 0; q < BF_COUNT; ++q) acc[q] = 0.f;

    // ---------------- P1: stage raw pred / gt with an 8-pixel halo, edge pixels replicated ----------------
    if constexpr (GRAD) {
        constexpr int NC4 = FRW / 4;   // 36 float4 per staged row
        for (int i = tid; i < FRH * NC4; i += kThreadsB) {
            const int rr = i / NC4, c4 = i - rr * NC4;
            const int gy = clampi(y0 - HALO + rr, 0, H - 1);
            const int gx = x0 - HALO + 4 * c4;
            float4 pv, gv;
            if (gx >= 0 && gx + 3 < W) {
                pv = __ldg(reinterpret_cast<const float4*>(predb + (size_t)gy * W + gx));
                gv = __ldg(reinterpret_cast<const float4*>(gtb + (size_t)gy * W + gx));
            } else {   // whole float4 outside (W % 4 == 0): replicate the border pixel
                const int cx = gx < 0 ? 0 : W - 1;
                const float ps = __ldg(predb + (size_t)gy * W + cx), gs = __ldg(gtb + (size_t)gy * W + cx);
                pv = make_float4(ps, ps, ps, ps);
                gv = make_float4(gs, gs, gs, gs);
            }
            *reinterpret_cast<float4*>(sm.sp + rr * FRW + 4 * c4) = pv;
            *reinterpret_cast<float4*>(sm.sg + rr * FRW + 4 * c4) = gv;
        }
        __syncthreads();

        // ---------------- P2: avg-pool pyramid in the reference's summation order, pooled logs ----------------
        {
            constexpr int BR = FRH / 8, BC = FRW / 8;   // 8 x 18 blocks of 8x8
            for (int item = tid; item < 2 * BR * BC; item += kThreadsB) {
                const int t = item / (BR * BC);
                const int blk = item - t * (BR * BC);
                const int by = blk / BC, bx = blk - by * BC;
                const float* src = (t == 0 ? sm.sp : sm.sg) + (by * 8) * FRW + bx * 8;
                float* dst = (t == 0 ? sm.pl : sm.pg);
                float v[8][8];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float4 lo = *reinterpret_cast<const float4*>(src + r * FRW);
                    const float4 hi = *reinterpret_cast<const float4*>(src + r * FRW + 4);
                    v[r][0] = lo.x; v[r][1] = lo.y; v[r][2] = lo.z; v[r][3] = lo.w;
                    v[r][4] = hi.x; v[r][5] = hi.y; v[r][6] = hi.z; v[r][7] = hi.w;
                }
                pool_scale<1>(v, a, sm.cc, dst, t, by, bx, y0, x0, H, W);
                pool_scale<2>(v, a, sm.cc, dst, t, by, bx, y0, x0, H, W);
                pool_scale<3>(v, a, sm.cc, dst, t, by, bx, y0, x0, H, W);
            }
        }
        __syncthreads();

        // ---------------- P3a: coefficients of scales 3 and 2; P3b: logs in place over the staged tile ----------------
#pragma unroll
        for (int s = 3; s >= 2; --s) {
            int Hs, Ws; float inv_nx, inv_ny;
            scale_dims(a, s, Hs, Ws, inv_nx, inv_ny);
            const int ch = FTH >> s, cw = FTW >> s, pw = (FTW >> s) + 2;
            const float* PL = sm.pl + fpool_off(s);
            const float* PG = sm.pg + fpool_off(s);
            float* CC = sm.cc + fcc_off(s);
            const float spread = 1.0f / (float)((1 << s) * (1 << s)) * 0.25f * a.w_grad;
            float ax = 0.f, ay = 0.f;
            for (int i = tid; i < ch * cw; i += kThreadsB) {
                const int cy = i / cw, cx = i - cy * cw;
                const int gyc = (y0 >> s) + cy, gxc = (x0 >> s) + cx;
                float coef = 0.f;
                if (gyc < Hs && gxc < Ws) {
                    const int c = (cy + 1) * pw + (cx + 1);
                    const float lp = PL[c], lg = PG[c];
                    float sx_r = 0.f, sx_l = 0.f, sy_d = 0.f, sy_u = 0.f;
                    if (gxc + 1 < Ws) { const float e = (PL[c + 1] - lp) - (PG[c + 1] - lg); ax += fabsf(e); sx_r = sgn3(e); }
                    if (gxc >= 1) sx_l = sgn3((lp - PL[c - 1]) - (lg - PG[c - 1]));
                    if (gyc + 1 < Hs) { const float e = (PL[c + pw] - lp) - (PG[c + pw] - lg); ay += fabsf(e); sy_d = sgn3(e); }
                    if (gyc >= 1) sy_u = sgn3((lp - PL[c - pw]) - (lg - PG[c - pw]));
                    coef = ((sx_l - sx_r) * inv_nx + (sy_u - sy_d) * inv_ny) * CC[i] * spread;
                }
                CC[i] = coef;
            }
            acc[BF_GX0 + 2 * s] += ax;
            acc[BF_GY0 + 2 * s] += ay;
        }
        // logs of the (FTH+2) x (FTW+2) ring+interior, in place: 2 pixels per step on the packed pipes
        {
            constexpr int LR = FTH + 2, LC2 = (FTW + 2) / 2;   // 50 rows x 65 pairs
            for (int i = tid; i < LR * LC2; i += kThreadsB) {
                const int rr = i / LC2, cp = i - rr * LC2;
                const int o = (rr + HALO - 1) * FRW + (HALO - 1) + 2 * cp;
                const float2 pv = make_float2(clamp_nan(sm.sp[o], a.eps_grad, 1000.0f), clamp_nan(sm.sp[o + 1], a.eps_grad, 1000.0f));
                const float2 gv = make_float2(clamp_nan(sm.sg[o], a.eps_grad, 1000.0f), clamp_nan(sm.sg[o + 1], a.eps_grad, 1000.0f));
                const float2 lpv = log_exact2(pv), lgv = log_exact2(gv);        // depth_loss.h:115-116
                sm.sp[o] = lpv.x; sm.sp[o + 1] = lpv.y;
                sm.sg[o] = lgv.x; sm.sg[o + 1] = lgv.y;
            }
        }
        __syncthreads();

        // ---------------- P3c: scale-1 coefficients + the coarser two gathered: what each pixel adds ----------------
        {
            constexpr int s = 1;
            int Hs, Ws; float inv_nx, inv_ny;
            scale_dims(a, s, Hs, Ws, inv_nx, inv_ny);
            constexpr int ch = FTH >> 1, cw = FTW >> 1, pw = (FTW >> 1) + 2;
            const float* PL = sm.pl + fpool_off(1);
            const float* PG = sm.pg + fpool_off(1);
            float* CC = sm.cc + fcc_off(1);
            const float* C2 = sm.cc + fcc_off(2);
            const float* C3 = sm.cc + fcc_off(3);
            const float spread = 0.25f * 0.25f * a.w_grad;
            float ax = 0.f, ay = 0.f;
            for (int i = tid; i < ch * cw; i += kThreadsB) {
                const int cy = i / cw, cx = i - cy * cw;
                const int gyc = (y0 >> 1) + cy, gxc = (x0 >> 1) + cx;
                float coef = 0.f;
                if (gyc < Hs && gxc < Ws) {
                    const int c = (cy + 1) * pw + (cx + 1);
                    const float lp = PL[c], lg = PG[c];
                    float sx_r = 0.f, sx_l = 0.f, sy_d = 0.f, sy_u = 0.f;
                    if (gxc + 1 < Ws) { const float e = (PL[c + 1] - lp) - (PG[c + 1] - lg); ax += fabsf(e); sx_r = sgn3(e); }
                    if (gxc >= 1) sx_l = sgn3((lp - PL[c - 1]) - (lg - PG[c - 1]));
                    if (gyc + 1 < Hs) { const float e = (PL[c + pw] - lp) - (PG[c + pw] - lg); ay += fabsf(e); sy_d = sgn3(e); }
                    if (gyc >= 1) sy_u = sgn3((lp - PL[c - pw]) - (lg - PG[c - pw]));
                    coef = ((sx_l - sx_r) * inv_nx + (sy_u - sy_d) * inv_ny) * CC[i] * spread;
                    coef += C2[(cy >> 1) * fcc_w(2) + (cx >> 1)] + C3[(cy >> 2) * fcc_w(3) + (cx >> 2)];
                }
                CC[i] = coef;
            }
            acc[BF_GX0 + 2] += ax;
            acc[BF_GY0 + 2] += ay;
        }
        __syncthreads();
    }

    // ---------------- P4: full-resolution pass.  One warp = 128 columns, marching down FRPW rows ----------------
    {
        const int xl = 4 * lane;
        const int gx0 = x0 + xl;
        const bool lane_in = gx0 < W;                       // W % 4 == 0: a lane is fully inside or outside
        const int r0 = warp * FRPW;

        // per-column camera geometry (depth_loss.h:290-300)
        float fxe = 1.f, fye = 1.f, rfx = 1.f, rfy = 1.f, cx = 0.f, cy = 0.f;
        float axk[4] = {0.f, 0.f, 0.f, 0.f}, xhk[4] = {0.f, 0.f, 0.f, 0.f};
        bool mk_ok = true;
        if constexpr (RP) {
            float fx, fy;
            load_K(a, b, fx, fy, cx, cy);
            fxe = fx + a.eps_rp;
            fye = fy + a.eps_rp;
            rfx = __frcp_rn(fxe);
            rfy = __frcp_rn(fye);
            mk_ok = markstein_safe(fxe) && markstein_safe(fye);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                axk[k] = (float)(gx0 + k) - cx;
                xhk[k] = __fdiv_rn(axk[k], fxe);
            }
        }
        float inv_nx0 = 0.f, inv_ny0 = 0.f;
        if constexpr (GRAD) { int Hs, Ws; scale_dims(a, 0, Hs, Ws, inv_nx0, inv_ny0); }
        const float g0scale = 0.25f * a.w_grad;             // 1/num_scales * weight
        float ab_w = 0.f, sm_nx = 0.f, sm_ny = 0.f;
        if constexpr (SMOOTH) {
            const float ab = 1.0f / ((float)(a.img_psum[b] / ((double)H * W)) + a.eps_smooth);   // depth_loss.h:192-193
            ab_w = ab * a.w_smooth;
            sm_nx = W > 1 ? (float)(1.0 / ((double)a.global_B * H * (W - 1))) : 0.f;
            sm_ny = H > 1 ? (float)(1.0 / ((double)a.global_B * (H - 1) * W)) : 0.f;
        }
        const float* __restrict__ rgbb = SMOOTH ? a.rgb + (size_t)b * 3 * H * W : nullptr;
        const size_t plane = (size_t)H * W;
        constexpr float kExpScale = -1.4426950408889634f / 3.0f;    // exp(-mean_c|dI|) = 2^(kExpScale * sum_c|dI|)

        // ---- row fetch: raw pred (+ right neighbour), gt, rgb (+ right neighbour) of image row gy (clamped) ----
        struct Row {
            float p[5];        // own 4 + right neighbour
            float g[4];
            float I[3][5];     // own 4 + right neighbour per channel
        };
        auto fetch = [&](int gy_raw, Row& R) {
            const int gy = clampi(gy_raw, 0, H - 1);
            if (lane_in) {
                const float4 p4 = __ldg(reinterpret_cast<const float4*>(predb + (size_t)gy * W + gx0));
                R.p[0] = p4.x; R.p[1] = p4.y; R.p[2] = p4.z; R.p[3] = p4.w;
                if constexpr (SI || RP) {
                    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gtb + (size_t)gy * W + gx0));
                    R.g[0] = g4.x; R.g[1] = g4.y; R.g[2] = g4.z; R.g[3] = g4.w;
                }
            } else {   // lanes right of the image hold the replicated border pixel (their edges vanish)
                const float ps = __ldg(predb + (size_t)gy * W + W - 1);
                R.p[0] = R.p[1] = R.p[2] = R.p[3] = ps;
                if constexpr (SI || RP) R.g[0] = R.g[1] = R.g[2] = R.g[3] = 0.f;
            }
            if constexpr (SMOOTH) {
                const bool in_img = (gy_raw >= 0) && (gy_raw < H);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (in_img && lane_in) v = ldg_stream(reinterpret_cast<const float4*>(rgbb + c * plane + (size_t)gy * W + gx0));
                    R.I[c][0] = v.x; R.I[c][1] = v.y; R.I[c][2] = v.z; R.I[c][3] = v.w;
                }
                // right neighbours: next lane's first pixel; the last lane reads the (edge-replicated) halo pixel
                const int gxn = clampi(gx0 + 4, 0, W - 1);
                float pn = __shfl_down_sync(0xffffffffu, R.p[0], 1);
                if (lane == 31) pn = __ldg(predb + (size_t)gy * W + gxn);
                R.p[4] = pn;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float in = __shfl_down_sync(0xffffffffu, R.I[c][0], 1);
                    if (lane == 31) in = (in_img && gx0 + 4 < W) ? __ldg(rgbb + c * plane + (size_t)gy * W + gx0 + 4) : 0.f;
                    R.I[c][4] = in;
                }
            }
        };

        // state carried from the row above: signed terms of the edge (y-1 -> y) per column
        float sy_up[4] = {0.f, 0.f, 0.f, 0.f};     // gradient matching scale 0: sign(e_y)
        float ty_up[4] = {0.f, 0.f, 0.f, 0.f};     // smoothness: w_y * sign(d_y)
        Row cur, nxt;
        fetch(y0 + r0 - 1, cur);

        for (int r = r0 - 1; r < r0 + FRPW; ++r) {
            const int gy = y0 + r;
            fetch(gy + 1, nxt);
            const bool emit = (r >= r0) && (gy < H);         // warp-uniform
            const bool cnt = emit && lane_in;                // lanes right of the image own no edges
            float out[4] = {0.f, 0.f, 0.f, 0.f};
            float sy_dn[4] = {0.f, 0.f, 0.f, 0.f}, ty_dn[4] = {0.f, 0.f, 0.f, 0.f};
            float lp[4] = {0.f, 0.f, 0.f, 0.f}, lg[4] = {0.f, 0.f, 0.f, 0.f};
            float smg[4] = {0.f, 0.f, 0.f, 0.f};

            if constexpr (GRAD) {
                // logs of rows r and r+1 from shared memory (edge-replicated, so border edges vanish)
                const float* lprow = sm.sp + (r + HALO) * FRW + HALO + xl;
                const float* lgrow = sm.sg + (r + HALO) * FRW + HALO + xl;
                const float4 a4 = *reinterpret_cast<const float4*>(lprow);
                const float4 b4 = *reinterpret_cast<const float4*>(lgrow);
                const float4 ad = *reinterpret_cast<const float4*>(lprow + FRW);
                const float4 bd = *reinterpret_cast<const float4*>(lgrow + FRW);
                lp[0] = a4.x; lp[1] = a4.y; lp[2] = a4.z; lp[3] = a4.w;
                lg[0] = b4.x; lg[1] = b4.y; lg[2] = b4.z; lg[3] = b4.w;
                const float lpd[4] = {ad.x, ad.y, ad.z, ad.w}, lgd[4] = {bd.x, bd.y, bd.z, bd.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float e = (lpd[k] - lp[k]) - (lgd[k] - lg[k]);          // depth_loss.h:151-163
                    sy_dn[k] = sgn3(e);
                    if (cnt) acc[BF_GY0] += fabsf(e);
                }
                if (emit) {
                    const float lpx[6] = {lprow[-1], lp[0], lp[1], lp[2], lp[3], lprow[4]};
                    const float lgx[6] = {lgrow[-1], lg[0], lg[1], lg[2], lg[3], lgrow[4]};
                    float sx[5];
#pragma unroll
                    for (int j = 0; j < 5; ++j) sx[j] = sgn3((lpx[j + 1] - lpx[j]) - (lgx[j + 1] - lgx[j]));   // :140-148,162
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (cnt) acc[BF_GX0] += fabsf((lpx[k + 2] - lpx[k + 1]) - (lgx[k + 2] - lgx[k + 1]));
                        out[k] = ((sx[k] - sx[k + 1]) * inv_nx0 + (sy_up[k] - sy_dn[k]) * inv_ny0) * g0scale;   // x 1/p below
                    }
                }
            }

            if constexpr (SMOOTH) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float s = fabsf(nxt.I[0][k] - cur.I[0][k]) + fabsf(nxt.I[1][k] - cur.I[1][k]) +
                                    fabsf(nxt.I[2][k] - cur.I[2][k]);
                    const float wy = ex2_approx(s * kExpScale);                 // depth_loss.h:218-227
                    const float d = nxt.p[k] - cur.p[k];
                    ty_dn[k] = wy * sgn3(d);
                    if (cnt) acc[BF_SMY] += wy * fabsf(d);
                }
                if (emit) {
                    float tx[5];                                                // tx[j]: edge (x_{j-1} -> x_j), j = 0 is the left neighbour's
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float s = fabsf(cur.I[0][k + 1] - cur.I[0][k]) + fabsf(cur.I[1][k + 1] - cur.I[1][k]) +
                                        fabsf(cur.I[2][k + 1] - cur.I[2][k]);
                        const float wx = ex2_approx(s * kExpScale);             // depth_loss.h:211-226
                        const float d = cur.p[k + 1] - cur.p[k];
                        tx[k + 1] = wx * sgn3(d);
                        if (cnt) acc[BF_SMX] += wx * fabsf(d);
                    }
                    float tl = __shfl_up_sync(0xffffffffu, tx[4], 1);
                    if (lane == 0) {
                        // left neighbour lives in another tile: evaluate that one edge here
                        tl = 0.f;
                        if (gx0 >= 1) {
                            const float pl_ = __ldg(predb + (size_t)gy * W + gx0 - 1);
                            float s = 0.f;
#pragma unroll
                            for (int c = 0; c < 3; ++c) s += fabsf(cur.I[c][0] - __ldg(rgbb + c * plane + (size_t)gy * W + gx0 - 1));
                            tl = ex2_approx(s * kExpScale) * sgn3(cur.p[0] - pl_);
                        }
                    }
                    tx[0] = tl;
#pragma unroll
                    for (int k = 0; k < 4; ++k)     // d L / d p_j without the mean-normalisation term (added per image later)
                        smg[k] = ((tx[k] - tx[k + 1]) * sm_nx + (ty_up[k] - ty_dn[k]) * sm_ny) * ab_w;
                }
            }

            if (emit) {
                // pointwise terms + assembly
                uchar4 mk = make_uchar4(1, 1, 1, 1);
                if (has_mask && lane_in) mk = __ldg(reinterpret_cast<const uchar4*>(a.mask + img + (size_t)gy * W + gx0));
                const unsigned char um[4] = {mk.x, mk.y, mk.z, mk.w};
                float ayv = 0.f, yh = 0.f;
                if constexpr (RP) {
                    ayv = (float)gy - cy;
                    yh = __fdiv_rn(ayv, fye);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float p = cur.p[k];
                    const float rp = rcp_approx(p);
                    float gsum = 0.f;
                    if constexpr (GRAD) {
                        const bool cm = (p >= a.eps_grad) && (p <= 1000.0f);      // clamp backward
                        gsum = cm ? out[k] * rp : 0.f;
                        gsum += sm.cc[fcc_off(1) + (r >> 1) * fcc_w(1) + ((xl + k) >> 1)];
                    }
                    if constexpr (SMOOTH) gsum += smg[k];
                    if constexpr (SI) {
                        const float g = cur.g[k];
                        const bool m = has_mask ? (um[k] != 0) : (g > a.eps_si);
                        const bool cm = (p >= a.eps_si) && (p <= 1000.0f);
                        float d;
                        if constexpr (GRAD) d = lp[k] - lg[k];
                        else d = log_exact(clamp_nan(p, a.eps_si, 1000.0f)) - log_exact(clamp_nan(g, a.eps_si, 1000.0f));
                        if (m && cm && dv.si_on) gsum += a.w_si * ((dv.si_c1 * d + dv.si_c2) * rp);
                    }
                    if constexpr (RP) {
                        const float g = cur.g[k];
                        const bool m = has_mask ? (um[k] != 0) : (g > a.eps_rp);
                        if (m && dv.rp_on && lane_in) {
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
                            const float ss = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(dX, dX), __fmul_rn(dY, dY)), __fmul_rn(dZ, dZ)), a.eps_rp);
                            const float re = rsqrt_approx(ss);
                            acc[BF_RP_E] += ss * re;                               // e = sqrt(ss)
                            gsum += a.w_rp * ((dX * xhk[k] + dY * yh + dZ) * re * dv.rp_inv_n);
                        }
                    }
                    out[k] = gsum * a.upstream;
                }
                if (a.grad && lane_in)
                    *reinterpret_cast<float4*>(a.grad + img + (size_t)gy * W + gx0) = make_float4(out[0], out[1], out[2], out[3]);
            }

            // roll the row state
#pragma unroll
            for (int k = 0; k < 4; ++k) { sy_up[k] = sy_dn[k]; ty_up[k] = ty_dn[k]; }
            cur = nxt;
        }
    }

    if (publish_partials(a, acc, tile, s_f, &s_last)) {
        finalize_results(a, s_d);
        if (tid == 0 && a.metrics) write_metric_results(a.stats, a.metrics, *a.results);
    }
}

}  // namespace cadl
