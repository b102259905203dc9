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
#include <cuda.h>   // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cooperative_groups.h>
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
// INTERIOR blocks store all their cells unconditionally; halo blocks only the cells next to the tile.
template <int S, bool INTERIOR>
__device__ __forceinline__ void pool_scale(const float (&v)[8][8], float eps, float* cc, float* dst, int t, int by, int bx) {
    constexpr int f = 1 << S, nc = 8 >> S;
    constexpr float inv_area = 1.0f / (float)(f * f);
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
    // local cell coordinates of this block's first cell (the staged region starts one coarsest cell before the tile)
    const int cy0 = by * nc - nc, cx0 = bx * nc - nc;
    if constexpr (INTERIOR) {
        float* drow = dst + fpool_off(S) + (cy0 + 1) * fpool_w(S) + (cx0 + 1);
        float* crow = cc + fcc_off(S) + cy0 * fcc_w(S) + cx0;
#pragma unroll
        for (int ci = 0; ci < nc; ++ci) {
            if constexpr (nc == 1) {
                drow[0] = log_exact(clamp_nan(q[0][0], eps, 1000.0f));
            } else {
#pragma unroll
                for (int cj = 0; cj < nc; cj += 2) {
                    const float2 l2 = log_exact2(make_float2(clamp_nan(q[ci][cj], eps, 1000.0f), clamp_nan(q[ci][cj + 1], eps, 1000.0f)));
                    drow[ci * fpool_w(S) + cj] = l2.x;       // (odd float index: the array has a 1-cell halo)
                    drow[ci * fpool_w(S) + cj + 1] = l2.y;
                }
            }
            if (t == 0) {
#pragma unroll
                for (int cj = 0; cj < nc; ++cj)
                    crow[ci * fcc_w(S) + cj] = in_range_pos(q[ci][cj], eps, 1000.0f) ? rcp_approx(q[ci][cj]) : 0.f;
            }
        }
    } else {
#pragma unroll
        for (int ci = 0; ci < nc; ++ci)
#pragma unroll
            for (int cj = 0; cj < nc; ++cj) {
                const int cy = cy0 + ci, cx = cx0 + cj;
                if (cy < -1 || cy > (FTH >> S) || cx < -1 || cx > (FTW >> S)) continue;
                dst[fpool_off(S) + (cy + 1) * fpool_w(S) + (cx + 1)] = log_exact(clamp_nan(q[ci][cj], eps, 1000.0f));
            }
    }
}

// Border tiles: copy the pooled logs of the last valid row / column of cells into the first row / column
// outside the image, so that an edge across the image border has residual exactly 0 (no per-edge predicates).
template <int S>
__device__ __forceinline__ void pooled_replicate(const PhaseBArgs& a, const FastSmem& sm, int tid, int y0, int x0, bool rows) {
    constexpr int ch = FTH >> S, cw = FTW >> S, pw = cw + 2, ph = ch + 2;
    const int Hs = a.H >> S, Ws = a.W >> S;
    float* PL = sm.pl + fpool_off(S);
    float* PG = sm.pg + fpool_off(S);
    if (rows) {
        const int top = (y0 == 0) ? 0 : -1;                                  // array row of cells at image row -1
        const int bot = (Hs - (y0 >> S) <= ch) ? Hs - (y0 >> S) + 1 : -1;    // array row of cells at image row Hs
        for (int i = tid; i < pw; i += kThreadsB) {
            if (top >= 0) { PL[top * pw + i] = PL[(top + 1) * pw + i]; PG[top * pw + i] = PG[(top + 1) * pw + i]; }
            if (bot >= 0) { PL[bot * pw + i] = PL[(bot - 1) * pw + i]; PG[bot * pw + i] = PG[(bot - 1) * pw + i]; }
        }
    } else {
        const int lft = (x0 == 0) ? 0 : -1;
        const int rgt = (Ws - (x0 >> S) <= cw) ? Ws - (x0 >> S) + 1 : -1;
        for (int i = tid; i < ph; i += kThreadsB) {
            if (lft >= 0) { PL[i * pw + lft] = PL[i * pw + lft + 1]; PG[i * pw + lft] = PG[i * pw + lft + 1]; }
            if (rgt >= 0) { PL[i * pw + rgt] = PL[i * pw + rgt - 1]; PG[i * pw + rgt] = PG[i * pw + rgt - 1]; }
        }
    }
}

// One scale's per-cell coefficient pass (branch-free): every cell evaluates its four edges, the two it owns
// also feed the loss sum; stores coefficient * (1/q) * spread [+ the two coarser scales when GATHER].
// Cells outside the image hold 1/q = 0, and edges across the border vanish by replication (above).
template <int S, bool GATHER>
__device__ __forceinline__ void coef_pass(const PhaseBArgs& a, const FastSmem& sm, int tid, int y0, int x0, float spread,
                                          float& acc_x, float& acc_y) {
    constexpr int ch = FTH >> S, cw = FTW >> S, pw = (FTW >> S) + 2;
    const int Hs = a.H >> S, Ws = a.W >> S;
    const float inv_nx = a.inv_nx[S] * spread, inv_ny = a.inv_ny[S] * spread;
    const float* PL = sm.pl + fpool_off(S);
    const float* PG = sm.pg + fpool_off(S);
    float* CC = sm.cc + fcc_off(S);
    const float* C2 = sm.cc + fcc_off(2);
    const float* C3 = sm.cc + fcc_off(3);
    float ax = 0.f, ay = 0.f;
#pragma unroll 2
    for (int i = tid; i < ch * cw; i += kThreadsB) {
        const int cy = i / cw, cx = i - cy * cw;                  // cw is a power of two
        const bool cv = ((y0 >> S) + cy < Hs) && ((x0 >> S) + cx < Ws);
        const int c = (cy + 1) * pw + (cx + 1);
        const float lp = PL[c], lg = PG[c];
        const float e_r = (PL[c + 1] - lp) - (PG[c + 1] - lg);    // depth_loss.h:140-148,162
        const float e_l = (lp - PL[c - 1]) - (lg - PG[c - 1]);
        const float e_d = (PL[c + pw] - lp) - (PG[c + pw] - lg);  // :151-159,163
        const float e_u = (lp - PL[c - pw]) - (lg - PG[c - pw]);
        if (cv) { ax += fabsf(e_r); ay += fabsf(e_d); }
        float coef = ((sgn3(e_l) - sgn3(e_r)) * inv_nx + (sgn3(e_u) - sgn3(e_d)) * inv_ny) * CC[i];
        if constexpr (GATHER) coef += C2[(cy >> 1) * fcc_w(2) + (cx >> 1)] + C3[(cy >> 2) * fcc_w(3) + (cx >> 2)];
        CC[i] = coef;
    }
    acc_x += ax;
    acc_y += ay;
}

// Thread-group barrier: the whole CTA (BAR == 0) or a named barrier over kThreadsB threads (warp-specialised kernel)
template <int BAR>
__device__ __forceinline__ void gsync() {
    if constexpr (BAR == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" :: "n"(BAR), "n"(kThreadsB) : "memory");
}

// Everything between "the raw pred/gt tile is in shared memory" and the full-resolution pass: border ring,
// avg-pool pyramid, pooled logs, per-cell coefficients of the three coarse scales, logs in place.
// Executed by a group of kThreadsB threads (tid = index inside the group).
template <int BAR>
__device__ __forceinline__ void fast_prelude(const PhaseBArgs& a, const FastSmem& sm, int tid, int y0, int x0,
                                             bool zero_filled, float (&acc)[BF_COUNT]) {
    const int H = a.H, W = a.W;
    if (zero_filled) {   // TMA staging: patch the replicated 1-pixel ring of border tiles
        const int r_top = (y0 == 0) ? HALO - 1 : -1;                       // staged row of image row -1
        const int r_bot = (H - y0 + HALO < FRH) ? H - y0 + HALO : -1;      // staged row of image row H
        const int c_lft = (x0 == 0) ? HALO - 1 : -1;
        const int c_rgt = (W - x0 + HALO < FRW) ? W - x0 + HALO : -1;
        if (r_top >= 0 || r_bot >= 0 || c_lft >= 0 || c_rgt >= 0) {       // block-uniform: the tile touches the border
            if (r_top >= 0 && tid < FRW) { sm.sp[r_top * FRW + tid] = sm.sp[(r_top + 1) * FRW + tid]; sm.sg[r_top * FRW + tid] = sm.sg[(r_top + 1) * FRW + tid]; }
            if (r_bot >= 0 && tid < FRW) { sm.sp[r_bot * FRW + tid] = sm.sp[(r_bot - 1) * FRW + tid]; sm.sg[r_bot * FRW + tid] = sm.sg[(r_bot - 1) * FRW + tid]; }
            gsync<BAR>();                                               // rows first, then columns (corners follow)
            if (c_lft >= 0 && tid < FRH) { sm.sp[tid * FRW + c_lft] = sm.sp[tid * FRW + c_lft + 1]; sm.sg[tid * FRW + c_lft] = sm.sg[tid * FRW + c_lft + 1]; }
            if (c_rgt >= 0 && tid < FRH) { sm.sp[tid * FRW + c_rgt] = sm.sp[tid * FRW + c_rgt - 1]; sm.sg[tid * FRW + c_rgt] = sm.sg[tid * FRW + c_rgt - 1]; }
        }
    }
    gsync<BAR>();

    // ---------------- P2: avg-pool pyramid in the reference's summation order, pooled logs ----------------
    {
        constexpr int BR = FRH / 8, BC = FRW / 8;   // 8 x 18 blocks of 8x8
        // items 0..191: interior blocks (6 x 16 per tensor) -- six whole warps on the unconditional path;
        // items 192..287: the ring of halo blocks (48 per tensor)
        constexpr int NI = (BR - 2) * (BC - 2), NH = BR * BC - NI;
        for (int item = tid; item < 2 * BR * BC; item += kThreadsB) {
            int t, by, bx;
            if (item < 2 * NI) {
                t = item >= NI ? 1 : 0;
                const int idx = item - t * NI;
                by = 1 + idx / (BC - 2);
                bx = 1 + idx - (by - 1) * (BC - 2);
            } else {
                const int h = item - 2 * NI;
                t = h >= NH ? 1 : 0;
                const int idx = h - t * NH;
                if (idx < BC) { by = 0; bx = idx; }
                else if (idx < 2 * BC) { by = BR - 1; bx = idx - BC; }
                else if (idx < 2 * BC + (BR - 2)) { by = 1 + idx - 2 * BC; bx = 0; }
                else { by = 1 + idx - 2 * BC - (BR - 2); bx = BC - 1; }
            }
            float* dst = (t == 0 ? sm.pl : sm.pg);
            // H, W and the block origin are multiples of 8: a block is entirely inside or outside the image
            const int gby = y0 - HALO + 8 * by, gbx = x0 - HALO + 8 * bx;
            const bool valid = gby >= 0 && gby < H && gbx >= 0 && gbx < W;
            const bool interior = by >= 1 && by <= BR - 2 && bx >= 1 && bx <= BC - 2;
            if (!valid) {   // cells outside the image: defined but inert (1/q = 0; logs patched by pooled_replicate)
#pragma unroll
                for (int S = 1; S <= 3; ++S) {
                    const int nc = 8 >> S, cy0 = by * nc - nc, cx0 = bx * nc - nc;
                    for (int ci = 0; ci < nc; ++ci)
                        for (int cj = 0; cj < nc; ++cj) {
                            const int cy = cy0 + ci, cx = cx0 + cj;
                            if (cy < -1 || cy > (FTH >> S) || cx < -1 || cx > (FTW >> S)) continue;
                            dst[fpool_off(S) + (cy + 1) * fpool_w(S) + (cx + 1)] = 0.f;
                            if (t == 0 && cy >= 0 && cy < (FTH >> S) && cx >= 0 && cx < (FTW >> S))
                                sm.cc[fcc_off(S) + cy * fcc_w(S) + cx] = 0.f;
                        }
                }
                continue;
            }
            const float* src = (t == 0 ? sm.sp : sm.sg) + (by * 8) * FRW + bx * 8;
            float v[8][8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float4 lo = *reinterpret_cast<const float4*>(src + r * FRW);
                const float4 hi = *reinterpret_cast<const float4*>(src + r * FRW + 4);
                v[r][0] = lo.x; v[r][1] = lo.y; v[r][2] = lo.z; v[r][3] = lo.w;
                v[r][4] = hi.x; v[r][5] = hi.y; v[r][6] = hi.z; v[r][7] = hi.w;
            }
            if (interior) {
                pool_scale<1, true>(v, a.eps_grad, sm.cc, dst, t, by, bx);
                pool_scale<2, true>(v, a.eps_grad, sm.cc, dst, t, by, bx);
                pool_scale<3, true>(v, a.eps_grad, sm.cc, dst, t, by, bx);
            } else {
                pool_scale<1, false>(v, a.eps_grad, sm.cc, dst, t, by, bx);
                pool_scale<2, false>(v, a.eps_grad, sm.cc, dst, t, by, bx);
                pool_scale<3, false>(v, a.eps_grad, sm.cc, dst, t, by, bx);
            }
        }
    }
    gsync<BAR>();
    if (y0 == 0 || x0 == 0 || H - y0 <= FTH || W - x0 <= FTW) {   // block-uniform: the tile touches the image border
        pooled_replicate<1>(a, sm, tid, y0, x0, true);
        pooled_replicate<2>(a, sm, tid, y0, x0, true);
        pooled_replicate<3>(a, sm, tid, y0, x0, true);
        gsync<BAR>();
        pooled_replicate<1>(a, sm, tid, y0, x0, false);
        pooled_replicate<2>(a, sm, tid, y0, x0, false);
        pooled_replicate<3>(a, sm, tid, y0, x0, false);
        gsync<BAR>();
    }

    // ---------------- P3a: coefficients of scales 3 and 2 ----------------
    const float wg = 0.25f * a.w_grad * a.upstream;      // 1/num_scales * weight * upstream
    coef_pass<3, false>(a, sm, tid, y0, x0, wg * (1.0f / 64.0f), acc[BF_GX3], acc[BF_GY3]);
    coef_pass<2, false>(a, sm, tid, y0, x0, wg * (1.0f / 16.0f), acc[BF_GX2], acc[BF_GY2]);
    // ---------------- P3b: logs of the (FTH+2) x (FTW+2) ring + interior, in place, 2 px per step ----------------
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
    gsync<BAR>();
    // ---------------- P3c: scale 1 + the two coarser gathered: what each pixel adds ----------------
    coef_pass<1, true>(a, sm, tid, y0, x0, wg * 0.25f, acc[BF_GX1], acc[BF_GY1]);
}

// The full-resolution pass over one staged tile, executed by 8 warps (warp = index inside the group).
template <int F, bool HAS_MASK>
__device__ __forceinline__ void fast_p4(const PhaseBArgs& a, const FastSmem& sm, int warp, int lane, int b, int y0, int x0,
                                        const float* s_c, float (&acc)[BF_COUNT]) {
    constexpr bool GRAD = (F & FB_GRAD) != 0;
    constexpr bool SMOOTH = (F & FB_SMOOTH) != 0;
    constexpr bool SI = (F & FB_SI) != 0;
    constexpr bool RP = (F & FB_RP) != 0;
    const int H = a.H, W = a.W;
    const int img = b * H * W;                               // B*H*W < 2^31 (checked on the host)
    const float* __restrict__ predb = a.pred + img;
    const float* __restrict__ gtb = a.gt ? a.gt + img : nullptr;
    const float up = a.upstream;
    // one warp = 128 columns, marching down FRPW rows; see file header
    // ---------------- P4: full-resolution pass.  One warp = 128 columns, marching down FRPW rows ----------------
    {
        const int xl = 4 * lane;
        const int gx0 = x0 + xl;
        const bool lane_in = gx0 < W;                       // W % 4 == 0: a lane is fully inside or outside
        const int r0 = warp * FRPW;
        const float c1 = s_c[0], c2 = s_c[1], rpn = s_c[2], abw = s_c[3];

        // per-column camera geometry (depth_loss.h:290-300)
        float fxe = 1.f, fye = 1.f, rfx = 1.f, rfy = 1.f, cyv = 0.f;
        float axk[4] = {0.f, 0.f, 0.f, 0.f}, xhk[4] = {0.f, 0.f, 0.f, 0.f};
        bool mk_ok = true;
        if constexpr (RP) {
            float fx, fy, cxv;
            load_K(a, b, fx, fy, cxv, cyv);
            fxe = fx + a.eps_rp;
            fye = fy + a.eps_rp;
            rfx = __frcp_rn(fxe);
            rfy = __frcp_rn(fye);
            mk_ok = markstein_safe(fxe) && markstein_safe(fye);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                axk[k] = (float)(gx0 + k) - cxv;
                xhk[k] = axk[k] * rfx;                       // d pX / d p: tolerance path
            }
        }
        const float inx0 = a.inv_nx[0] * 0.25f * a.w_grad * up, iny0 = a.inv_ny[0] * 0.25f * a.w_grad * up;
        const float snx = a.sm_nx * abw, sny = a.sm_ny * abw;
        const float eps_g = a.eps_grad, eps_s = a.eps_si, eps_r = a.eps_rp;
        const float* __restrict__ rgbb = SMOOTH ? a.rgb + (size_t)b * 3 * H * W : nullptr;
        const int plane = H * W;
        constexpr float kExpScale = -1.4426950408889634f / 3.0f;    // exp(-mean_c|dI|) = 2^(kExpScale * sum_c|dI|)
        const int gxr = clampi(gx0 + 4, 0, W - 1);                  // right neighbour column of the last lane
        const bool right_in = gx0 + 4 < W;
        const bool endlane = (lane == 31) || (lane == 0 && gx0 >= 1);
        const int hx = (lane == 31) ? gxr : (gx0 >= 1 ? gx0 - 1 : 0);     // column of the end lanes' halo pixel
        const bool h_rgb_ok = (lane == 31) ? right_in : true;

        // ---- row state ----
        float pc[5], Ic[3][5];              // current row: own 4 + right neighbour
        float pn[5], In[3][5];              // next row (index 4 filled by finish_row once the loads have landed)
        float hn_p = 0.f, hn_I[3] = {0.f, 0.f, 0.f};   // end lanes' halo pixel of the next row
        float hc_p = 0.f, hc_I[3] = {0.f, 0.f, 0.f};   // ... of the current row
        float lpc[4] = {0.f, 0.f, 0.f, 0.f}, lgc[4] = {0.f, 0.f, 0.f, 0.f};   // logs of the current row
        float sy_up[4] = {0.f, 0.f, 0.f, 0.f}, ty_up[4] = {0.f, 0.f, 0.f, 0.f};

        // issue the global loads of one image row (clamped at the borders); no use of the values here
        auto fetch = [&](int gy_raw, float (&p)[5], float (&I)[3][5], float& hp, float (&hI)[3]) {
            const bool in_img = (gy_raw >= 0) && (gy_raw < H);
            const int ro = clampi(gy_raw, 0, H - 1) * W;
            if (lane_in) {
                const float4 p4 = __ldg(reinterpret_cast<const float4*>(predb + ro + gx0));
                p[0] = p4.x; p[1] = p4.y; p[2] = p4.z; p[3] = p4.w;
            } else {   // lanes right of the image hold the replicated border pixel (their edges vanish)
                const float ps = __ldg(predb + ro + W - 1);
                p[0] = p[1] = p[2] = p[3] = ps;
            }
            if constexpr (SMOOTH) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (in_img && lane_in) v = ldg_stream(reinterpret_cast<const float4*>(rgbb + c * plane + ro + gx0));
                    I[c][0] = v.x; I[c][1] = v.y; I[c][2] = v.z; I[c][3] = v.w;
                }
                // the two lanes at the warp's ends also fetch the pixel beyond their end: lane 31 its right
                // neighbour, lane 0 its left neighbour (same registers, same instructions, different lanes)
                if (endlane) {
                    hp = __ldg(predb + ro + hx);
                    const bool ok = in_img && h_rgb_ok;
#pragma unroll
                    for (int c = 0; c < 3; ++c) hI[c] = ok ? __ldg(rgbb + c * plane + ro + hx) : 0.f;
                }
            }
        };
        // right neighbours across lanes -- called when the row becomes current, long after its loads were issued
        auto finish_row = [&](float (&p)[5], float (&I)[3][5], float hp, const float (&hI)[3]) {
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
        // terms of the vertical edges (row -> row+1): signed gradient-matching residual and smoothness term
        auto yterms = [&](int r, bool count, float (&sy)[4], float (&ty)[4], float (&lpn)[4], float (&lgn)[4]) {
            if constexpr (GRAD) {
                const float4 ad = *reinterpret_cast<const float4*>(sm.sp + (r + 1 + HALO) * FRW + HALO + xl);
                const float4 bd = *reinterpret_cast<const float4*>(sm.sg + (r + 1 + HALO) * FRW + HALO + xl);
                lpn[0] = ad.x; lpn[1] = ad.y; lpn[2] = ad.z; lpn[3] = ad.w;
                lgn[0] = bd.x; lgn[1] = bd.y; lgn[2] = bd.z; lgn[3] = bd.w;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float e = (lpn[k] - lpc[k]) - (lgn[k] - lgc[k]);      // depth_loss.h:151-163
                    sy[k] = sgn3(e);
                    if (count) acc[BF_GY0] += fabsf(e);
                }
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

        // prologue: the row above this warp's first row only contributes its lower edges
        {
            float hp0 = 0.f, hI0[3] = {0.f, 0.f, 0.f};
            fetch(y0 + r0 - 1, pc, Ic, hp0, hI0);
            fetch(y0 + r0, pn, In, hn_p, hn_I);
        }
        if constexpr (GRAD) {
            const float4 a4 = *reinterpret_cast<const float4*>(sm.sp + (r0 - 1 + HALO) * FRW + HALO + xl);
            const float4 b4 = *reinterpret_cast<const float4*>(sm.sg + (r0 - 1 + HALO) * FRW + HALO + xl);
            lpc[0] = a4.x; lpc[1] = a4.y; lpc[2] = a4.z; lpc[3] = a4.w;
            lgc[0] = b4.x; lgc[1] = b4.y; lgc[2] = b4.z; lgc[3] = b4.w;
        }
        {
            float lpn[4], lgn[4];
            yterms(r0 - 1, false, sy_up, ty_up, lpn, lgn);
            if constexpr (GRAD) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { lpc[k] = lpn[k]; lgc[k] = lgn[k]; }
            }
        }
        finish_row(pn, In, hn_p, hn_I);
        hc_p = hn_p;
#pragma unroll
        for (int c = 0; c < 3; ++c) hc_I[c] = hn_I[c];
#pragma unroll
        for (int k = 0; k < 5; ++k) pc[k] = pn[k];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int k = 0; k < 5; ++k) Ic[c][k] = In[c][k];

        for (int r = r0; r < r0 + FRPW; ++r) {
            const int gy = y0 + r;
            if (gy >= H) break;                              // warp-uniform
            // 1. issue next row's loads; they are consumed at step 4, after ~2/3 of this row's arithmetic
            fetch(gy + 1, pn, In, hn_p, hn_I);
            float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (SI || RP) {
                if (lane_in) g4 = __ldg(reinterpret_cast<const float4*>(gtb + gy * W + gx0));
            }
            uchar4 mk4 = make_uchar4(0, 0, 0, 0);
            if constexpr (HAS_MASK) {
                if (lane_in) mk4 = __ldg(reinterpret_cast<const uchar4*>(a.mask + img + gy * W + gx0));
            }

            // 2. horizontal edges of the current row
            float gm[4] = {0.f, 0.f, 0.f, 0.f}, smg[4] = {0.f, 0.f, 0.f, 0.f};
            if constexpr (GRAD) {
                const float* lprow = sm.sp + (r + HALO) * FRW + HALO + xl;
                const float* lgrow = sm.sg + (r + HALO) * FRW + HALO + xl;
                const float lpx[6] = {lprow[-1], lpc[0], lpc[1], lpc[2], lpc[3], lprow[4]};
                const float lgx[6] = {lgrow[-1], lgc[0], lgc[1], lgc[2], lgc[3], lgrow[4]};
                float sx[5];
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const float e = (lpx[j + 1] - lpx[j]) - (lgx[j + 1] - lgx[j]);   // depth_loss.h:140-148,162
                    sx[j] = sgn3(e);
                    if (j >= 1 && lane_in) acc[BF_GX0] += fabsf(e);                  // the 4 edges this lane owns
                }
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
                    tl = 0.f;                                 // left neighbour lives in another tile: evaluate that edge here
                    if (gx0 >= 1) {                           // (its pixel was fetched with the row: hc_p, hc_I)
                        const float s = fabsf(Ic[0][0] - hc_I[0]) + fabsf(Ic[1][0] - hc_I[1]) + fabsf(Ic[2][0] - hc_I[2]);
                        tl = ex2_approx(s * kExpScale) * sgn3(pc[0] - hc_p);
                    }
                }
                tx[0] = tl;
#pragma unroll
                for (int k = 0; k < 4; ++k) smg[k] = (tx[k] - tx[k + 1]) * snx;
            }

            // 3. pointwise terms
            const float gcur[4] = {g4.x, g4.y, g4.z, g4.w};
            const bool um[4] = {mk4.x != 0, mk4.y != 0, mk4.z != 0, mk4.w != 0};
            float ayv = 0.f, yh = 0.f;
            if constexpr (RP) {
                ayv = (float)gy - cyv;
                yh = ayv * rfy;                               // d pY / d p (tolerance path)
            }
            const float2 ccv = GRAD ? *reinterpret_cast<const float2*>(sm.cc + fcc_off(1) + (r >> 1) * fcc_w(1) + (xl >> 1))
                                    : make_float2(0.f, 0.f);
            float rpk[4], pw[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float p = pc[k];
                rpk[k] = rcp_approx(p);
                float gsum = (k < 2 ? ccv.x : ccv.y);
                if constexpr (SI) {
                    const float g = gcur[k];
                    const bool m = HAS_MASK ? um[k] : (g > eps_s);
                    float d;
                    if constexpr (GRAD) d = lpc[k] - lgc[k];
                    else d = log_exact(clamp_nan(p, eps_s, 1000.0f)) - log_exact(clamp_nan(g, eps_s, 1000.0f));
                    if (m && in_range_pos(p, eps_s, 1000.0f)) gsum = fmaf(fmaf(c1, d, c2), rpk[k], gsum);
                }
                if constexpr (RP) {
                    const float g = gcur[k];
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

            // 4. vertical edges (needs the next row), then assembly and the 128-bit store
            float sy_dn[4] = {0.f, 0.f, 0.f, 0.f}, ty_dn[4] = {0.f, 0.f, 0.f, 0.f};
            float lpn[4] = {0.f, 0.f, 0.f, 0.f}, lgn[4] = {0.f, 0.f, 0.f, 0.f};
            yterms(r, lane_in, sy_dn, ty_dn, lpn, lgn);
            float out[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float gsum = pw[k];
                if constexpr (SMOOTH) gsum += fmaf(ty_up[k] - ty_dn[k], sny, smg[k]);
                if constexpr (GRAD) {
                    const float gmk = fmaf(sy_up[k] - sy_dn[k], iny0, gm[k]);
                    gsum = in_range_pos(pc[k], eps_g, 1000.0f) ? fmaf(gmk, rpk[k], gsum) : gsum;   // clamp backward
                }
                out[k] = gsum;
            }
            if (a.grad && lane_in)
                *reinterpret_cast<float4*>(a.grad + img + gy * W + gx0) = make_float4(out[0], out[1], out[2], out[3]);

            // roll the row state
            finish_row(pn, In, hn_p, hn_I);
            hc_p = hn_p;
#pragma unroll
            for (int c = 0; c < 3; ++c) hc_I[c] = hn_I[c];
#pragma unroll
            for (int k = 0; k < 4; ++k) { sy_up[k] = sy_dn[k]; ty_up[k] = ty_dn[k]; lpc[k] = lpn[k]; lgc[k] = lgn[k]; }
#pragma unroll
            for (int k = 0; k < 5; ++k) pc[k] = pn[k];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int k = 0; k < 5; ++k) Ic[c][k] = In[c][k];
        }
    }

}

template <int F, bool HAS_MASK>
__global__ void __launch_bounds__(kThreadsB, 2) phase_b_fast_kernel(const PhaseBArgs a,
                                                                    const __grid_constant__ CUtensorMap tm_pred,
                                                                    const __grid_constant__ CUtensorMap tm_gt) {
    extern __shared__ __align__(128) float smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ float s_f[kThreadsB / 32][BF_COUNT];
    __shared__ double s_d[8];
    __shared__ float s_c[8];
    __shared__ int s_last;
    FastSmem sm;
    sm.sp = smem_raw;
    sm.sg = sm.sp + FRH * FRW;
    sm.pl = sm.sg + FRH * FRW;
    sm.pg = sm.pl + kFPoolCells;
    sm.cc = sm.pg + kFPoolCells;

    constexpr bool GRAD = (F & FB_GRAD) != 0;
    constexpr bool SMOOTH = (F & FB_SMOOTH) != 0;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x;
    const int tx = tile % a.tiles_x;
    const int ty = (tile / a.tiles_x) % a.tiles_y;
    const int b = tile / (a.tiles_x * a.tiles_y);
    const int x0 = tx * FTW, y0 = ty * FTH;
    const int H = a.H, W = a.W;
    const int img = b * H * W;                               // B*H*W < 2^31 (checked on the host)
    const float* __restrict__ predb = a.pred + img;
    const float* __restrict__ gtb = a.gt ? a.gt + img : nullptr;
    const float up = a.upstream;

    // Scalars every pixel needs, derived once per CTA from the phase-A statistics;
    // weights and the upstream gradient are folded in here so the pixel loop has no extra multiplies.
    if (tid == 0) {
        const double n = a.stats[ST_SI_N], S = a.stats[ST_SI_S], nr = a.stats[ST_RP_N];
        s_c[0] = n > 0.0 ? (float)(2.0 / n) * a.w_si * up : 0.f;                                   // c1
        s_c[1] = n > 0.0 ? (float)(-2.0 * (double)a.lambda * S / (n * n)) * a.w_si * up : 0.f;     // c2
        s_c[2] = nr > 0.0 ? (float)(1.0 / nr) * a.w_rp * up : 0.f;                                 // 1/n (reprojection)
        s_c[3] = SMOOTH ? (1.0f / ((float)(a.img_psum[b] / ((double)H * W)) + a.eps_smooth)) * a.w_smooth * up : 0.f;  // a_b (:192-193)
    }

    float acc[BF_COUNT];
#pragma unroll
    for (int q = 0; q < BF_COUNT; ++q) acc[q] = 0.f;

    if constexpr (GRAD) {
        // ---------------- P1: stage raw pred / gt with an 8-pixel halo ----------------
        if (a.use_tma) {
            // TMA: one thread issues two 3-D box loads (144 x 64 x 1 floats each); pixels outside the image
            // arrive as zeros.  Only the 1-pixel ring around the image needs the replicated border value
            // (pooled cells outside the image are invalid anyway), so border tiles patch one row / column.
            if (tid == 0) mbar_init(&s_bar, 1);
            __syncthreads();
            if (tid == 0) {
                mbar_expect_tx(&s_bar, 2u * FRH * FRW * sizeof(float));
                tma_load_3d(sm.sp, &tm_pred, x0 - HALO, y0 - HALO, b, &s_bar);
                tma_load_3d(sm.sg, &tm_gt, x0 - HALO, y0 - HALO, b, &s_bar);
            }
            mbar_wait(&s_bar, 0);
        } else {
            // cp.async fallback (no tensor maps): 16-byte copies, border pixels replicated
            for (int rr = warp; rr < FRH; rr += kThreadsB / 32) {
                const int rowo = clampi(y0 - HALO + rr, 0, H - 1) * W;
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    const int c4 = pass * 32 + lane;
                    if (c4 < FRW / 4) {
                        const int gx = x0 - HALO + 4 * c4;
                        float* dp = sm.sp + rr * FRW + 4 * c4;
                        float* dg = sm.sg + rr * FRW + 4 * c4;
                        if (gx >= 0 && gx + 3 < W) {
                            cp_async16(dp, predb + rowo + gx);
                            cp_async16(dg, gtb + rowo + gx);
                        } else {   // whole float4 outside (W % 4 == 0): replicate the border pixel
                            const int cx = gx < 0 ? 0 : W - 1;
                            const float ps = __ldg(predb + rowo + cx), gs = __ldg(gtb + rowo + cx);
                            *reinterpret_cast<float4*>(dp) = make_float4(ps, ps, ps, ps);
                            *reinterpret_cast<float4*>(dg) = make_float4(gs, gs, gs, gs);
                        }
                    }
                }
            }
            cp_async_wait_all();
        }
        fast_prelude<0>(a, sm, tid, y0, x0, a.use_tma != 0, acc);
    }
    __syncthreads();
    fast_p4<F, HAS_MASK>(a, sm, warp, lane, b, y0, x0, s_c, acc);

    if (publish_partials(a, acc, tile, s_f, &s_last)) {
        finalize_results(a, s_d);
        if (a.metrics) write_metric_results(a.stats, a.metrics, *a.results, tid);
    }
}


// ================================================================================================
// Fast streaming kernel for the pointwise terms alone (BASELINE config 2: reprojection fwd+bwd).
// One warp per 128-pixel row segment, 2 segments in flight per warp; 12 B/px of HBM traffic.
// Requires W % 4 == 0 and 16-byte aligned tensors (same dispatch condition as the tile fast path).
// ================================================================================================
// COUNT (reprojection alone, cooperative launch): the kernel counts the valid pixels itself in a first sweep over gt
// (the only statistic this term needs), meets at a grid-wide barrier, and then runs the gradient sweep -- no separate
// phase-A launch, no reduce-kernel epilogue, no launch gap (~8 us of a 43 us step at config 2).
#ifndef CADL_PT_MINB
#define CADL_PT_MINB 3
#endif
#ifndef CADL_PT_NB
#define CADL_PT_NB 3
#endif
template <int F, bool HAS_MASK, bool COUNT = false>
__global__ void __launch_bounds__(kThreadsB, CADL_PT_MINB) phase_b_point_fast_kernel(const PhaseBArgs a) {
    __shared__ float s_f[kThreadsB / 32][BF_COUNT];
    __shared__ double s_d[8];
    __shared__ float s_c[4];
    __shared__ int s_last;
    __shared__ unsigned s_cnt[kThreadsB / 32];
    constexpr bool SI = (F & FB_SI) != 0, RP = (F & FB_RP) != 0;
    static_assert(!COUNT || F == FB_RP, "the in-kernel count exists for the reprojection term alone");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float up = a.upstream;
    if constexpr (COUNT) {
        const int segs = (a.W + 127) >> 7, items = a.H * segs, b = blockIdx.y;
        const int wstride = gridDim.x * (kThreadsB / 32);
        unsigned cnt = 0;
        const int dq = wstride / segs, dr = wstride - dq * segs;
        int it = blockIdx.x * (kThreadsB / 32) + warp;
        int y = it / segs, sg = it - y * segs;
        constexpr int NC = 6;                                   // loads in flight per lane
        for (; it < items; it += NC * wstride) {
            float4 g[NC];
            uchar4 u[NC];
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                const int x = (sg << 7) + 4 * lane;
                g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                u[j] = make_uchar4(0, 0, 0, 0);
                if (it + j * wstride < items && x < a.W) {
                    const int o = (b * a.H + y) * a.W + x;
                    if constexpr (HAS_MASK) u[j] = __ldg(reinterpret_cast<const uchar4*>(a.mask + o));
                    else g[j] = __ldg(reinterpret_cast<const float4*>(a.gt + o));
                }
                y += dq; sg += dr;
                if (sg >= segs) { sg -= segs; ++y; }
            }
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                if constexpr (HAS_MASK) cnt += (u[j].x != 0) + (u[j].y != 0) + (u[j].z != 0) + (u[j].w != 0);
                else cnt += (g[j].x > a.eps_rp) + (g[j].y > a.eps_rp) + (g[j].z > a.eps_rp) + (g[j].w > a.eps_rp);   // depth_loss.h:318-320
            }
        }
        cnt = warp_sum(cnt);
        if (lane == 0) s_cnt[warp] = cnt;
        __syncthreads();
        if (tid == 0) {
            unsigned t = 0;
            for (int w = 0; w < kThreadsB / 32; ++w) t += s_cnt[w];
            atomicAdd(&a.hdr->icount[AI_RP_N], (unsigned long long)t);      // integer: order-free, exact
        }
        cooperative_groups::this_grid().sync();
        if (tid == 0) {
            const double nr = (double)*reinterpret_cast<volatile unsigned long long*>(&a.hdr->icount[AI_RP_N]);
            if (blockIdx.x == 0 && blockIdx.y == 0) const_cast<double*>(a.stats)[ST_RP_N] = nr;   // what finalize_results reads
            s_c[0] = 0.f; s_c[1] = 0.f;
            s_c[2] = nr > 0.0 ? (float)(1.0 / nr) * a.w_rp * up : 0.f;
        }
    } else {
    pdl_wait();        // launched behind phase A with programmatic stream serialization: its statistics
    if (tid == 0) {
        const double n = a.stats[ST_SI_N], S = a.stats[ST_SI_S], nr = a.stats[ST_RP_N];
        s_c[0] = n > 0.0 ? (float)(2.0 / n) * a.w_si * up : 0.f;
        s_c[1] = n > 0.0 ? (float)(-2.0 * (double)a.lambda * S / (n * n)) * a.w_si * up : 0.f;
        s_c[2] = nr > 0.0 ? (float)(1.0 / nr) * a.w_rp * up : 0.f;
    }
    }
    __syncthreads();
    const float c1 = s_c[0], c2 = s_c[1], rpn = s_c[2];
    const float eps_s = a.eps_si, eps_r = a.eps_rp;
    float acc[BF_COUNT];
#pragma unroll
    for (int q = 0; q < BF_COUNT; ++q) acc[q] = 0.f;

    // grid = (blocks per image, B): per-image constants once per CTA; a warp walks whole rows of its image
    const int H = a.H, W = a.W;
    const int b = blockIdx.y;
    const int segs = (W + 127) >> 7;
    float fxe = 1.f, fye = 1.f, rfx = 1.f, rfy = 1.f, cxv = 0.f, cyv = 0.f;
    bool mk_ok = true;
    if constexpr (RP) {
        float fx, fy;
        load_K(a, b, fx, fy, cxv, cyv);
        fxe = fx + eps_r; fye = fy + eps_r;               // depth_loss.h:299-300
        rfx = __frcp_rn(fxe); rfy = __frcp_rn(fye);
        mk_ok = markstein_safe(fxe) && markstein_safe(fye);
    }
    // Work items = 128-pixel row segments of this image, dealt round-robin to the warps of the image's CTAs (whole
    // rows per warp left some warps with 4 rows and others with 3 at config 2).  NB items per batch: all loads of
    // a batch are issued before any arithmetic.
    const float2 rfx2 = make_float2(rfx, rfx), rfy2 = make_float2(rfy, rfy), fxe2 = make_float2(fxe, fxe), fye2 = make_float2(fye, fye);
    float2 acc2 = make_float2(0.f, 0.f);     // sum of e, two lanes of the packed pipe
    const int wstride = gridDim.x * (kThreadsB / 32);
    const int items = H * segs;
    constexpr int NB = CADL_PT_NB;
    // (row, segment) of an item advance by a fixed (dq, dr) per stride: no integer division in the loop
    const int dq = wstride / segs, dr = wstride - dq * segs;
    int i0 = blockIdx.x * (kThreadsB / 32) + warp;
    int y0 = i0 / segs, s0 = i0 - y0 * segs;
    for (; i0 < items; i0 += NB * wstride) {
      {
        float4 p4[NB], g4[NB];
        uchar4 u4[NB];
        int ys[NB], xs[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          ys[j] = y0;
          xs[j] = (s0 << 7) + 4 * lane;
          const bool on = (i0 + j * wstride < items) && xs[j] < W;
          if (!on) xs[j] = -1;
          p4[j] = make_float4(1.f, 1.f, 1.f, 1.f); g4[j] = make_float4(0.f, 0.f, 0.f, 0.f); u4[j] = make_uchar4(0, 0, 0, 0);
          if (on) {
            const int o = (b * H + y0) * W + xs[j];
            p4[j] = __ldg(reinterpret_cast<const float4*>(a.pred + o));
            g4[j] = __ldg(reinterpret_cast<const float4*>(a.gt + o));
            if constexpr (HAS_MASK) u4[j] = __ldg(reinterpret_cast<const uchar4*>(a.mask + o));
          }
          y0 += dq; s0 += dr;
          if (s0 >= segs) { s0 -= segs; ++y0; }
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int y = ys[j], x = xs[j];
          if (x < 0) continue;
          const int row = b * H + y;
          const float ayv = RP ? (float)y - cyv : 0.f;
          const float yh = ayv * rfy;                         // d pY / d p, tolerance path
          const int off = row * W + x;
          const bool um[4] = {u4[j].x != 0, u4[j].y != 0, u4[j].z != 0, u4[j].w != 0};
          const float p[4] = {p4[j].x, p4[j].y, p4[j].z, p4[j].w}, g[4] = {g4[j].x, g4[j].y, g4[j].z, g4[j].w};
          const float xf = (float)x;
          float lp[4] = {0.f, 0.f, 0.f, 0.f}, lg[4] = {0.f, 0.f, 0.f, 0.f};
          if constexpr (SI) {
#pragma unroll
            for (int k = 0; k < 4; k += 2) {
              const float2 a2 = log_exact2(make_float2(clamp_nan(p[k], eps_s, 1000.0f), clamp_nan(p[k + 1], eps_s, 1000.0f)));
              const float2 b2 = log_exact2(make_float2(clamp_nan(g[k], eps_s, 1000.0f), clamp_nan(g[k + 1], eps_s, 1000.0f)));
              lp[k] = a2.x; lp[k + 1] = a2.y; lg[k] = b2.x; lg[k + 1] = b2.y;
            }
          }
          float out[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float gsum = 0.f;
            if constexpr (SI) {
              const bool m = HAS_MASK ? um[k] : (g[k] > eps_s);
              if (m && in_range_pos(p[k], eps_s, 1000.0f)) gsum = fmaf(c1, lp[k] - lg[k], c2) * rcp_approx(p[k]);
            }
            out[k] = gsum;
          }
          if constexpr (RP) {
            // same operations, same order as depth_loss.h:299-315 -- X = ((u - cx) * d) / (fx + eps) -- two pixels
            // per instruction on the packed fp32x2 pipes (the pairs line up with the float4 loads)
            const float2 ay2 = make_float2(ayv, ayv), yh2 = make_float2(yh, yh);
            const float2 rpn2 = make_float2(rpn, rpn), eps2 = make_float2(eps_r, eps_r);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float2 pp = make_float2(p[2 * h], p[2 * h + 1]), gg = make_float2(g[2 * h], g[2 * h + 1]);
              const float2 ax = make_float2((xf + (float)(2 * h)) - cxv, (xf + (float)(2 * h + 1)) - cxv);   // (float)(x+k) is exact
              const float2 xh = __fmul2_rn(ax, rfx2);
              const float2 tpx = __fmul2_rn(ax, pp), tgx = __fmul2_rn(ax, gg), tpy = __fmul2_rn(ay2, pp), tgy = __fmul2_rn(ay2, gg);
              float2 pX, gX, pY, gY;
              if (mk_ok) {
                // Markstein: q0 = t * rb, rem = t - q0 * b (exact), q = q0 + rem * rb = RN(t / b)
                float2 q0 = __fmul2_rn(tpx, rfx2); pX = __ffma2_rn(__ffma2_rn(make_float2(-q0.x, -q0.y), fxe2, tpx), rfx2, q0);
                q0 = __fmul2_rn(tgx, rfx2);        gX = __ffma2_rn(__ffma2_rn(make_float2(-q0.x, -q0.y), fxe2, tgx), rfx2, q0);
                q0 = __fmul2_rn(tpy, rfy2);        pY = __ffma2_rn(__ffma2_rn(make_float2(-q0.x, -q0.y), fye2, tpy), rfy2, q0);
                q0 = __fmul2_rn(tgy, rfy2);        gY = __ffma2_rn(__ffma2_rn(make_float2(-q0.x, -q0.y), fye2, tgy), rfy2, q0);
              } else {
                pX = make_float2(__fdiv_rn(tpx.x, fxe), __fdiv_rn(tpx.y, fxe)); gX = make_float2(__fdiv_rn(tgx.x, fxe), __fdiv_rn(tgx.y, fxe));
                pY = make_float2(__fdiv_rn(tpy.x, fye), __fdiv_rn(tpy.y, fye)); gY = make_float2(__fdiv_rn(tgy.x, fye), __fdiv_rn(tgy.y, fye));
              }
              const float2 dX = __fadd2_rn(pX, make_float2(-gX.x, -gX.y)), dY = __fadd2_rn(pY, make_float2(-gY.x, -gY.y));
              const float2 dZ = __fadd2_rn(pp, make_float2(-gg.x, -gg.y));
              const float2 ss = __fadd2_rn(__ffma2_rn(dZ, dZ, __ffma2_rn(dY, dY, __fmul2_rn(dX, dX))), eps2);   // :313-315
              float2 re = make_float2(rsqrt_approx(ss.x), rsqrt_approx(ss.y));
              const bool m0 = HAS_MASK ? um[2 * h] : (gg.x > eps_r), m1 = HAS_MASK ? um[2 * h + 1] : (gg.y > eps_r);
              re.x = m0 ? re.x : 0.f; re.y = m1 ? re.y : 0.f;
              acc2 = __ffma2_rn(ss, re, acc2);                                                    // e = sqrt(ss)
              const float2 t = __fmul2_rn(__ffma2_rn(dX, xh, __ffma2_rn(dY, yh2, dZ)), re);
              const float2 o = __ffma2_rn(t, rpn2, make_float2(out[2 * h], out[2 * h + 1]));
              out[2 * h] = o.x; out[2 * h + 1] = o.y;
            }
          }
          if (a.grad) *reinterpret_cast<float4*>(a.grad + off) = make_float4(out[0], out[1], out[2], out[3]);
        }
      }
    }
    acc[BF_RP_E] = acc2.x + acc2.y;
    if (publish_partials(a, acc, blockIdx.y * gridDim.x + blockIdx.x, s_f, &s_last)) {
        if (COUNT && tid == 0) a.hdr->icount[AI_RP_N] = 0ull;      // everybody has read it: leave the counter clean
        finalize_results(a, s_d);
        if (a.metrics) write_metric_results(a.stats, a.metrics, *a.results, tid);
    }
}

}  // namespace cadl
