// cadl phase B: the gradient pass.  Two kernels:
//
//  * phase_b_tile_kernel<F>  -- terms with stencils (gradient matching over 4 scales, edge-aware
//    smoothness) fused with the pointwise terms (SI, reprojection).  One CTA owns a TH x TW tile;
//    pred/gt are staged in shared memory with an 8-pixel halo (= one cell of the coarsest scale), the
//    avg-pool pyramid is built from that staging in the reference's summation order, rgb is streamed
//    row by row through registers (read once from HBM, never staged), and the gradient tile is written
//    once with 128-bit stores.
//  * phase_b_point_kernel<F> -- SI and/or reprojection alone: pure streaming, no staging.
//
// Both write per-CTA partial loss sums as fp64 rows; the last CTA (ticket) reduces them in a fixed
// order and writes cadl_results.
#pragma once
#include "cadl_common.cuh"
#include "cadl_phase_a.cuh"
#include "cadl_args.cuh"

namespace cadl {

// ---- shared-memory geometry of the tile kernel -------------------------------------------------
constexpr int RH = TH + 2 * HALO;      // 48 staged rows
constexpr int RW = TW + 2 * HALO;      // 144 staged cols (576 B pitch, 16 B multiple)
constexpr int LH = TH + 2;             // log rows  (-1 .. TH)
constexpr int LW = TW + 8;             // log pitch: interior col 0 at index 4 (float4 aligned)
__host__ __device__ constexpr int pool_h(int s) { return TH / (1 << s) + 2; }
__host__ __device__ constexpr int pool_w(int s) { return TW / (1 << s) + 2; }
__host__ __device__ constexpr int cc_h(int s) { return TH / (1 << s); }
__host__ __device__ constexpr int cc_w(int s) { return TW / (1 << s); }
constexpr int kPoolCells = pool_h(1) * pool_w(1) + pool_h(2) * pool_w(2) + pool_h(3) * pool_w(3);
constexpr int kCCCells = cc_h(1) * cc_w(1) + cc_h(2) * cc_w(2) + cc_h(3) * cc_w(3);
__host__ __device__ constexpr int pool_off(int s) {
    return s == 1 ? 0 : (s == 2 ? pool_h(1) * pool_w(1) : pool_h(1) * pool_w(1) + pool_h(2) * pool_w(2));
}
__host__ __device__ constexpr int cc_off(int s) {
    return s == 1 ? 0 : (s == 2 ? cc_h(1) * cc_w(1) : cc_h(1) * cc_w(1) + cc_h(2) * cc_w(2));
}
constexpr size_t kTileSmemFloats = 2 * RH * RW + 2 * LH * LW + 2 * kPoolCells + kCCCells;
constexpr size_t kTileSmemBytes = kTileSmemFloats * sizeof(float);

struct TileSmem {
    float* sp;   // raw pred  [RH][RW]
    float* sg;   // raw gt    [RH][RW]
    float* lp;   // log pred  [LH][LW]
    float* lg;   // log gt    [LH][LW]
    float* pl;   // pooled log pred, scales 1..3 (with 1-cell halo)
    float* pg;   // pooled log gt
    float* cc;   // per-cell gradient coefficient, scales 1..3 (interior cells)
};

__device__ __forceinline__ TileSmem carve(float* base) {
    TileSmem s;
    s.sp = base;
    s.sg = s.sp + RH * RW;
    s.lp = s.sg + RH * RW;
    s.lg = s.lp + LH * LW;
    s.pl = s.lg + LH * LW;
    s.pg = s.pl + kPoolCells;
    s.cc = s.pg + kPoolCells;
    return s;
}

// Final reduction + results, executed by ONE CTA (all of its threads, >= 64) after every partial row is visible.
// Written as a short dependency chain -- one round of independent global loads, block reductions, six result
// terms on six threads, one writer -- because it runs with the rest of the GPU idle or waiting (the first version,
// a lane-strided dependent load loop per quantity and one thread doing every division, took 20-30 us).
// Deterministic: fixed thread-strided order, fixed shuffle tree, fixed warp order.
__device__ void finalize_results(const PhaseBArgs& a, double* s_d) {
    (void)s_d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
    const double* part = a.b_part;
    constexpr int NQ = BF_COUNT + 1;                 // + the a_b-weighted smoothness total
    __shared__ double s_w[32][NQ];
    __shared__ double s_tot[NQ];
    __shared__ double s_term[4 + 2];
    const bool smooth = (a.terms & CADL_TERM_SMOOTH) != 0;
    const int tpi = a.tiles_x * a.tiles_y;

    // ---- 1. global loads: partial rows (three per thread and trip), per-image smoothness shares ----
    double accq[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) accq[q] = 0.0;
    for (int i = tid; i < a.b_rows; i += 3 * blockDim.x) {
        const int i2 = i + blockDim.x, i3 = i + 2 * blockDim.x;
        double r0[BF_COUNT], r1[BF_COUNT], r2[BF_COUNT];
#pragma unroll
        for (int q = 0; q < BF_COUNT; ++q) {      // __ldcg: written by other SMs, never cached in this L1
            r0[q] = __ldcg(part + (size_t)i * BF_COUNT + q);
            r1[q] = i2 < a.b_rows ? __ldcg(part + (size_t)i2 * BF_COUNT + q) : 0.0;
            r2[q] = i3 < a.b_rows ? __ldcg(part + (size_t)i3 * BF_COUNT + q) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < BF_COUNT; ++q) accq[q] = ((accq[q] + r0[q]) + r1[q]) + r2[q];
    }
    if (smooth) {
        // per-image sums (the partial rows of an image are contiguous), then a_b-weighted
        if (tpi == 1) {   // rows already folded per image (streaming kernel): one thread per image
            for (int img = tid; img < a.B; img += blockDim.x) {
                double Lb;
                float off;
                smooth_image_share(a, img, __ldcg(part + (size_t)img * BF_COUNT + BF_SMX),
                                   __ldcg(part + (size_t)img * BF_COUNT + BF_SMY), Lb, off);
                a.img_sm[2 * img] = Lb;
                a.img_off[img] = off;
                accq[BF_COUNT] += Lb;
            }
        } else {
            for (int img = warp; img < a.B; img += nwarp) {
                double sx = 0.0, sy = 0.0;
#pragma unroll 4
                for (int i = lane; i < tpi; i += 32) {
                    sx += __ldcg(part + ((size_t)img * tpi + i) * BF_COUNT + BF_SMX);
                    sy += __ldcg(part + ((size_t)img * tpi + i) * BF_COUNT + BF_SMY);
                }
                sx = warp_sum(sx);
                sy = warp_sum(sy);
                if (lane == 0) {
                    double Lb;
                    float off;
                    smooth_image_share(a, img, sx, sy, Lb, off);
                    a.img_sm[2 * img] = Lb;
                    a.img_off[img] = off;
                    accq[BF_COUNT] += Lb;
                }
            }
        }
    }
    // ---- 2. block reduction ----
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const double v = warp_sum(accq[q]);
        if (lane == 0) s_w[warp][q] = v;
    }
    __syncthreads();
    if (tid < NQ) {
        double t = 0.0;
        for (int w = 0; w < nwarp; ++w) t += s_w[w][tid];
        s_tot[tid] = t;
    }
    __syncthreads();
    // ---- 3. the result terms, one thread each ----
    const double* st = a.stats;
    if (tid < 4) {            // gradient matching, scale tid                       depth_loss.h:162-163
        double term = 0.0;
        if ((a.terms & CADL_TERM_GRAD) && tid < a.num_scales) {
            const int Hs = a.H >> tid, Ws = a.W >> tid;
            const double nx = (double)a.global_B * Hs * (Ws - 1);
            const double ny = (double)a.global_B * (Hs - 1) * Ws;
            term = s_tot[BF_GX0 + 2 * tid] / nx + s_tot[BF_GY0 + 2 * tid] / ny;
        }
        s_term[tid] = term;
    } else if (tid == 32) {   // scale-invariant                                      depth_loss.h:58-63
        double si = 0.0;
        if (a.terms & CADL_TERM_SI) {
            const double n = st[ST_SI_N];
            if (n > 0.0) si = st[ST_SI_Q] / n - (double)a.lambda * st[ST_SI_S] * st[ST_SI_S] / (n * n);
        }
        s_term[4] = si;
    } else if (tid == 33) {   // reprojection                                         depth_loss.h:323-330
        double rp = 0.0;
        if (a.terms & CADL_TERM_REPROJ) {
            const double n = st[ST_RP_N];
            if (n > 0.0) rp = s_tot[BF_RP_E] / n;
        }
        s_term[5] = rp;
    }
    __syncthreads();
    if (tid == 0) {
        cadl_results& r = *a.results;
        double gm = 0.0;
        for (int s = 0; s < a.num_scales; ++s) gm += s_term[s];
        gm = (a.terms & CADL_TERM_GRAD) ? gm / (double)a.num_scales : 0.0;
        const double si = s_term[4], rp = s_term[5], sm = smooth ? s_tot[BF_COUNT] : 0.0;
        r.n_si = (a.terms & CADL_TERM_SI) ? (int64_t)st[ST_SI_N] : 0;
        r.n_reproj = (a.terms & CADL_TERM_REPROJ) ? (int64_t)st[ST_RP_N] : 0;
        r.d_si = si; r.d_grad = gm; r.d_smooth = sm; r.d_reproj = rp;
        r.loss_si = (float)si; r.loss_grad = (float)gm; r.loss_smooth = (float)sm; r.loss_reproj = (float)rp;
        // depth_loss.h:427-430, in float like the reference's tensor arithmetic
        float tot = 0.f;
        if (a.terms & CADL_TERM_SI) tot = a.w_si * r.loss_si;
        if (a.terms & CADL_TERM_GRAD) tot = tot + a.w_grad * r.loss_grad;
        if (a.terms & CADL_TERM_SMOOTH) tot = tot + a.w_smooth * r.loss_smooth;
        if (a.terms & CADL_TERM_REPROJ) tot = tot + a.w_rp * r.loss_reproj;
        r.loss_total = tot;
        r.d_total = (double)a.w_si * si + (double)a.w_grad * gm + (double)a.w_smooth * sm + (double)a.w_rp * rp;
        a.hdr->ticket_b = 0u;
    }
}

// Block partial row + ticket; returns true in the last block.
__device__ __forceinline__ bool publish_partials(const PhaseBArgs& a, float (&acc)[BF_COUNT], int row,
                                                 float (*s_f)[BF_COUNT], int* s_last) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int q = 0; q < BF_COUNT; ++q) {
        float v = warp_sum(acc[q]);
        if (lane == 0) s_f[warp][q] = v;
    }
    __syncthreads();
    if (tid < BF_COUNT) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += (double)s_f[w][tid];
        a.b_part[(size_t)row * BF_COUNT + tid] = t;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned t = atomicAdd(&a.hdr->ticket_b, 1u);
        *s_last = (t == (unsigned)a.b_rows - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (*s_last) __threadfence();
    return *s_last != 0;
}

// Per-pixel camera geometry of the reprojection term (depth_loss.h:283-300).
struct RpGeom {
    float ax, ay;     // (u - cx), (v - cy)
    float fxe, fye;   // fx + eps, fy + eps
    float xh, yh;     // ax / fxe, ay / fye  (d pX / d p)
};

// Pointwise terms for one pixel; returns the weighted gradient contribution (upstream applied by caller).
template <int F>
__device__ __forceinline__ float pointwise_px(const PhaseBArgs& a, const Derived& dv, float p, float g,
                                              bool has_mask, bool um, float d_si, const RpGeom& q,
                                              float& rp_e_acc) {
    float gr = 0.f;
    if constexpr (F & FB_SI) {
        bool m = has_mask ? um : (g > a.eps_si);
        bool cm = (p >= a.eps_si) && (p <= 1000.0f);           // clamp backward (closed interval)
        if (m && cm && dv.si_on) gr += a.w_si * ((dv.si_c1 * d_si + dv.si_c2) / p);
    }
    if constexpr (F & FB_RP) {
        bool m = has_mask ? um : (g > a.eps_rp);
        if (m && dv.rp_on) {
            // The reference subtracts two back-projected points, X_p - X_g with X = (u-cx)*d/(fx+eps)
            // (depth_loss.h:299-311).  Where pred ~ gt that difference is dominated by the rounding of
            // the two quotients, and e ~ sqrt(eps) amplifies it into the gradient (1e-4 of max|g|), so
            // the same operations are done in the same order; the factored form x_hat*(p-g) would be
            // more accurate but would not reproduce the reference's values.
            const float dX = __fdiv_rn(__fmul_rn(q.ax, p), q.fxe) - __fdiv_rn(__fmul_rn(q.ax, g), q.fxe);
            const float dY = __fdiv_rn(__fmul_rn(q.ay, p), q.fye) - __fdiv_rn(__fmul_rn(q.ay, g), q.fye);
            const float dZ = p - g;
            const float ss = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(dX, dX), __fmul_rn(dY, dY)), __fmul_rn(dZ, dZ)), a.eps_rp);
            const float e = sqrtf(ss);                                                // :313-315
            rp_e_acc += e;
            // autograd: dL/dp = (dX * xh + dY * yh + dZ) / (e * n)
            gr += a.w_rp * (((dX * q.xh + dY * q.yh + dZ) / e) * dv.rp_inv_n);
        }
    }
    return gr;
}


// ================================================================================================
// pointwise kernel: SI and/or reprojection (no stencil)
// ================================================================================================
template <int F>
__global__ void __launch_bounds__(kThreadsB) phase_b_point_kernel(const PhaseBArgs a) {
    __shared__ float s_f[kThreadsB / 32][BF_COUNT];
    __shared__ double s_d[8];
    __shared__ int s_last;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const Derived dv = derive(a);
    const bool has_mask = a.mask != nullptr;
    float acc[BF_COUNT];
#pragma unroll
    for (int q = 0; q < BF_COUNT; ++q) acc[q] = 0.f;

    const int segs = (a.W + TW - 1) / TW;
    const long long items = (long long)a.B * a.H * segs;
    const long long wstride = (long long)gridDim.x * (kThreadsB / 32);
    for (long long it = (long long)blockIdx.x * (kThreadsB / 32) + warp; it < items; it += wstride) {
        const int seg = (int)(it % segs);
        const long long row = it / segs;
        const int y = (int)(row % a.H);
        const int b = (int)(row / a.H);
        const int x = seg * TW + 4 * lane;
        if (x >= a.W) continue;
        const size_t off = ((size_t)b * a.H + y) * a.W + x;
        float p[4], g[4];
        bool um[4] = {true, true, true, true};
        const bool full = a.vec_ok && (x + 3 < a.W);
        if (full) {
            float4 p4 = __ldg(reinterpret_cast<const float4*>(a.pred + off));
            float4 g4 = __ldg(reinterpret_cast<const float4*>(a.gt + off));
            p[0] = p4.x; p[1] = p4.y; p[2] = p4.z; p[3] = p4.w;
            g[0] = g4.x; g[1] = g4.y; g[2] = g4.z; g[3] = g4.w;
            if (has_mask) {
                uchar4 u = __ldg(reinterpret_cast<const uchar4*>(a.mask + off));
                um[0] = u.x != 0; um[1] = u.y != 0; um[2] = u.z != 0; um[3] = u.w != 0;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                bool in = x + k < a.W;
                p[k] = in ? __ldg(a.pred + off + k) : 1.f;
                g[k] = in ? __ldg(a.gt + off + k) : 0.f;
                um[k] = in ? (has_mask ? (__ldg(a.mask + off + k) != 0) : true) : false;
            }
        }
        RpGeom q{};
        float cx = 0.f;
        if constexpr (F & FB_RP) {
            float fx, fy, cy;
            load_K(a, b, fx, fy, cx, cy);
            q.fxe = fx + a.eps_rp;                               // depth_loss.h:299-300
            q.fye = fy + a.eps_rp;
            q.ay = (float)y - cy;
            q.yh = q.ay / q.fye;
        }
        float out[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float d_si = 0.f;
            if constexpr (F & FB_SI) {
                d_si = logf(clampf(p[k], a.eps_si, 1000.0f)) - logf(clampf(g[k], a.eps_si, 1000.0f));
            }
            if constexpr (F & FB_RP) {
                q.ax = (float)(x + k) - cx;
                q.xh = q.ax / q.fxe;
            }
            bool in = x + k < a.W;
            bool mk = has_mask ? um[k] : in;
            float gr = 0.f;
            if (in) gr = pointwise_px<F>(a, dv, p[k], g[k], has_mask, mk, d_si, q, acc[BF_RP_E]);
            out[k] = gr * a.upstream;
        }
        if (a.grad) {
            if (full) {
                *reinterpret_cast<float4*>(a.grad + off) = make_float4(out[0], out[1], out[2], out[3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (x + k < a.W) a.grad[off + k] = out[k];
            }
        }
    }
    if (publish_partials(a, acc, blockIdx.x, s_f, &s_last)) {
        finalize_results(a, s_d);
        if (a.metrics) write_metric_results(a.stats, a.metrics, *a.results, tid);
    }
}

// ================================================================================================
// tile kernel: stencil terms (+ pointwise terms fused)
// ================================================================================================
template <int F>
__global__ void __launch_bounds__(kThreadsB, 2) phase_b_tile_kernel(const PhaseBArgs a) {
    extern __shared__ __align__(128) float smem_raw[];
    __shared__ float s_f[kThreadsB / 32][BF_COUNT];
    __shared__ double s_d[8];
    __shared__ int s_last;
    const TileSmem sm = carve(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const int tile = blockIdx.x;
    const int tx = tile % a.tiles_x;
    const int ty = (tile / a.tiles_x) % a.tiles_y;
    const int b = tile / (a.tiles_x * a.tiles_y);
    const int x0 = tx * TW, y0 = ty * TH;
    const int H = a.H, W = a.W;
    const size_t img = (size_t)b * H * W;
    const bool has_mask = a.mask != nullptr;
    const Derived dv = derive(a);
    constexpr bool GRAD = (F & FB_GRAD) != 0;
    constexpr bool SMOOTH = (F & FB_SMOOTH) != 0;
    // halo actually needed: 8 for the pooled scales, 1 otherwise
    const int halo = (GRAD && a.num_scales > 1) ? HALO : 1;

    float acc[BF_COUNT];
#pragma unroll
    for (int q = 0; q < BF_COUNT; ++q) acc[q] = 0.f;

    // ---------------- phase 1: stage raw pred / gt (zero outside the image) ----------------
    {
        const int r_lo = HALO - halo, r_hi = HALO + TH + halo;       // staged row range
        const int c_lo = (HALO - halo) & ~3, c_hi = ((HALO + TW + halo) + 3) & ~3;
        const int nrow = r_hi - r_lo, ncol4 = (c_hi - c_lo) >> 2;
        for (int i = tid; i < nrow * ncol4; i += kThreadsB) {
            const int rr = r_lo + i / ncol4, cc = c_lo + 4 * (i % ncol4);
            const int gy = y0 - HALO + rr, gx = x0 - HALO + cc;
            float4 pv = make_float4(0.f, 0.f, 0.f, 0.f), gv = pv;
            if (gy >= 0 && gy < H) {
                const size_t o = img + (size_t)gy * W + gx;
                if (a.vec_ok && gx >= 0 && gx + 3 < W) {
                    pv = __ldg(reinterpret_cast<const float4*>(a.pred + o));
                    if (GRAD || (F & (FB_SI | FB_RP))) gv = __ldg(reinterpret_cast<const float4*>(a.gt + o));
                } else {
                    float* pp = &pv.x; float* gg = &gv.x;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (gx + k >= 0 && gx + k < W) {
                            pp[k] = __ldg(a.pred + o + k);
                            if (GRAD || (F & (FB_SI | FB_RP))) gg[k] = __ldg(a.gt + o + k);
                        }
                    }
                }
            }
            *reinterpret_cast<float4*>(sm.sp + rr * RW + cc) = pv;
            *reinterpret_cast<float4*>(sm.sg + rr * RW + cc) = gv;
        }
    }
    __syncthreads();

    // ---------------- phase 2: avg-pool pyramid (reference summation order) + logs ----------------
    if constexpr (GRAD) {
        if (a.num_scales > 1) {
            // one thread per (8x8 block, tensor): sums for scales 1..3 from registers, row-major
            // sequential inside each window == ATen avg_pool2d's loop order (CPU and CUDA).
            constexpr int BR = RH / 8, BC = RW / 8;   // 6 x 18 blocks
            for (int item = tid; item < 2 * BR * BC; item += kThreadsB) {
                const int t = item / (BR * BC);
                const int blk = item % (BR * BC);
                const int by = blk / BC, bx = blk % BC;
                const float* src = (t == 0 ? sm.sp : sm.sg) + (by * 8) * RW + bx * 8;
                float* dst = (t == 0 ? sm.pl : sm.pg);
                float v[8][8];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    float4 lo = *reinterpret_cast<const float4*>(src + r * RW);
                    float4 hi = *reinterpret_cast<const float4*>(src + r * RW + 4);
                    v[r][0] = lo.x; v[r][1] = lo.y; v[r][2] = lo.z; v[r][3] = lo.w;
                    v[r][4] = hi.x; v[r][5] = hi.y; v[r][6] = hi.z; v[r][7] = hi.w;
                }
#pragma unroll
                for (int s = 1; s <= 3; ++s) {
                    if (s >= a.num_scales) break;
                    const int f = 1 << s, nc = 8 >> s;       // cells per block side
                    const int Hs = H >> s, Ws = W >> s;
                    const float inv_area = 1.0f / (float)(f * f);
#pragma unroll
                    for (int ci = 0; ci < nc; ++ci) {
#pragma unroll
                        for (int cj = 0; cj < nc; ++cj) {
                            float sum = 0.f;
#pragma unroll
                            for (int r = 0; r < f; ++r)
#pragma unroll
                                for (int c = 0; c < f; ++c) sum += v[ci * f + r][cj * f + c];
                            // local cell coords relative to the tile's first cell
                            const int cy = by * nc + ci - (HALO >> s);
                            const int cx = bx * nc + cj - (HALO >> s);
                            if (cy < -1 || cy > (TH >> s) || cx < -1 || cx > (TW >> s)) continue;
                            const int gy = (y0 >> s) + cy, gx = (x0 >> s) + cx;
                            const bool valid = gy >= 0 && gy < Hs && gx >= 0 && gx < Ws;
                            const float q = sum * inv_area;                     // sum / (f*f), exact
                            const float ql = valid ? logf(clampf(q, a.eps_grad, 1000.0f)) : 0.f;
                            dst[pool_off(s) + (cy + 1) * pool_w(s) + (cx + 1)] = ql;
                            if (t == 0 && cy >= 0 && cy < (TH >> s) && cx >= 0 && cx < (TW >> s)) {
                                const bool cm = (q >= a.eps_grad) && (q <= 1000.0f);
                                sm.cc[cc_off(s) + cy * cc_w(s) + cx] = (valid && cm) ? (1.0f / q) : 0.f;
                            }
                        }
                    }
                }
            }
        }
        // scale-0 logs over the (TH+2) x (TW+2) ring
        for (int i = tid; i < LH * (TW + 2); i += kThreadsB) {
            const int ry = i / (TW + 2) - 1, rx = i % (TW + 2) - 1;
            const float pv = sm.sp[(ry + HALO) * RW + rx + HALO];
            const float gv = sm.sg[(ry + HALO) * RW + rx + HALO];
            sm.lp[(ry + 1) * LW + rx + 4] = logf(clampf(pv, a.eps_grad, 1000.0f));   // depth_loss.h:115
            sm.lg[(ry + 1) * LW + rx + 4] = logf(clampf(gv, a.eps_grad, 1000.0f));   // depth_loss.h:116
        }
        __syncthreads();

        // ---------------- phase 3: per-cell gradient coefficients for scales 1..3 ----------------
        for (int s = 1; s < a.num_scales && s <= 3; ++s) {
            int Hs, Ws; float inv_nx, inv_ny;
            scale_dims(a, s, Hs, Ws, inv_nx, inv_ny);
            const int ch = TH >> s, cw = TW >> s, pw = (TW >> s) + 2;
            const float* PL = sm.pl + (s == 1 ? pool_off(1) : (s == 2 ? pool_off(2) : pool_off(3)));
            const float* PG = sm.pg + (s == 1 ? pool_off(1) : (s == 2 ? pool_off(2) : pool_off(3)));
            float* CC = sm.cc + (s == 1 ? cc_off(1) : (s == 2 ? cc_off(2) : cc_off(3)));
            const float spread = 1.0f / (float)((1 << s) * (1 << s)) / (float)a.num_scales;
            float ax = 0.f, ay = 0.f;
            for (int i = tid; i < ch * cw; i += kThreadsB) {
                const int cy = i / cw, cx = i % cw;
                const int gy = (y0 >> s) + cy, gx = (x0 >> s) + cx;
                float coef = 0.f;
                if (gy < Hs && gx < Ws) {
                    const int c = (cy + 1) * pw + (cx + 1);
                    const float lp = PL[c], lg = PG[c];
                    float sx_r = 0.f, sx_l = 0.f, sy_d = 0.f, sy_u = 0.f;
                    if (gx + 1 < Ws) {   // own right edge: depth_loss.h:140-148,162
                        const float e = (PL[c + 1] - lp) - (PG[c + 1] - lg);
                        ax += fabsf(e);
                        sx_r = sgnf(e);
                    }
                    if (gx >= 1) sx_l = sgnf((lp - PL[c - 1]) - (lg - PG[c - 1]));
                    if (gy + 1 < Hs) {   // own lower edge: depth_loss.h:151-159,163
                        const float e = (PL[c + pw] - lp) - (PG[c + pw] - lg);
                        ay += fabsf(e);
                        sy_d = sgnf(e);
                    }
                    if (gy >= 1) sy_u = sgnf((lp - PL[c - pw]) - (lg - PG[c - pw]));
                    coef = ((sx_l - sx_r) * inv_nx + (sy_u - sy_d) * inv_ny) * CC[i] * spread;
                }
                CC[i] = coef * a.w_grad;
            }
            acc[BF_GX0 + 2 * s] += ax;
            acc[BF_GY0 + 2 * s] += ay;
        }
        __syncthreads();
    }

    // ---------------- phase 4: full-resolution pass, one warp per row group ----------------
    {
        constexpr int RPW = TH / (kThreadsB / 32);   // rows per warp
        const int xl = 4 * lane;                     // local column of this lane's float4
        const int gx0 = x0 + xl;
        float fx = 1.f, fy = 1.f, cx = 0.f, cy = 0.f;
        float axk[4] = {0.f, 0.f, 0.f, 0.f}, xhk[4] = {0.f, 0.f, 0.f, 0.f};
        if constexpr (F & FB_RP) {
            load_K(a, b, fx, fy, cx, cy);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                axk[k] = (float)(gx0 + k) - cx;                 // depth_loss.h:299
                xhk[k] = axk[k] / (fx + a.eps_rp);
            }
        }
        float inv_nx0 = 0.f, inv_ny0 = 0.f;
        if constexpr (GRAD) { int Hs, Ws; scale_dims(a, 0, Hs, Ws, inv_nx0, inv_ny0); }
        float ab = 0.f, sm_nx = 0.f, sm_ny = 0.f;
        if constexpr (SMOOTH) {
            const double HWd = (double)H * W;
            ab = 1.0f / ((float)(a.img_psum[b] / HWd) + a.eps_smooth);      // depth_loss.h:192-193
            sm_nx = W > 1 ? (float)(1.0 / ((double)a.global_B * H * (W - 1))) : 0.f;
            sm_ny = H > 1 ? (float)(1.0 / ((double)a.global_B * (H - 1) * W)) : 0.f;
        }
        const float inv_ns = 1.0f / (float)a.num_scales;

        // rgb rows streamed through registers: prev (y-1), cur (y), next (y+1)
        const float* rgb_b = SMOOTH ? a.rgb + (size_t)b * 3 * H * W : nullptr;
        const size_t plane = (size_t)H * W;
        auto load_rgb_row = [&](int gy, float (&I)[3][6]) {
            // I[c][0] = x-1, I[c][1..4] = own 4 px, I[c][5] = x+4 ; zero outside the image.
            // gy is warp-uniform, so every lane reaches the shuffles.
            const bool row_ok = (gy >= 0) && (gy < H);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
#pragma unroll
                for (int k = 0; k < 6; ++k) I[c][k] = 0.f;
                const float* rowp = rgb_b + c * plane + (size_t)(row_ok ? gy : 0) * W;
                if (row_ok && gx0 < W) {
                    if (a.vec_ok && gx0 + 3 < W) {
                        float4 v = ldg_stream(reinterpret_cast<const float4*>(rowp + gx0));
                        I[c][1] = v.x; I[c][2] = v.y; I[c][3] = v.z; I[c][4] = v.w;
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (gx0 + k < W) I[c][1 + k] = __ldg(rowp + gx0 + k);
                    }
                }
                // neighbours across lanes; warp-edge lanes fetch the halo pixel themselves
                float left = __shfl_up_sync(0xffffffffu, I[c][4], 1);
                float right = __shfl_down_sync(0xffffffffu, I[c][1], 1);
                if (lane == 0) left = (row_ok && gx0 - 1 >= 0 && gx0 - 1 < W) ? __ldg(rowp + gx0 - 1) : 0.f;
                if (lane == 31) right = (row_ok && gx0 + 4 < W) ? __ldg(rowp + gx0 + 4) : 0.f;
                I[c][0] = left;
                I[c][5] = right;
            }
        };
        auto edge_w = [](float a0, float b0, float a1, float b1, float a2, float b2) {
            // exp(-mean_c |dI|)   depth_loss.h:211-227
            float m = (fabsf(a0 - b0) + fabsf(a1 - b1) + fabsf(a2 - b2)) * (1.0f / 3.0f);
            return expf(-m);
        };

        const int r_begin = warp * RPW;
        float Icur[3][6], Inxt[3][6];
        float wy_up[4] = {0.f, 0.f, 0.f, 0.f};   // w_y of the edge (y-1 -> y), per column
        if constexpr (SMOOTH) {
            float Iprev[3][6];
            load_rgb_row(y0 + r_begin - 1, Iprev);
            load_rgb_row(y0 + r_begin, Icur);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                wy_up[k] = edge_w(Iprev[0][1 + k], Icur[0][1 + k], Iprev[1][1 + k], Icur[1][1 + k],
                                  Iprev[2][1 + k], Icur[2][1 + k]);
        }

        for (int r = r_begin; r < r_begin + RPW; ++r) {
            const int gy = y0 + r;
            if constexpr (SMOOTH) load_rgb_row(gy + 1, Inxt);
            const bool row_in = gy < H;
            // raw pred / gt for own pixels and the 4 neighbours
            const float* prow = sm.sp + (r + HALO) * RW + HALO + xl;
            const float* grow = sm.sg + (r + HALO) * RW + HALO + xl;
            const float4 p4 = *reinterpret_cast<const float4*>(prow);
            const float4 g4 = *reinterpret_cast<const float4*>(grow);
            const float p[4] = {p4.x, p4.y, p4.z, p4.w};
            const float g[4] = {g4.x, g4.y, g4.z, g4.w};
            float out[4] = {0.f, 0.f, 0.f, 0.f};

            uchar4 mk = make_uchar4(1, 1, 1, 1);
            if (has_mask && row_in) {
                const size_t o = img + (size_t)gy * W + gx0;
                if (a.vec_ok && gx0 + 3 < W) mk = __ldg(reinterpret_cast<const uchar4*>(a.mask + o));
                else {
                    unsigned char* mm = &mk.x;
#pragma unroll
                    for (int k = 0; k < 4; ++k) mm[k] = (gx0 + k < W) ? __ldg(a.mask + o + k) : 0;
                }
            }
            const unsigned char um[4] = {mk.x, mk.y, mk.z, mk.w};

            float lp[4] = {0, 0, 0, 0}, lg[4] = {0, 0, 0, 0};
            if constexpr (GRAD) {
                const float* lprow = sm.lp + (r + 1) * LW + 4 + xl;
                const float* lgrow = sm.lg + (r + 1) * LW + 4 + xl;
                const float4 a4 = *reinterpret_cast<const float4*>(lprow);
                const float4 b4 = *reinterpret_cast<const float4*>(lgrow);
                lp[0] = a4.x; lp[1] = a4.y; lp[2] = a4.z; lp[3] = a4.w;
                lg[0] = b4.x; lg[1] = b4.y; lg[2] = b4.z; lg[3] = b4.w;
                const float lp_l = lprow[-1], lp_r = lprow[4], lg_l = lgrow[-1], lg_r = lgrow[4];
                const float4 lpu = *reinterpret_cast<const float4*>(lprow - LW);
                const float4 lpd = *reinterpret_cast<const float4*>(lprow + LW);
                const float4 lgu = *reinterpret_cast<const float4*>(lgrow - LW);
                const float4 lgd = *reinterpret_cast<const float4*>(lgrow + LW);
                const float lpx[6] = {lp_l, lp[0], lp[1], lp[2], lp[3], lp_r};
                const float lgx[6] = {lg_l, lg[0], lg[1], lg[2], lg[3], lg_r};
                const float lpuv[4] = {lpu.x, lpu.y, lpu.z, lpu.w}, lpdv[4] = {lpd.x, lpd.y, lpd.z, lpd.w};
                const float lguv[4] = {lgu.x, lgu.y, lgu.z, lgu.w}, lgdv[4] = {lgd.x, lgd.y, lgd.z, lgd.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int gx = gx0 + k;
                    if (!row_in || gx >= W) continue;
                    float sx_r = 0.f, sx_l = 0.f, sy_d = 0.f, sy_u = 0.f;
                    if (gx + 1 < W) {
                        const float e = (lpx[k + 2] - lpx[k + 1]) - (lgx[k + 2] - lgx[k + 1]);
                        acc[BF_GX0] += fabsf(e);
                        sx_r = sgnf(e);
                    }
                    if (gx >= 1) sx_l = sgnf((lpx[k + 1] - lpx[k]) - (lgx[k + 1] - lgx[k]));
                    if (gy + 1 < H) {
                        const float e = (lpdv[k] - lp[k]) - (lgdv[k] - lg[k]);
                        acc[BF_GY0] += fabsf(e);
                        sy_d = sgnf(e);
                    }
                    if (gy >= 1) sy_u = sgnf((lp[k] - lpuv[k]) - (lg[k] - lguv[k]));
                    const bool cm = (p[k] >= a.eps_grad) && (p[k] <= 1000.0f);
                    float c0 = cm ? (((sx_l - sx_r) * inv_nx0 + (sy_u - sy_d) * inv_ny0) / p[k]) * inv_ns : 0.f;
                    float gsum = c0 * a.w_grad;
                    // coarse scales: coefficient of the cell this pixel belongs to
                    const int lx = xl + k;
                    if (a.num_scales > 1) gsum += sm.cc[cc_off(1) + (r >> 1) * cc_w(1) + (lx >> 1)];
                    if (a.num_scales > 2) gsum += sm.cc[cc_off(2) + (r >> 2) * cc_w(2) + (lx >> 2)];
                    if (a.num_scales > 3) gsum += sm.cc[cc_off(3) + (r >> 3) * cc_w(3) + (lx >> 3)];
                    out[k] += gsum;
                }
            }

            if constexpr (SMOOTH) {
                const float p_l = prow[-1], p_r = prow[4];
                const float4 pu4 = *reinterpret_cast<const float4*>(prow - RW);
                const float4 pd4 = *reinterpret_cast<const float4*>(prow + RW);
                const float px[6] = {p_l, p[0], p[1], p[2], p[3], p_r};
                const float pu[4] = {pu4.x, pu4.y, pu4.z, pu4.w}, pd[4] = {pd4.x, pd4.y, pd4.z, pd4.w};
                // x-edge weights: wx[j] is the edge between px[j] and px[j+1], j = 0..4
                float wx[5];
#pragma unroll
                for (int j = 0; j < 5; ++j)
                    wx[j] = edge_w(Icur[0][j], Icur[0][j + 1], Icur[1][j], Icur[1][j + 1], Icur[2][j], Icur[2][j + 1]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int gx = gx0 + k;
                    const float wy_dn = edge_w(Icur[0][1 + k], Inxt[0][1 + k], Icur[1][1 + k], Inxt[1][1 + k],
                                               Icur[2][1 + k], Inxt[2][1 + k]);
                    if (row_in && gx < W) {
                        float gx_term = 0.f, gy_term = 0.f;
                        if (gx + 1 < W) {            // own right edge
                            const float d = px[k + 2] - px[k + 1];
                            acc[BF_SMX] += wx[k + 1] * fabsf(d);
                            gx_term -= wx[k + 1] * sgnf(d);
                        }
                        if (gx >= 1) gx_term += wx[k] * sgnf(px[k + 1] - px[k]);
                        if (gy + 1 < H) {            // own lower edge
                            const float d = pd[k] - p[k];
                            acc[BF_SMY] += wy_dn * fabsf(d);
                            gy_term -= wy_dn * sgnf(d);
                        }
                        if (gy >= 1) gy_term += wy_up[k] * sgnf(p[k] - pu[k]);
                        out[k] += a.w_smooth * ab * (gx_term * sm_nx + gy_term * sm_ny);
                    }
                    wy_up[k] = wy_dn;
                }
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int j = 0; j < 6; ++j) Icur[c][j] = Inxt[c][j];
            }

            if constexpr ((F & (FB_SI | FB_RP)) != 0) {
                RpGeom q{};
                if constexpr (F & FB_RP) {
                    q.fxe = fx + a.eps_rp;
                    q.fye = fy + a.eps_rp;
                    q.ay = (float)gy - cy;                      // depth_loss.h:300
                    q.yh = q.ay / q.fye;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (!row_in || gx0 + k >= W) continue;
                    float d_si = 0.f;
                    if constexpr (F & FB_SI) {
                        if (GRAD && a.eps_si == a.eps_grad) d_si = lp[k] - lg[k];
                        else d_si = logf(clampf(p[k], a.eps_si, 1000.0f)) - logf(clampf(g[k], a.eps_si, 1000.0f));
                    }
                    q.ax = axk[k];
                    q.xh = xhk[k];
                    out[k] += pointwise_px<F>(a, dv, p[k], g[k], has_mask, um[k] != 0, d_si, q, acc[BF_RP_E]);
                }
            }

            if (a.grad && row_in && gx0 < W) {
                float* o = a.grad + img + (size_t)gy * W + gx0;
                if (a.vec_ok && gx0 + 3 < W) {
                    *reinterpret_cast<float4*>(o) =
                        make_float4(out[0] * a.upstream, out[1] * a.upstream, out[2] * a.upstream, out[3] * a.upstream);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (gx0 + k < W) o[k] = out[k] * a.upstream;
                }
            }
        }
    }

    if (publish_partials(a, acc, tile, s_f, &s_last)) {
        finalize_results(a, s_d);
        if (a.metrics) write_metric_results(a.stats, a.metrics, *a.results, tid);
    }
}

// grad[b, :] -= off[b]   (the mean-normalisation term of the smoothness gradient)
__global__ void __launch_bounds__(256) smooth_offset_kernel(float* grad, const float* off, int HW, int vec_ok) {
    const int b = blockIdx.y;
    const float o = off[b];
    float* g = grad + (size_t)b * HW;
    if (vec_ok) {
        float4* g4 = reinterpret_cast<float4*>(g);
        const int n4 = HW >> 2;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
            float4 v = g4[i];
            v.x -= o; v.y -= o; v.z -= o; v.w -= o;
            g4[i] = v;
        }
    } else {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) g[i] -= o;
    }
}

// autograd backward: out = in * (*upstream); no traffic when upstream == 1 and in place
__global__ void __launch_bounds__(256) scale_grad_kernel(const float* in, const float* up, float* out, size_t n,
                                                         int vec_ok) {
    const float u = __ldg(up);
    if (u == 1.0f && in == out) return;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if (vec_ok) {
        const float4* i4 = reinterpret_cast<const float4*>(in);
        float4* o4 = reinterpret_cast<float4*>(out);
        const size_t n4 = n >> 2;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            float4 v = i4[i];
            v.x *= u; v.y *= u; v.z *= u; v.w *= u;
            o4[i] = v;
        }
        for (size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i] * u;
    } else {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i] * u;
    }
}

// metrics-only finalisation (no loss terms requested)
__global__ void metrics_finalize_kernel(const double* stats, uint32_t which, cadl_results* r) {
    if (blockIdx.x == 0) write_metric_results(stats, which, *r, threadIdx.x);   // launched with 32 threads
}

}  // namespace cadl
