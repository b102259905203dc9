// cadl phase A: one streaming pass over pred/gt that produces every batch-global scalar the
// gradient pass needs (SI n, sum d, sum d^2; reprojection n; per-image sum(pred) for the smoothness
// normaliser) and, fused into the same pass, both metric variants.  HBM-bound: 8 B/px, 128-bit
// loads, per-thread fp32 accumulators -> warp shuffle -> fp64 per-block partial rows -> the last
// block (ticket) reduces the rows in a fixed order (deterministic; no floating-point atomics).
#pragma once
#include "cadl_common.cuh"
#include "cadl_math.cuh"

namespace cadl {

// what a phase-A instantiation computes
constexpr int FA_SI = 1, FA_RP = 2, FA_PSUM = 4, FA_EV = 8, FA_TR = 16;

struct PhaseAArgs {
    const float* pred;
    const float* gt;
    const uint8_t* mask;
    int B, HW;
    int blocks_per_img;
    int vec_ok;
    float eps_si, eps_rp, min_d, max_d;
    WsHeader* hdr;
    double* stats;
    double* img_psum;
    double* a_part;
};

// delta-threshold counting with the reference's IEEE semantics at SFU cost: the quotients come from
// rcp.approx (<= 2 ulp), and only when max(p/g, g/p) lies within a guard band of a threshold (probability
// ~1e-5 per pixel) are the two IEEE divisions of depth_metrics.h:221 / trainer :433 actually performed.
__device__ __forceinline__ float ratio_for_thresholds(float p, float g, float rp, float rg) {
    float ratio = fmaxf(p * rg, g * rp);
    const bool near = (fabsf(ratio - 1.25f) < 4e-6f) || (fabsf(ratio - 1.5625f) < 5e-6f) ||
                      (fabsf(ratio - 1.953125f) < 6e-6f) || !(ratio == ratio);
    if (near) ratio = fmaxf(__fdiv_rn(p, g), __fdiv_rn(g, p));
    return ratio;
}

// One pair of pixels (packed fp32x2 logs).  lpv/lgv: logs of clamp(x, eps_si, 1000) when F needs them.
template <int F>
__device__ __forceinline__ void phase_a_px(float p, float g, float lp, float lg, bool has_mask, bool um,
                                           const PhaseAArgs& a, float (&af)[AF_COUNT], unsigned (&ai)[AI_COUNT]) {
    if constexpr (F & FA_PSUM) af[AF_PSUM] += p;            // depth_loss.h:192 (mean over H,W)
    if constexpr (F & FA_SI) {
        // depth_loss.h:38-47
        const bool m = has_mask ? um : (g > a.eps_si);
        const float d = lp - lg;
        if (m) {
            ai[AI_SI_N] += 1u;
            af[AF_SI_S] += d;
            af[AF_SI_Q] += d * d;
        }
    }
    if constexpr (F & FA_RP) {
        const bool m = has_mask ? um : (g > a.eps_rp);      // depth_loss.h:318-320
        if (m) ai[AI_RP_N] += 1u;
    }
    if constexpr ((F & (FA_EV | FA_TR)) != 0) {
        // Both metric variants share their per-pixel terms whenever clamping leaves pred unchanged and
        // x + 1e-8f == x; the general (rare) cases are evaluated separately below.
        const float pc = clamp_nan(p, a.min_d, a.max_d);                       // depth_metrics.h:66
        const float psi = clamp_nan(p, a.eps_si, 1000.0f), gsi = clamp_nan(g, a.eps_si, 1000.0f);
        const bool ev_ok = (F & FA_EV) && (g > a.min_d) && (g < a.max_d) && (has_mask ? um : true);   // :154-161
        const bool tr_ok = (F & FA_TR) && (g > 0.0f);                          // trainer :410
        if (ev_ok || tr_ok) {
            const float rg = rcp_approx(g);
            // ---- eval terms (pred clamped) ----
            const float lpc = (pc == psi) ? lp : logf(pc);
            const float lge = (g == gsi) ? lg : logf(g);
            const float diff = pc - g, ad = fabsf(diff), sq = diff * diff;     // torch::pow(x,2) == x*x
            const float ld = lpc - lge;
            const float ratio = ratio_for_thresholds(pc, g, rcp_approx(pc), rg);
            if (ev_ok) {
                af[AF_EV_ABSREL] += ad * rg;                                   // :170
                af[AF_EV_SQREL] += sq * rg;                                    // :177
                af[AF_EV_SQ] += sq;                                            // :184
                af[AF_EV_LOGSQ] += ld * ld;                                    // :191-192
                af[AF_EV_ABS] += ad;                                           // :199
                af[AF_EV_LOG10] += fabsf(ld) * 0.43429448190325182765f;        // :206 (log10 x = ln x / ln 10)
                ai[AI_EV_N] += 1u;
                ai[AI_EV_C1] += (ratio < 1.25f) ? 1u : 0u;                     // :224-229
                ai[AI_EV_C2] += (ratio < 1.25f * 1.25f) ? 1u : 0u;
                ai[AI_EV_C3] += (ratio < 1.25f * 1.25f * 1.25f) ? 1u : 0u;
                af[AF_EV_SUMP] += pc;                                          // :84
                af[AF_EV_SUMG] += g;                                           // :85
            }
            if (tr_ok) {
                // trainer :419-436: no clamp, log(x + 1e-8)
                const float p8 = p + 1e-8f, g8 = g + 1e-8f;
                float adt = ad, sqt = sq, ldt = ld, rt = ratio;
                if (!(pc == p) || !(p8 == psi) || !(g8 == gsi)) {              // rare: recompute unshared
                    adt = fabsf(p - g);
                    sqt = adt * adt;
                    ldt = logf(p8) - logf(g8);
                    rt = fmaxf(__fdiv_rn(p, g), __fdiv_rn(g, p));
                }
                af[AF_TR_ABSREL] += adt * rg;
                af[AF_TR_SQREL] += sqt * rg;
                af[AF_TR_SQ] += sqt;
                af[AF_TR_LOGSQ] += ldt * ldt;
                ai[AI_TR_N] += 1u;
                ai[AI_TR_C1] += (rt < 1.25f) ? 1u : 0u;
                ai[AI_TR_C2] += (rt < 1.5625f) ? 1u : 0u;
                ai[AI_TR_C3] += (rt < 1.953125f) ? 1u : 0u;
            }
        }
    }
}

template <int F>
__device__ __forceinline__ void phase_a_quad(const float (&p)[4], const float (&g)[4], bool has_mask,
                                             const bool (&um)[4], const PhaseAArgs& a, float (&af)[AF_COUNT],
                                             unsigned (&ai)[AI_COUNT]) {
    float lp[4] = {0.f, 0.f, 0.f, 0.f}, lg[4] = {0.f, 0.f, 0.f, 0.f};
    if constexpr ((F & (FA_SI | FA_EV | FA_TR)) != 0) {
#pragma unroll
        for (int k = 0; k < 4; k += 2) {
            const float2 a2 = log_exact2(make_float2(clamp_nan(p[k], a.eps_si, 1000.0f), clamp_nan(p[k + 1], a.eps_si, 1000.0f)));
            const float2 b2 = log_exact2(make_float2(clamp_nan(g[k], a.eps_si, 1000.0f), clamp_nan(g[k + 1], a.eps_si, 1000.0f)));
            lp[k] = a2.x; lp[k + 1] = a2.y;
            lg[k] = b2.x; lg[k + 1] = b2.y;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) phase_a_px<F>(p[k], g[k], lp[k], lg[k], has_mask, um[k], a, af, ai);
}

// Deterministic block-wide sum of one double per thread (fixed shuffle/tree order).
__device__ __forceinline__ double block_sum_double(double v, double* scratch /*>=8*/) {
    v = warp_sum(v);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (warp == 0) {
        r = (lane < (int)(blockDim.x >> 5)) ? scratch[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;  // valid in warp 0
}

template <int F>
__global__ void __launch_bounds__(kThreadsA) phase_a_kernel(const PhaseAArgs a) {
    constexpr bool NEED_P = (F & (FA_SI | FA_PSUM | FA_EV | FA_TR)) != 0;
    constexpr bool NEED_G = (F & (FA_SI | FA_RP | FA_EV | FA_TR)) != 0;
    __shared__ float s_f[kThreadsA / 32][AF_COUNT];
    __shared__ unsigned s_i[AI_COUNT];
    __shared__ double s_d[8];
    __shared__ int s_last;

    const int b = blockIdx.y, k = blockIdx.x;
    const int tid = threadIdx.x;
    const bool has_mask = a.mask != nullptr;
    const size_t base = (size_t)b * a.HW;

    float af[AF_COUNT];
    unsigned ai[AI_COUNT];
#pragma unroll
    for (int q = 0; q < AF_COUNT; ++q) af[q] = 0.f;
#pragma unroll
    for (int q = 0; q < AI_COUNT; ++q) ai[q] = 0u;
    if (tid < AI_COUNT) s_i[tid] = 0u;

    if (a.vec_ok) {
        const int nvec = a.HW >> 2;
        const int v0 = (int)((long long)nvec * k / a.blocks_per_img);
        const int v1 = (int)((long long)nvec * (k + 1) / a.blocks_per_img);
        const float4* p4 = reinterpret_cast<const float4*>(a.pred + base);
        const float4* g4 = reinterpret_cast<const float4*>(a.gt + base);
        const uchar4* m4 = reinterpret_cast<const uchar4*>(a.mask ? a.mask + base : nullptr);
        for (int i = v0 + tid; i < v1; i += 2 * kThreadsA) {
            const int j = i + kThreadsA;
            const bool hj = j < v1;
            float4 p0 = make_float4(0, 0, 0, 0), p1 = p0, g0 = p0, g1 = p0;
            uchar4 u0 = make_uchar4(1, 1, 1, 1), u1 = u0;
            if (NEED_P) { p0 = __ldg(p4 + i); if (hj) p1 = __ldg(p4 + j); }
            if (NEED_G) { g0 = __ldg(g4 + i); if (hj) g1 = __ldg(g4 + j); }
            if (has_mask) { u0 = __ldg(m4 + i); if (hj) u1 = __ldg(m4 + j); }
            {
                const float pp[4] = {p0.x, p0.y, p0.z, p0.w}, gg[4] = {g0.x, g0.y, g0.z, g0.w};
                const bool mm[4] = {u0.x != 0, u0.y != 0, u0.z != 0, u0.w != 0};
                phase_a_quad<F>(pp, gg, has_mask, mm, a, af, ai);
            }
            if (hj) {
                const float pp[4] = {p1.x, p1.y, p1.z, p1.w}, gg[4] = {g1.x, g1.y, g1.z, g1.w};
                const bool mm[4] = {u1.x != 0, u1.y != 0, u1.z != 0, u1.w != 0};
                phase_a_quad<F>(pp, gg, has_mask, mm, a, af, ai);
            }
        }
    } else {
        const int e0 = (int)((long long)a.HW * k / a.blocks_per_img);
        const int e1 = (int)((long long)a.HW * (k + 1) / a.blocks_per_img);
        for (int i = e0 + tid; i < e1; i += kThreadsA) {
            float p = NEED_P ? __ldg(a.pred + base + i) : 0.f;
            float g = NEED_G ? __ldg(a.gt + base + i) : 0.f;
            bool um = has_mask ? (__ldg(a.mask + base + i) != 0) : true;
            float lp = 0.f, lg = 0.f;
            if constexpr ((F & (FA_SI | FA_EV | FA_TR)) != 0) {
                lp = log_exact(clamp_nan(p, a.eps_si, 1000.0f));
                lg = log_exact(clamp_nan(g, a.eps_si, 1000.0f));
            }
            phase_a_px<F>(p, g, lp, lg, has_mask, um, a, af, ai);
        }
    }

    // ---- block reduction: fp32 warp shuffle, then fp64 across warps in warp order ----
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int q = 0; q < AF_COUNT; ++q) {
        float v = warp_sum(af[q]);
        if (lane == 0) s_f[warp][q] = v;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < AI_COUNT; ++q) {
        unsigned v = warp_sum(ai[q]);
        if (lane == 0 && v) atomicAdd(&s_i[q], v);
    }
    const int blk = b * a.blocks_per_img + k;
    if (tid < AF_COUNT) {
        double acc = 0.0;
#pragma unroll
        for (int w = 0; w < kThreadsA / 32; ++w) acc += (double)s_f[w][tid];
        a.a_part[(size_t)blk * AF_COUNT + tid] = acc;
    }
    __syncthreads();
    if (tid < AI_COUNT && s_i[tid]) atomicAdd(&a.hdr->icount[tid], (unsigned long long)s_i[tid]);

    // ---- ticket: the last block to arrive reduces all partial rows in a fixed order ----
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned total = gridDim.x * gridDim.y;
        unsigned t = atomicAdd(&a.hdr->ticket_a, 1u);
        s_last = (t == total - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    const int nblk = a.B * a.blocks_per_img;
    const volatile double* part = a.a_part;
    for (int q = 0; q < AF_COUNT; ++q) {
        if (q == AF_PSUM) continue;
        double acc = 0.0;
        for (int i = tid; i < nblk; i += kThreadsA) acc += part[(size_t)i * AF_COUNT + q];
        double r = block_sum_double(acc, s_d);
        if (tid == 0) {
            int st = -1;
            switch (q) {
                case AF_SI_S: st = ST_SI_S; break;
                case AF_SI_Q: st = ST_SI_Q; break;
                case AF_EV_ABSREL: st = ST_EV_ABSREL; break;
                case AF_EV_SQREL: st = ST_EV_SQREL; break;
                case AF_EV_SQ: st = ST_EV_SQ; break;
                case AF_EV_LOGSQ: st = ST_EV_LOGSQ; break;
                case AF_EV_ABS: st = ST_EV_ABS; break;
                case AF_EV_LOG10: st = ST_EV_LOG10; break;
                case AF_EV_SUMP: st = ST_EV_SUMP; break;
                case AF_EV_SUMG: st = ST_EV_SUMG; break;
                case AF_TR_ABSREL: st = ST_TR_ABSREL; break;
                case AF_TR_SQREL: st = ST_TR_SQREL; break;
                case AF_TR_SQ: st = ST_TR_SQ; break;
                case AF_TR_LOGSQ: st = ST_TR_LOGSQ; break;
                default: break;
            }
            if (st >= 0) a.stats[st] = r;
        }
    }
    // per-image sum(pred): the blocks of image b are contiguous rows
    for (int img = warp; img < a.B; img += kThreadsA / 32) {
        double acc = 0.0;
        for (int i = lane; i < a.blocks_per_img; i += 32)
            acc += part[((size_t)img * a.blocks_per_img + i) * AF_COUNT + AF_PSUM];
        acc = warp_sum(acc);
        if (lane == 0) a.img_psum[img] = acc;
    }
    if (tid < AI_COUNT) {
        unsigned long long c = atomicExch(&a.hdr->icount[tid], 0ull);  // read + leave clean
        int st = -1;
        switch (tid) {
            case AI_SI_N: st = ST_SI_N; break;
            case AI_RP_N: st = ST_RP_N; break;
            case AI_EV_N: st = ST_EV_N; break;
            case AI_EV_C1: st = ST_EV_C1; break;
            case AI_EV_C2: st = ST_EV_C2; break;
            case AI_EV_C3: st = ST_EV_C3; break;
            case AI_TR_N: st = ST_TR_N; break;
            case AI_TR_C1: st = ST_TR_C1; break;
            case AI_TR_C2: st = ST_TR_C2; break;
            case AI_TR_C3: st = ST_TR_C3; break;
            default: break;
        }
        if (st >= 0) a.stats[st] = (double)c;
    }
    if (tid == 0) a.hdr->ticket_a = 0u;
}

}  // namespace cadl
