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
    unsigned near_lo[3], near_span[3];   // guard bands around k log2(1.25) as bit patterns: exact delta counts
    WsHeader* hdr;
    double* stats;
    double* img_psum;
    double* a_part;
};

// Logs in phase A feed sums only (nothing here decides a sign), so they are lg2.approx (one MUFU) in log2 units and
// the per-block sums are converted once: x ln2, x ln2^2, x log10(2).  Round 1 evaluated a logf replica per value
// (~16 instructions each); the 1e-5 tolerance of the losses and metrics leaves four orders of magnitude of room.
constexpr float kLn2f = 0.693147180559945309f;
constexpr float kLog2_125 = 0.32192809488736235f;        // log2(1.25): the delta thresholds are 1.25^k

// Per-thread accumulators of phase A.  Both metric variants take the same per-pixel terms whenever
// min < gt < max and pred needs no clamping (the trainers' +1e-8 inside the log is applied as an exact
// first-order correction): those pixels go to ONE set of "common" accumulators.  The few others
// (gt outside (min,max) but > 0, or pred outside [min,max]) are evaluated separately for the trainer variant
// and kept in per-thread shared-memory slots so they cost no registers.
struct AccA {
    float psum, si_s, si_q;                          // si_s, si_q in log2 units
    unsigned si_n, rp_n;
    float c_absrel, c_sqrel, c_sq, c_logsq;          // eval == train   (c_logsq in log2^2 units)
    unsigned c_n, c_c1, c_c2, c_c3;
    float e_abs, e_l10, e_sump, e_sumg;              // eval-only quantities (e_l10 in log2 units)
    float t_logsq;                                   // train: correction of sum ld^2 for the +1e-8 inside the logs (natural units)
};
constexpr int kToSlots = 8;   // train-only slow path: absrel, sqrel, sq, logsq, n, c1, c2, c3

template <int F, bool HAS_MASK>
__device__ __forceinline__ void phase_a_px(float p, float g, bool um, const PhaseAArgs& a, AccA& A, float* s_to) {
    if constexpr (F & FA_PSUM) A.psum += p;                 // depth_loss.h:192 (mean over H,W)
    if constexpr (F & FA_SI) {
        // depth_loss.h:38-47
        const bool m = HAS_MASK ? um : (g > a.eps_si);
        const float d2 = lg2_approx(clamp_nan(p, a.eps_si, 1000.0f)) - lg2_approx(clamp_nan(g, a.eps_si, 1000.0f));
        const float d = m ? d2 : 0.f;
        A.si_n += m ? 1u : 0u;
        A.si_s += d;
        A.si_q = fmaf(d, d, A.si_q);
    }
    if constexpr (F & FA_RP) {
        const bool m = HAS_MASK ? um : (g > a.eps_rp);      // depth_loss.h:318-320
        if (m) A.rp_n += 1u;
    }
    if constexpr ((F & (FA_EV | FA_TR)) != 0) {
        constexpr bool EV = (F & FA_EV) != 0, TR = (F & FA_TR) != 0;
        // eval terms, pred clamped AFTER masking (depth_metrics.h:154-161, :66)
        const bool ev_ok = (g > a.min_d) && (g < a.max_d) && (HAS_MASK ? um : true);
        const float pc = clamp_nan(p, a.min_d, a.max_d);
        const float rg = rcp_approx(g);
        const float diff = pc - g, ad = fabsf(diff), sq = diff * diff;              // :170-199 (pow(x,2) == x*x)
        const float ld = lg2_approx(pc) - lg2_approx(g);                            // log2 units; g > min_d > 0 where it counts
        if (ev_ok) {
            // delta thresholds (:221-229, trainer :433): max(p/g, g/p) < 1.25^k  <=>  |log2 p - log2 g| < k log2 1.25.
            // Only when |ld| lies within a guard band of a threshold (the error of the two approximate logs; ~1e-5
            // of the pixels) are the reference's two IEEE divisions actually performed.
            const float ald = fabsf(ld);
            const unsigned u = (unsigned)__float_as_int(ald);
            const bool near = (u - a.near_lo[0] <= a.near_span[0]) || (u - a.near_lo[1] <= a.near_span[1]) ||
                              (u - a.near_lo[2] <= a.near_span[2]);
            bool b1 = ald < kLog2_125, b2 = ald < 2.f * kLog2_125, b3 = ald < 3.f * kLog2_125;
            if (near) {
                const float ratio = fmaxf(__fdiv_rn(pc, g), __fdiv_rn(g, pc));
                b1 = ratio < 1.25f; b2 = ratio < 1.5625f; b3 = ratio < 1.953125f;
            }
            A.c_absrel = fmaf(ad, rg, A.c_absrel);
            A.c_sqrel = fmaf(sq, rg, A.c_sqrel);
            A.c_sq += sq;
            A.c_logsq = fmaf(ld, ld, A.c_logsq);
            A.c_n += 1u;
            A.c_c1 += b1 ? 1u : 0u;
            A.c_c2 += b2 ? 1u : 0u;
            A.c_c3 += b3 ? 1u : 0u;
            if constexpr (EV) {
                A.e_abs += ad;
                A.e_l10 += ald;                                                     // x log10(2) at the end (:206)
                A.e_sump += pc;
                A.e_sumg += g;
            }
            if constexpr (TR) {
                // trainer :429: log(x + 1e-8).  Only x < 0.25 changes under +1e-8f (half an ulp of 0.25 is 1.5e-8);
                // there (x + 1e-8f) - x is exact and log(x + dx) = log x + dx/x to far below one ulp, so the train
                // variant adds  (ld + c)^2 - ld^2 = c (2 ld + c)  to the common sum (natural-log units).  Rare branch.
                if ((p < 0.25f || g < 0.25f) && pc == p) {
                    const float dp = (p + 1e-8f) - p, dg = (g + 1e-8f) - g;
                    const float c = fmaf(dp, rcp_approx(p), -dg * rg);
                    A.t_logsq = fmaf(c, fmaf(2.f * kLn2f, ld, c), A.t_logsq);
                }
            }
        }
        if constexpr (TR) {
            // trainer-only pixels: gt > 0 outside (min,max), or pred clamped by the eval variant
            const bool slow = (g > 0.0f) && (!ev_ok || !(pc == p));                  // trainer :410
            if (slow) {
                const float adt = fabsf(p - g), sqt = adt * adt;                     // :419-436, no clamp
                const float ldt = logf(p + 1e-8f) - logf(g + 1e-8f);
                const float rt = fmaxf(__fdiv_rn(p, g), __fdiv_rn(g, p));
                const int t = threadIdx.x;
                // "to" = train-only additions; where the pixel was also counted as common (ev_ok, clamped
                // pred) the common contribution is taken back out so train = common + to stays exact in form
                const float sgn = ev_ok ? 1.f : 0.f;
                const float ldn = ld * kLn2f;                                        // what the common sum received (natural units)
                s_to[0 * kThreadsA + t] += adt * rg - sgn * ad * rg;
                s_to[1 * kThreadsA + t] += sqt * rg - sgn * sq * rg;
                s_to[2 * kThreadsA + t] += sqt - sgn * sq;
                s_to[3 * kThreadsA + t] += ldt * ldt - sgn * ldn * ldn;
                const float rc = ev_ok ? fmaxf(__fdiv_rn(pc, g), __fdiv_rn(g, pc)) : 3.0f;
                s_to[4 * kThreadsA + t] += 1.f - sgn;
                s_to[5 * kThreadsA + t] += ((rt < 1.25f) ? 1.f : 0.f) - sgn * ((rc < 1.25f) ? 1.f : 0.f);
                s_to[6 * kThreadsA + t] += ((rt < 1.5625f) ? 1.f : 0.f) - sgn * ((rc < 1.5625f) ? 1.f : 0.f);
                s_to[7 * kThreadsA + t] += ((rt < 1.953125f) ? 1.f : 0.f) - sgn * ((rc < 1.953125f) ? 1.f : 0.f);
            }
        }
    }
}

template <int F, bool HAS_MASK>
__device__ __forceinline__ void phase_a_quad(const float (&p)[4], const float (&g)[4], const bool (&um)[4],
                                             const PhaseAArgs& a, AccA& A, float* s_to) {
#pragma unroll
    for (int k = 0; k < 4; ++k) phase_a_px<F, HAS_MASK>(p[k], g[k], um[k], a, A, s_to);
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

template <int F, bool HAS_MASK>
__global__ void __launch_bounds__(kThreadsA, 4) phase_a_kernel(const PhaseAArgs a) {
    constexpr bool NEED_P = (F & (FA_SI | FA_PSUM | FA_EV | FA_TR)) != 0;
    constexpr bool NEED_G = (F & (FA_SI | FA_RP | FA_EV | FA_TR)) != 0;
    __shared__ float s_f[kThreadsA / 32][AF_COUNT];
    __shared__ int s_last;
    __shared__ float s_to[kToSlots * kThreadsA];
    __shared__ unsigned s_iw[kThreadsA / 32][AI_COUNT];

    pdl_trigger();   // the pooled-pyramid kernel that follows reads only pred/gt: let it fill SMs as this grid drains
    const int b = blockIdx.y, k = blockIdx.x;
    const int tid = threadIdx.x;
    constexpr bool has_mask = HAS_MASK;
    const size_t base = (size_t)b * a.HW;

    AccA A;
    memset(&A, 0, sizeof(A));
#pragma unroll
    for (int q = 0; q < kToSlots; ++q) s_to[q * kThreadsA + tid] = 0.f;   // private slots: no sync needed

    if (a.vec_ok) {
        const int nvec = a.HW >> 2;
        const int v0 = (int)((long long)nvec * k / a.blocks_per_img);
        const int v1 = (int)((long long)nvec * (k + 1) / a.blocks_per_img);
        const float4* p4 = reinterpret_cast<const float4*>(a.pred + base);
        const float4* g4 = reinterpret_cast<const float4*>(a.gt + base);
        const uchar4* m4 = reinterpret_cast<const uchar4*>(a.mask ? a.mask + base : nullptr);
        // UNR independent 128-bit loads per tensor in flight per thread (light variants need more to cover DRAM latency)
        constexpr int UNR = (F & (FA_EV | FA_TR)) ? 2 : ((F == FA_RP) ? 8 : 4);
        for (int i = v0 + tid; i < v1; i += UNR * kThreadsA) {
            float4 pv[UNR], gv[UNR];
            uchar4 uv[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int j = i + u * kThreadsA;
                pv[u] = make_float4(0, 0, 0, 0); gv[u] = pv[u]; uv[u] = make_uchar4(1, 1, 1, 1);
                if (j < v1) {
                    if (NEED_P) pv[u] = __ldg(p4 + j);
                    if (NEED_G) gv[u] = __ldg(g4 + j);
                    if (has_mask) uv[u] = __ldg(m4 + j);
                }
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                if (i + u * kThreadsA < v1) {
                    const float pp[4] = {pv[u].x, pv[u].y, pv[u].z, pv[u].w}, gg[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
                    const bool mm[4] = {uv[u].x != 0, uv[u].y != 0, uv[u].z != 0, uv[u].w != 0};
                    phase_a_quad<F, HAS_MASK>(pp, gg, mm, a, A, s_to);
                }
            }
        }
    } else {
        const int e0 = (int)((long long)a.HW * k / a.blocks_per_img);
        const int e1 = (int)((long long)a.HW * (k + 1) / a.blocks_per_img);
        for (int i = e0 + tid; i < e1; i += kThreadsA) {
            float p = NEED_P ? __ldg(a.pred + base + i) : 0.f;
            float g = NEED_G ? __ldg(a.gt + base + i) : 0.f;
            bool um = has_mask ? (__ldg(a.mask + base + i) != 0) : true;
            phase_a_px<F, HAS_MASK>(p, g, um, a, A, s_to);
        }
    }

    // ---- fold the common / train-only accumulators into the per-variant sums ----
    float af[AF_COUNT];
    unsigned ai[AI_COUNT];
    constexpr bool EVc = (F & FA_EV) != 0, TRc = (F & FA_TR) != 0;
    af[AF_SI_S] = A.si_s * kLn2f; af[AF_SI_Q] = A.si_q * (kLn2f * kLn2f); af[AF_PSUM] = A.psum;      // log2 -> natural units
    af[AF_EV_ABSREL] = EVc ? A.c_absrel : 0.f; af[AF_EV_SQREL] = EVc ? A.c_sqrel : 0.f;
    af[AF_EV_SQ] = EVc ? A.c_sq : 0.f; af[AF_EV_LOGSQ] = EVc ? A.c_logsq * (kLn2f * kLn2f) : 0.f;
    af[AF_EV_ABS] = A.e_abs; af[AF_EV_LOG10] = A.e_l10 * 0.30102999566398119521f;   // log10 x = log2 x * log10(2)
    af[AF_EV_SUMP] = A.e_sump; af[AF_EV_SUMG] = A.e_sumg;
    af[AF_TR_ABSREL] = TRc ? A.c_absrel + s_to[0 * kThreadsA + tid] : 0.f;
    af[AF_TR_SQREL] = TRc ? A.c_sqrel + s_to[1 * kThreadsA + tid] : 0.f;
    af[AF_TR_SQ] = TRc ? A.c_sq + s_to[2 * kThreadsA + tid] : 0.f;
    af[AF_TR_LOGSQ] = TRc ? A.c_logsq * (kLn2f * kLn2f) + A.t_logsq + s_to[3 * kThreadsA + tid] : 0.f;
    ai[AI_SI_N] = A.si_n; ai[AI_RP_N] = A.rp_n;
    ai[AI_EV_N] = EVc ? A.c_n : 0u; ai[AI_EV_C1] = EVc ? A.c_c1 : 0u;
    ai[AI_EV_C2] = EVc ? A.c_c2 : 0u; ai[AI_EV_C3] = EVc ? A.c_c3 : 0u;
    // the train-only slots hold small signed integers as floats (exact)
    ai[AI_TR_N] = TRc ? (unsigned)((int)A.c_n + (int)s_to[4 * kThreadsA + tid]) : 0u;
    ai[AI_TR_C1] = TRc ? (unsigned)((int)A.c_c1 + (int)s_to[5 * kThreadsA + tid]) : 0u;
    ai[AI_TR_C2] = TRc ? (unsigned)((int)A.c_c2 + (int)s_to[6 * kThreadsA + tid]) : 0u;
    ai[AI_TR_C3] = TRc ? (unsigned)((int)A.c_c3 + (int)s_to[7 * kThreadsA + tid]) : 0u;

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
        if (lane == 0) s_iw[warp][q] = v;
    }
    const int blk = b * a.blocks_per_img + k;
    if ((F & ~FA_RP) && tid < AF_COUNT) {
        double acc = 0.0;
#pragma unroll
        for (int w = 0; w < kThreadsA / 32; ++w) acc += (double)s_f[w][tid];
        a.a_part[(size_t)blk * AF_COUNT + tid] = acc;
    }
    __syncthreads();
    if (tid < AI_COUNT) {
        unsigned t = 0;
#pragma unroll
        for (int w = 0; w < kThreadsA / 32; ++w) t += s_iw[w][tid];
        if (t) atomicAdd(&a.hdr->icount[tid], (unsigned long long)t);
    }

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
    const double* part = a.a_part;
    // One partial row per thread and trip, two halves of the quantities (register budget): every load of a trip is
    // independent -- a dependent chain of L2 round trips per quantity cost ~15 us.  Fixed thread-strided order,
    // fixed shuffle tree, fixed warp order: deterministic.  __ldcg: rows written by other SMs (fence + ticket).
    __shared__ double s_wa[kThreadsA / 32][AF_COUNT];
    constexpr int kHalf = (AF_COUNT + 1) / 2;
    if constexpr ((F & ~FA_RP) != 0)          // the reprojection count alone has no float partial rows
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        double accq[kHalf];
#pragma unroll
        for (int q = 0; q < kHalf; ++q) accq[q] = 0.0;
        for (int i = tid; i < nblk; i += 3 * kThreadsA) {       // three rows per trip: 24 independent loads in flight
            const int i2 = i + kThreadsA, i3 = i + 2 * kThreadsA;
            double r0[kHalf], r1[kHalf], r2[kHalf];
#pragma unroll
            for (int q = 0; q < kHalf; ++q) {
                const bool on = h * kHalf + q < AF_COUNT;
                r0[q] = on ? __ldcg(part + (size_t)i * AF_COUNT + h * kHalf + q) : 0.0;
                r1[q] = (on && i2 < nblk) ? __ldcg(part + (size_t)i2 * AF_COUNT + h * kHalf + q) : 0.0;
                r2[q] = (on && i3 < nblk) ? __ldcg(part + (size_t)i3 * AF_COUNT + h * kHalf + q) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < kHalf; ++q) accq[q] = ((accq[q] + r0[q]) + r1[q]) + r2[q];
        }
#pragma unroll
        for (int q = 0; q < kHalf; ++q) {
            if (h * kHalf + q < AF_COUNT) {
                const double v = warp_sum(accq[q]);
                if (lane == 0) s_wa[warp][h * kHalf + q] = v;
            }
        }
    }
    __syncthreads();
    if (tid < AF_COUNT) {
        const int q = tid;
        bool on = q != AF_PSUM;
        if (!(F & FA_SI) && (q == AF_SI_S || q == AF_SI_Q)) on = false;
        if (!(F & FA_EV) && q >= AF_EV_ABSREL && q <= AF_EV_SUMG) on = false;
        if (!(F & FA_TR) && q >= AF_TR_ABSREL && q <= AF_TR_LOGSQ) on = false;
        if (on) {
            double r = 0.0;
#pragma unroll
            for (int w = 0; w < kThreadsA / 32; ++w) r += s_wa[w][q];
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
    if constexpr (F & FA_PSUM)
    for (int img = warp; img < a.B; img += kThreadsA / 32) {
        double acc = 0.0;
#pragma unroll 4
        for (int i = lane; i < a.blocks_per_img; i += 32)
            acc += __ldcg(part + ((size_t)img * a.blocks_per_img + i) * AF_COUNT + AF_PSUM);
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
        // (only what this instantiation counts: a metrics-only pass may run beside the loss statistics' pass)
        bool mine = true;
        if (!(F & FA_SI) && tid == AI_SI_N) mine = false;
        if (!(F & FA_RP) && tid == AI_RP_N) mine = false;
        if (!(F & FA_EV) && tid >= AI_EV_N && tid <= AI_EV_C3) mine = false;
        if (!(F & FA_TR) && tid >= AI_TR_N && tid <= AI_TR_C3) mine = false;
        if (st >= 0 && mine) a.stats[st] = (double)c;
    }
    if (tid == 0) a.hdr->ticket_a = 0u;
}

}  // namespace cadl
