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
// Statistics the pooled-sum pass can produce on its way (it reads every pred/gt value anyway): what phase A would
// compute for the loss terms -- SI n / sum d / sum d^2 (the reprojection count is the same mask on this path) and the
// per-image sum(pred) of the smoothness normaliser.  With them the loss step needs no phase A at all; the metric
// variants, if asked for, run as a metrics-only phase A on the auxiliary stream (cadl_api.cu: cadl_stack_fwd_bwd).
constexpr int PS_SI = 1, PS_PSUM = 2;
struct PoolStatsRec {               // fixed offset in the workspace header; zero between calls
    unsigned long long n;           // valid pixels                                   depth_loss.h:52, :323
    unsigned long long s_hi, s_lo;  // sum d,   log2 units, signed fixed point        depth_loss.h:61
    unsigned long long q_hi, q_lo;  // sum d^2, log2^2 units                          depth_loss.h:58
    unsigned int ticket, flags;     // flags: fix_split bits of quantity 0 (s) and 1 (q)
};
struct PoolStatsArgs {
    const uint8_t* mask;
    PoolStatsRec* rec;
    unsigned long long* img_words;  // per image: hi, lo, flags, pad (sum of pred, signed fixed point); zero between calls
    double* stats;                  // ST_* vector: the last CTA writes ST_SI_N, ST_SI_S, ST_SI_Q (and ST_RP_N when want_rp)
    double* img_psum;               // per-image sum(pred), written by the last CTA
    int want_rp;
};
// signed variant of fix_split: hi is a two's-complement count of 2^-16 units (unsigned wrap-around addition is exact)
__device__ __forceinline__ void fix_split_signed(double v, unsigned long long& hi, unsigned long long& lo, unsigned& flag, int q) {
    hi = 0ull; lo = 0ull;
    if (!(fabs(v) < 1.0e14)) { flag |= (v != v) ? (1u << q) : (v > 0.0 ? (1u << (8 + q)) : (1u << (16 + q))); return; }
    const double sc = v * 65536.0, h = floor(sc);
    hi = (unsigned long long)(long long)h;
    lo = (unsigned long long)__double2ll_rn((sc - h) * 1099511627776.0);          // 2^40: lo in 2^-56 units
}
__device__ __forceinline__ double fix_join_signed(unsigned long long hi, unsigned long long lo, unsigned flags, int q) {
    if (flags & (1u << q)) return __longlong_as_double(0x7ff8000000000000ll);
    if ((flags & (1u << (8 + q))) && (flags & (1u << (16 + q)))) return __longlong_as_double(0x7ff8000000000000ll);   // inf - inf
    if (flags & (1u << (8 + q))) return __longlong_as_double(0x7ff0000000000000ll);
    if (flags & (1u << (16 + q))) return __longlong_as_double(0xfff0000000000000ll);
    return (double)(long long)hi * (1.0 / 65536.0) + (double)lo * (1.0 / 72057594037927936.0);
}

template <int SF, bool HAS_MASK>
__global__ void __launch_bounds__(256, CADL_PYR_MINB) pyr_pool_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                       int B, int H, int W, float eps, PyrArrays py,
                                                       unsigned int* img_rec_words, const PoolStatsArgs ps) {
    // Launched with programmatic stream serialization: the pooling loop needs nothing from the kernel before it in the
    // stream (the previous step's gradient pass or metrics pass, or phase A) and runs while that one drains.  Everything
    // that kernel may still read or write -- the per-image records, the statistics -- is touched only behind the
    // pdl_wait() below, which also makes this grid END after its predecessor: the kernels behind it wait for THIS grid.
    pdl_trigger();
    const int W8 = W >> 3, H8 = H >> 3;
    unsigned si_n = 0u;
    float si_s = 0.f, si_q = 0.f;
    const int lane = threadIdx.x & 31;
    const int total = B * H8 * W8;
    // grid-stride over the 8x8 blocks: the grid is either one thread per block or, when the kernel runs beside
    // phase A on a second stream, one CTA per SM
    // (warp-uniform trips: with statistics the warp reduces its per-image sums together; a lane past the end works on
    //  the last block again and neither stores nor counts)
    for (int idx0 = blockIdx.x * blockDim.x + threadIdx.x - lane; idx0 < total; idx0 += gridDim.x * blockDim.x) {
    const bool act = idx0 + lane < total;
    if (SF == 0 && !act) break;
    const int idx = act ? idx0 + lane : total - 1;
    const int bx = idx % W8, by = (idx / W8) % H8, b = idx / (W8 * H8);
    float bsum = 0.f;
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
        if constexpr (SF != 0) {
            unsigned long long mbits = 0ull;
            if constexpr (HAS_MASK) mbits = __ldg(reinterpret_cast<const unsigned long long*>(ps.mask + (size_t)b * H * W + (size_t)(by * 8 + r) * W + bx * 8));
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if constexpr (SF & PS_PSUM) bsum += vp[c];                                  // depth_loss.h:192
                if constexpr (SF & PS_SI) {                                                 // depth_loss.h:38-47 (phase_a_px)
                    const bool m = (HAS_MASK ? ((mbits >> (8 * c)) & 0xffull) != 0ull : (vg[c] > eps)) && act;
                    const float d2 = lg2_approx(clamp_nan(vp[c], eps, 1000.0f)) - lg2_approx(clamp_nan(vg[c], eps, 1000.0f));
                    const float d = m ? d2 : 0.f;
                    si_n += m ? 1u : 0u;
                    si_s += d;
                    si_q = fmaf(d, d, si_q);
                }
            }
        }
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
            if (act) {
                *reinterpret_cast<float4*>(py.lp[0] + o) = make_float4(a0.x, a0.y, a1.x, a1.y);
                *reinterpret_cast<float4*>(py.lg[0] + o) = make_float4(b0.x, b0.y, b1.x, b1.y);
                *reinterpret_cast<float4*>(py.rq[0] + o) = make_float4(rq[0], rq[1], rq[2], rq[3]);
            }
        }
        if ((r & 3) == 3) {   // a row of 2 scale-2 cells
            const float q0 = s2p[0] * 0.0625f, q1 = s2p[1] * 0.0625f, g0 = s2g[0] * 0.0625f, g1 = s2g[1] * 0.0625f;
            const float2 a = log_exact2(make_float2(clamp_nan(q0, eps, 1000.0f), clamp_nan(q1, eps, 1000.0f)));
            const float2 c = log_exact2(make_float2(clamp_nan(g0, eps, 1000.0f), clamp_nan(g1, eps, 1000.0f)));
            const size_t o = ((size_t)b * H2 + (by * 2 + (r >> 2))) * W2 + bx * 2;
            if (act) {
                *reinterpret_cast<float2*>(py.lp[1] + o) = a;
                *reinterpret_cast<float2*>(py.lg[1] + o) = c;
                *reinterpret_cast<float2*>(py.rq[1] + o) = make_float2(in_range_pos(q0, eps, 1000.0f) ? rcp_approx(q0) : 0.f,
                                                                       in_range_pos(q1, eps, 1000.0f) ? rcp_approx(q1) : 0.f);
            }
        }
    }
    {
        const float q = s3p * 0.015625f, g = s3g * 0.015625f;
        const float2 l = log_exact2(make_float2(clamp_nan(q, eps, 1000.0f), clamp_nan(g, eps, 1000.0f)));
        const size_t o = ((size_t)b * H3 + by) * W3 + bx;
        if (act) {
            py.lp[2][o] = l.x;
            py.lg[2][o] = l.y;
            py.rq[2][o] = in_range_pos(q, eps, 1000.0f) ? rcp_approx(q) : 0.f;
        }
    }
    if constexpr ((SF & PS_PSUM) != 0) {
        // this block's sum(pred) -> its image's fixed-point words: one pair of atomics per warp when the 32 blocks lie in
        // one image (almost always), else per lane.  Integer atomics: any order, same bits.
        const int b0 = __shfl_sync(0xffffffffu, b, 0);
        float v = act ? bsum : 0.f;
        const bool together = __all_sync(0xffffffffu, b == b0);
        if (together) v = warp_sum(v);
        if (together ? lane == 0 : act) {
            unsigned long long hi, lo;
            unsigned fl = 0u;
            fix_split_signed((double)v, hi, lo, fl, 0);
            unsigned long long* w = ps.img_words + 4 * (size_t)b;
            if (hi) atomicAdd(w, hi);
            if (lo) atomicAdd(w + 1, lo);
            if (fl) atomicOr(reinterpret_cast<unsigned*>(w + 2), fl);
        }
    }
    }   // blocks
    pdl_wait();
    // the streaming kernel's per-image records (their offset depends on the shape, and one workspace serves calls of
    // different shapes: what lies there may be another shape's partial sums) and pyr_coef_kernel's totals behind them
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < (B + 1) * 32; i += blockDim.x) img_rec_words[i] = 0u;      // 128 bytes each
    if constexpr (SF != 0) {
        __shared__ float s_w[8][2];
        __shared__ unsigned s_n[8];
        __shared__ int s_last;
        const int warp = threadIdx.x >> 5;
        if constexpr ((SF & PS_SI) != 0) {
            const unsigned n = warp_sum(si_n);
            const float s1 = warp_sum(si_s), q1 = warp_sum(si_q);
            if (lane == 0) { s_n[warp] = n; s_w[warp][0] = s1; s_w[warp][1] = q1; }
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned long long n8 = 0ull;
                for (int w = 0; w < 8; ++w) n8 += s_n[w];
                if (n8) atomicAdd(&ps.rec->n, n8);
            } else if (threadIdx.x == 32 || threadIdx.x == 64) {
                const int q = threadIdx.x == 32 ? 0 : 1;
                double t = 0.0;
                for (int w = 0; w < 8; ++w) t += (double)s_w[w][q];
                unsigned long long hi, lo;
                unsigned fl = 0u;
                fix_split_signed(t, hi, lo, fl, q);
                if (hi) atomicAdd(q ? &ps.rec->q_hi : &ps.rec->s_hi, hi);
                if (lo) atomicAdd(q ? &ps.rec->q_lo : &ps.rec->s_lo, lo);
                if (fl) atomicOr(&ps.rec->flags, fl);
            }
        }
        // ticket: the last CTA turns the fixed-point words into the doubles the gradient pass reads, and leaves them zero
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = atomicAdd(&ps.rec->ticket, 1u) == gridDim.x - 1u;
        __syncthreads();
        if (s_last) {
            __threadfence();
            if constexpr ((SF & PS_SI) != 0) {
                if (threadIdx.x == 0) {
                    const unsigned fl = __ldcg(&ps.rec->flags);
                    const double n = (double)__ldcg(&ps.rec->n);
                    const double S = fix_join_signed(__ldcg(&ps.rec->s_hi), __ldcg(&ps.rec->s_lo), fl, 0);
                    const double Q = fix_join_signed(__ldcg(&ps.rec->q_hi), __ldcg(&ps.rec->q_lo), fl, 1);
                    ps.stats[ST_SI_N] = n;
                    ps.stats[ST_SI_S] = S * 0.69314718055994531;                          // log2 -> natural units
                    ps.stats[ST_SI_Q] = Q * (0.69314718055994531 * 0.69314718055994531);
                    if (ps.want_rp) ps.stats[ST_RP_N] = n;
                    ps.rec->n = 0ull; ps.rec->s_hi = 0ull; ps.rec->s_lo = 0ull; ps.rec->q_hi = 0ull; ps.rec->q_lo = 0ull;
                    ps.rec->flags = 0u;
                }
            }
            if constexpr ((SF & PS_PSUM) != 0) {
                for (int i = threadIdx.x; i < B; i += blockDim.x) {
                    unsigned long long* w = ps.img_words + 4 * (size_t)i;
                    const unsigned fl = __ldcg(reinterpret_cast<const unsigned*>(w + 2));
                    ps.img_psum[i] = fix_join_signed(__ldcg(w), __ldcg(w + 1), fl, 0);
                    w[0] = 0ull; w[1] = 0ull; w[2] = 0ull;
                }
            }
            if (threadIdx.x == 0) ps.rec->ticket = 0u;
        }
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
    unsigned long long* rec;   // fixed-point totals hi[6], lo[6], flags (zeroed by pyr_pool_kernel): GX1, GY1, GX2, GY2, GX3, GY3
};

// Latency-bound gathers: fewer, fatter threads win -- at 124 registers (2 CTAs per SM) the compiler keeps the loads of a
// whole level in flight; the step without metrics 108.1 -> 106.0 us against the 64-register build (3 CTAs: 106.7),
// with phase A beside it unchanged.
#ifndef CADL_COEF_MINB
#define CADL_COEF_MINB 2
#endif
__global__ void __launch_bounds__(256, CADL_COEF_MINB) pyr_coef_kernel(const PyrCoefArgs a) {
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
    if (threadIdx.x < 6) {
        // this CTA's sums -> the totals, as fixed-point integer atomics: any order gives the same bits, and the
        // gradient pass reads one record instead of folding a row per CTA
        const int q = threadIdx.x;
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += (double)s_f[w][q];
        unsigned long long hi, lo;
        unsigned fl = 0u;
        fix_split(t, hi, lo, fl, q);
        if (hi) atomicAdd(a.rec + q, hi);
        if (lo) atomicAdd(a.rec + 6 + q, lo);
        if (fl) atomicOr(reinterpret_cast<unsigned*>(a.rec + 12), fl);
    }
}

}  // namespace cadl
