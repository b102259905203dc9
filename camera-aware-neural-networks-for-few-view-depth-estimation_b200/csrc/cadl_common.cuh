// cadl -- shared device helpers (sm_100a).  Internal to csrc/; the public surface is include/cadl.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "cadl.h"

namespace cadl {

// ------------------------------------------------------------------------------------------------
// Exchangeable statistics vector written by phase A (see include/cadl.h: cadl_stack_reduce).
// All entries are doubles so one all-reduce(sum) moves them; counts are exact below 2^53.
// ------------------------------------------------------------------------------------------------
enum Stat : int {
    ST_SI_N = 0,   // #valid                       depth_loss.h:52
    ST_SI_S = 1,   // sum d                        depth_loss.h:61
    ST_SI_Q = 2,   // sum d^2                      depth_loss.h:58
    ST_RP_N = 3,   // #valid                       depth_loss.h:323
    ST_EV_N = 8,   // depth_metrics.h:60
    ST_EV_ABSREL = 9, ST_EV_SQREL = 10, ST_EV_SQ = 11, ST_EV_LOGSQ = 12, ST_EV_ABS = 13,
    ST_EV_LOG10 = 14, ST_EV_C1 = 15, ST_EV_C2 = 16, ST_EV_C3 = 17, ST_EV_SUMP = 18, ST_EV_SUMG = 19,
    ST_TR_N = 20,  // tensorboard_trainer_enhanced.h:414
    ST_TR_ABSREL = 21, ST_TR_SQREL = 22, ST_TR_SQ = 23, ST_TR_LOGSQ = 24,
    ST_TR_C1 = 25, ST_TR_C2 = 26, ST_TR_C3 = 27,
    ST_COUNT = 32
};

// Phase-A per-thread float accumulators (order of the per-block partial rows).
enum A_Float : int {
    AF_SI_S = 0, AF_SI_Q, AF_PSUM,
    AF_EV_ABSREL, AF_EV_SQREL, AF_EV_SQ, AF_EV_LOGSQ, AF_EV_ABS, AF_EV_LOG10, AF_EV_SUMP, AF_EV_SUMG,
    AF_TR_ABSREL, AF_TR_SQREL, AF_TR_SQ, AF_TR_LOGSQ,
    AF_COUNT
};
// Phase-A integer counters (exact; accumulated with integer atomics: order-independent).
enum A_Int : int {
    AI_SI_N = 0, AI_RP_N, AI_EV_N, AI_EV_C1, AI_EV_C2, AI_EV_C3, AI_TR_N, AI_TR_C1, AI_TR_C2, AI_TR_C3,
    AI_COUNT
};
// Phase-B per-tile partial sums.
enum B_Float : int {
    BF_GX0 = 0, BF_GY0, BF_GX1, BF_GY1, BF_GX2, BF_GY2, BF_GX3, BF_GY3,  // sum|e_x|, sum|e_y| per scale
    BF_SMX, BF_SMY,                                                     // sum w_x|dx p|, sum w_y|dy p|
    BF_RP_E,                                                            // sum_m e  (reprojection)
    BF_COUNT
};

struct WsHeader {
    unsigned int ticket_a;
    unsigned int ticket_b;
    unsigned int pad[2];
    unsigned long long icount[16];
    unsigned long long pool_rec[8];     // PoolStatsRec (cadl_phase_b_stream.cuh): fixed offset, zero between calls
};

struct WsLayout {
    size_t header, stats, img_words, img_psum, a_part, b_part, img_sm, img_off, img_rec, total;
    size_t pyr_lp[3], pyr_lg[3], pyr_rq[3], pyr_c1;   // streaming fast path (cadl_phase_b_stream.cuh); 0 = absent
    int a_blocks_per_img, a_blocks, b_tiles;
    int pyr_blocks;   // CTAs of pyr_pool_kernel / pyr_coef_kernel (0: shape not a multiple of 8)
};

// Capacities of the partial-row arrays and caps of the grid-stride grids.  They are LAYOUT constants (the workspace size
// must not depend on the device): sized for B200's 148 SMs x 8 CTAs; launch geometry that has to match the device
// (resident waves, cooperative grids) queries the SM count and the occupancy instead.
constexpr int kLayoutSms = 148;
constexpr int kGridCap = kLayoutSms * 8;
constexpr int kPointBlocks = kGridCap;      // partial rows of the pointwise kernels

constexpr int kThreadsA = 256;
constexpr int kThreadsB = 256;
constexpr int TH = 32;    // tile rows
constexpr int TW = 128;   // tile cols (one warp-row of float4)
constexpr int HALO = 8;   // 2^(CADL_MAX_SCALES-1)

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ inline int a_blocks_per_image(int B, int HW) {
    // ~4 resident blocks per SM over 148 SMs in one wave (fat threads amortise the block-end reduction),
    // at least 1024 px per block, each block inside one image
    int target = (kLayoutSms * 4) / B;          // rounded DOWN: one block too many per image starts a second, almost empty wave
    if (target < 1) target = 1;
    int by_size = (HW + 1023) / 1024;
    int n = target < by_size ? target : by_size;
    return n < 1 ? 1 : n;
}

__host__ inline WsLayout ws_layout(int B, int H, int W) {
    WsLayout L;
    int HW = H * W;
    L.a_blocks_per_img = a_blocks_per_image(B, HW);
    L.a_blocks = L.a_blocks_per_img * B;
    int tx = (W + TW - 1) / TW, ty = (H + TH - 1) / TH;   // generic tiling (32 x 128) >= fast tiling (48 x 128)
    L.b_tiles = tx * ty * B;
    size_t o = 0;
    L.header = o;   o = align_up(o + sizeof(WsHeader), 256);
    L.stats = o;    o = align_up(o + sizeof(double) * ST_COUNT, 256);
    L.img_words = o; o = align_up(o + 32 * (size_t)B, 256);                    // pyr_pool_kernel's per-image fixed-point words (fixed offset; zero between calls)
    L.img_psum = o; o = align_up(o + sizeof(double) * B, 256);
    L.a_part = o;   o = align_up(o + sizeof(double) * (size_t)L.a_blocks * AF_COUNT, 256);
    const bool pyr = (H % 8 == 0) && (W % 8 == 0);
    L.pyr_blocks = pyr ? (int)(((size_t)B * (H / 8) * (W / 8) + 255) / 256) : 0;
    size_t b_rows = (size_t)L.b_tiles;
    if (b_rows < (size_t)kPointBlocks) b_rows = kPointBlocks;
    // streaming path: [B per-image rows (unused since round 2)][pyr_coef_kernel rows]
    if (pyr && b_rows < (size_t)B + L.pyr_blocks) b_rows = (size_t)B + L.pyr_blocks;
    L.b_part = o;   o = align_up(o + sizeof(double) * b_rows * BF_COUNT, 256);
    L.img_sm = o;   o = align_up(o + sizeof(double) * (size_t)B * 2, 256);
    L.img_off = o;  o = align_up(o + sizeof(float) * (size_t)B, 256);
    L.img_rec = o;  o = align_up(o + 128 * ((size_t)B + 1), 256);              // ImgRec per image (cadl_stream3_host.h) + pyr_coef_kernel's totals
    for (int s = 0; s < 3; ++s) {
        const size_t cells = pyr ? (size_t)B * (H >> (s + 1)) * (W >> (s + 1)) : 0;
        L.pyr_lp[s] = pyr ? o : 0; o = align_up(o + sizeof(float) * cells, 256);
        L.pyr_lg[s] = pyr ? o : 0; o = align_up(o + sizeof(float) * cells, 256);
        L.pyr_rq[s] = pyr ? o : 0; o = align_up(o + sizeof(float) * cells, 256);
    }
    L.pyr_c1 = pyr ? o : 0;
    o = align_up(o + (pyr ? sizeof(float) * (size_t)B * (H / 2) * (W / 2) : 0), 256);
    L.total = o;
    return L;
}

// torch::clamp semantics: NaN propagates (fminf/fmaxf would drop it).
__device__ __forceinline__ float clampf(float x, float lo, float hi) {
    return x < lo ? lo : (x > hi ? hi : x);
}
// at::sgn for real floats: 0 at 0 (abs backward)
__device__ __forceinline__ float sgnf(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned warp_sum(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 128-bit load (read-once data: rgb): do not allocate in L1
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// Metrics results from the phase-A statistics (depth_metrics.h:69-85; trainer :418-436).  Called by at least 32
// threads with t = thread index: one value per thread (a single thread doing the ~25 double divisions one after
// the other was ~4 us of serial tail).
__device__ inline void write_metric_results(const double* st, uint32_t which, cadl_results& r, int t) {
    if (which & CADL_METRICS_EVAL) {
        const double n = st[ST_EV_N];
        if (t < 12) {
            // getZeroMetrics when nothing is valid, depth_metrics.h:238-253
            const int src = t == 0 ? ST_EV_ABSREL : t == 1 ? ST_EV_SQREL : t == 2 ? ST_EV_SQ : t == 3 ? ST_EV_LOGSQ
                          : t == 4 ? ST_EV_ABS : t == 5 ? ST_EV_LOG10 : t == 6 ? ST_EV_C1 : t == 7 ? ST_EV_C2
                          : t == 8 ? ST_EV_C3 : t == 10 ? ST_EV_SUMP : ST_EV_SUMG;
            float v = 0.f;
            if (n > 0.0) {
                if (t == 9) v = (float)n;                                  // static_cast<float>(num_valid), :83
                else if (t == 2 || t == 3) v = sqrtf((float)(st[src] / n));
                else v = (float)(st[src] / n);
            }
            r.eval[t] = v;
        } else if (t < 16) {
            const int k = t - 12;
            r.eval_counts[k] = (int64_t)st[k == 0 ? ST_EV_N : ST_EV_C1 + (k - 1)];
        }
    }
    if (which & CADL_METRICS_TRAIN) {
        const double n = st[ST_TR_N];
        if (t >= 16 && t < 24) {
            const int k = t - 16;
            const int src = k == 0 ? ST_TR_ABSREL : k == 1 ? ST_TR_SQREL : k == 2 ? ST_TR_SQ : k == 3 ? ST_TR_LOGSQ
                          : k == 4 ? ST_TR_C1 : k == 5 ? ST_TR_C2 : ST_TR_C3;
            float v = 0.f;
            if (n > 0.0 && k < 7) v = (k == 2 || k == 3) ? sqrtf((float)(st[src] / n)) : (float)(st[src] / n);
            r.train[k] = v;
        } else if (t >= 24 && t < 28) {
            const int k = t - 24;
            r.train_counts[k] = (int64_t)st[k == 0 ? ST_TR_N : ST_TR_C1 + (k - 1)];
        }
    }
}

}  // namespace cadl
