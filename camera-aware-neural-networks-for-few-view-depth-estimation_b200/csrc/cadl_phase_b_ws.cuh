// cadl phase B, warp-specialised persistent form of the aligned fast path (sm_100a).
//
// One CTA of 512 threads per SM, two thread groups, two shared-memory slots (2 x 100 KB):
//   group P (warps 0-7)  : TMA-stages tile k+1 and runs its prelude (ring patch, avg-pool pyramid, pooled logs,
//                          coarse coefficients, logs in place) ...
//   group M (warps 8-15) : ... while it runs the full-resolution pass of tile k.
// The groups meet only through mbarriers (full[slot]: P -> M, empty[slot]: M -> P); inside a group the phase
// barriers are named barriers over 256 threads, so one group never waits for the other's phases.  This removes
// the CTA-wide barrier stalls of the plain fast kernel and keeps both an LDS/ALU-heavy and an FP/LDG-heavy
// instruction stream resident on every scheduler.  Same device functions, same values as phase_b_fast_kernel.
#pragma once
#include "cadl_phase_b_fast.cuh"

namespace cadl {

constexpr int kWsThreads = 2 * kThreadsB;
constexpr size_t kWsSmemBytes = 2 * kFastSmemBytes;

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// Reduce the quantities [q0, q1) of one group's per-thread partial sums and write them to the tile's row.
template <int BAR>
__device__ __forceinline__ void group_publish(const PhaseBArgs& a, float (&acc)[BF_COUNT], int q0, int q1, int row,
                                              float (*s_f)[BF_COUNT], int gtid) {
    const int gw = gtid >> 5, lane = gtid & 31;
    for (int q = q0; q < q1; ++q) {
        const float v = warp_sum(acc[q]);
        if (lane == 0) s_f[gw][q] = v;
        acc[q] = 0.f;
    }
    gsync<BAR>();
    if (gtid >= q0 && gtid < q1) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kThreadsB / 32; ++w) t += (double)s_f[w][gtid];
        a.b_part[(size_t)row * BF_COUNT + gtid] = t;
    }
    gsync<BAR>();
}

template <int F, bool HAS_MASK>
__global__ void __launch_bounds__(kWsThreads, 1) phase_b_ws_kernel(const PhaseBArgs a,
                                                                   const __grid_constant__ CUtensorMap tm_pred,
                                                                   const __grid_constant__ CUtensorMap tm_gt) {
    static_assert((F & FB_GRAD) != 0, "the warp-specialised kernel exists for the variants with a prelude");
    extern __shared__ __align__(128) float smem_raw[];
    __shared__ __align__(8) unsigned long long bar_tma[2], bar_full[2], bar_empty[2];
    __shared__ float s_fP[kThreadsB / 32][BF_COUNT], s_fM[kThreadsB / 32][BF_COUNT];
    __shared__ double s_d[kWsThreads / 32];
    __shared__ int s_last;
    constexpr bool SMOOTH = (F & FB_SMOOTH) != 0;

    const int tid = threadIdx.x;
    const bool is_p = tid < kThreadsB;
    const int gtid = tid & (kThreadsB - 1), gw = gtid >> 5, lane = gtid & 31;
    const int H = a.H, W = a.W;
    const int tiles_per_img = a.tiles_x * a.tiles_y;
    const int num_tiles = tiles_per_img * a.B;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bar_tma[s], 1);
            mbar_init(&bar_full[s], kThreadsB);
            mbar_init(&bar_empty[s], kThreadsB);
        }
    }
    __syncthreads();

    float acc[BF_COUNT];
#pragma unroll
    for (int q = 0; q < BF_COUNT; ++q) acc[q] = 0.f;

    if (is_p) {
        // ================= group P: staging + prelude of the tiles, one slot ahead of group M =================
        int k = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++k) {
            const int s = k & 1, ph = (k >> 1) & 1;
            const int tx = tile % a.tiles_x, ty = (tile / a.tiles_x) % a.tiles_y, b = tile / tiles_per_img;
            const int x0 = tx * FTW, y0 = ty * FTH;
            FastSmem sm;
            sm.sp = smem_raw + (size_t)s * kFastSmemFloats;
            sm.sg = sm.sp + FRH * FRW;
            sm.pl = sm.sg + FRH * FRW;
            sm.pg = sm.pl + kFPoolCells;
            sm.cc = sm.pg + kFPoolCells;
            if (k >= 2) mbar_wait(&bar_empty[s], ph ^ 1);           // group M has finished with this slot
            if (gtid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // our earlier generic writes to the slot
                mbar_expect_tx(&bar_tma[s], 2u * FRH * FRW * sizeof(float));
                tma_load_3d(sm.sp, &tm_pred, x0 - HALO, y0 - HALO, b, &bar_tma[s]);
                tma_load_3d(sm.sg, &tm_gt, x0 - HALO, y0 - HALO, b, &bar_tma[s]);
            }
            mbar_wait(&bar_tma[s], ph);
            fast_prelude<1>(a, sm, gtid, y0, x0, true, acc);
            group_publish<1>(a, acc, BF_GX1, BF_GY3 + 1, tile, s_fP, gtid);   // coarse-scale loss sums of this tile
            mbar_arrive(&bar_full[s]);                               // release: slot s is ready for the full-res pass
        }
    } else {
        // ================= group M: full-resolution pass =================
        // batch-global scalars once; the smoothness normaliser a_b changes with the image
        float sc[4];
        {
            const double n = a.stats[ST_SI_N], S = a.stats[ST_SI_S], nr = a.stats[ST_RP_N];
            const float up = a.upstream;
            sc[0] = n > 0.0 ? (float)(2.0 / n) * a.w_si * up : 0.f;
            sc[1] = n > 0.0 ? (float)(-2.0 * (double)a.lambda * S / (n * n)) * a.w_si * up : 0.f;
            sc[2] = nr > 0.0 ? (float)(1.0 / nr) * a.w_rp * up : 0.f;
            sc[3] = 0.f;
        }
        int k = 0, b_cur = -1;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++k) {
            const int s = k & 1, ph = (k >> 1) & 1;
            const int tx = tile % a.tiles_x, ty = (tile / a.tiles_x) % a.tiles_y, b = tile / tiles_per_img;
            const int x0 = tx * FTW, y0 = ty * FTH;
            FastSmem sm;
            sm.sp = smem_raw + (size_t)s * kFastSmemFloats;
            sm.sg = sm.sp + FRH * FRW;
            sm.pl = sm.sg + FRH * FRW;
            sm.pg = sm.pl + kFPoolCells;
            sm.cc = sm.pg + kFPoolCells;
            if (SMOOTH && b != b_cur) {
                sc[3] = (1.0f / ((float)(a.img_psum[b] / ((double)H * W)) + a.eps_smooth)) * a.w_smooth * a.upstream;   // :192-193
                b_cur = b;
            }
            mbar_wait(&bar_full[s], ph);                             // acquire: prelude of this tile is done
            fast_p4<F, HAS_MASK>(a, sm, gw, lane, b, y0, x0, sc, acc);
            mbar_arrive(&bar_empty[s]);                              // the slot may be overwritten
            // full-resolution loss sums of this tile: GX0, GY0 | SMX, SMY, RP_E
            group_publish<2>(a, acc, BF_GX0, BF_GY0 + 1, tile, s_fM, gtid);
            group_publish<2>(a, acc, BF_SMX, BF_RP_E + 1, tile, s_fM, gtid);
        }
    }

    // ---- every CTA has written the rows of its tiles: the last CTA to arrive reduces them ----
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

}  // namespace cadl
