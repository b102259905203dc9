// cadl gradient pass, round-2 form (aligned shapes; replaces phase_b_stream_kernel + stream_finish_kernel of round 1).
//
// One warp owns 128 columns and marches down its rows; every stencil edge is evaluated once; no CTA-wide barrier.
// What changed against round 1, each change answering a line of the round-1 ncu profile
// (profiles/r01_v4_streaming.txt: 194 lane-instructions per pixel, 61 % issue-active, 12 warps per SM, 20 us offset pass):
//
//  1. TMA ROW RING.  Rows arrive through cp.async.bulk.tensor (one elected lane, one 136-pixel box per tensor and
//     row: the warp's 128 columns plus a 4-pixel halo either side, zero-filled outside the image) into a private
//     shared-memory ring of kS3Depth rows per warp, completion on one mbarrier per slot.  Bytes in flight no longer
//     cost registers (the one-row register prefetch of round 1 capped the kernel at ~2 TB/s once the arithmetic got
//     short), the halo pixel is an ordinary LDS, and no lane does global address arithmetic for the five streams.
//  2. TWO-TIER LOGS.  The gradient-matching term needs sign((lp1 - lp0) - (lg1 - lg0)) bit-identical to the
//     reference (depth_loss.h:140-163), which in round 1 meant a logf replica for every pixel.  Here every pixel gets
//     lg2.approx (one MUFU) and the residual is formed from d = lg2 p - lg2 g; its sign is the reference's sign
//     whenever |e| exceeds kBand, which bounds the total error of both computations (cadl_selftest(2) measures the
//     lg2.approx error).  Only a warp-row that holds an edge inside the band (about 1 % of them on BASELINE's data)
//     recomputes its edges with log_exact in the reference's operation order (exact_tier, out of line).  The loss
//     sums and the SI term only have to meet 1e-5 and use d directly.
//  3. PACKED REPROJECTION.  The reference-order back-projection (four correctly rounded quotients per pixel) runs on
//     fp32x2 FMAs: the lane's (u - cx) pairs are loop constants, so the pairs line up with the float4 reads.
//  4. NO OFFSET KERNEL.  The smoothness gradient is a_b * G_j - a_b * L_b / (HW) and L_b needs the whole image: each
//     warp subtracts the offset from the rows IT wrote (L2-resident) once the image's sums are complete.  Per-image
//     sums travel through fixed-point integer atomics (order-independent, hence deterministic).
//
// Cooperative launch: warps wait for other warps' image sums, so the grid must be co-resident.
#pragma once
#include <cuda.h>   // CUtensorMap (type only)
#include <type_traits>
#include "cadl_args.cuh"
#include "cadl_stream3_host.h"

namespace cadl {

// Edge residuals (log2 units) below this magnitude take the exact tier.  Bound of |e_approx * ln2 - e_reference|:
// approximate side: four lg2.approx results (absolute error < 2.3e-6 incl. the rounding of a result of magnitude < 20:
// measured by cadl_selftest(2), tests/test_math_gpu.py, profiles/lg2_probe.py) and three fp32 subtractions of
// magnitude < 32 (3 x 2^-20): 1.2e-5 in log2 units = 8.3e-6; reference side: four logf (1 ulp of <= 13.9: 2^-20 each)
// and three subtractions: 6.7e-6.  Together 1.5e-5 = 2.16e-5 in log2 units; 2^-15 = 3.05e-5 leaves 29 % margin (and the
// worst case needs all four values near the lower clamp, where equal inputs give equal logs and cancel).
constexpr float kBand = 3.0517578125e-05f;
constexpr float kLn2 = 0.693147180559945309f;

// |mag| with the sign of s (mag >= 0): one LOP3
__device__ __forceinline__ float with_sign(float mag, float s) {
    return __int_as_float((__float_as_int(s) & (int)0x80000000) | __float_as_int(mag));
}
// same, and exactly 0 where s == 0 (sign(0) = 0: at::sgn / abs backward).  A NaN s is flagged by the caller.
__device__ __forceinline__ float with_sign0(float mag, float s) { return with_sign(s != 0.f ? mag : 0.f, s); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
// sign with NaN passed through (the exact tier)
__device__ __forceinline__ float sgn3n(float x) { return (x != x) ? x : sgn3(x); }

// one lane of a converged warp (the compiler then knows the guarded block runs on exactly one thread: TMA operands
// go straight to uniform registers instead of a per-active-lane loop)
__device__ __forceinline__ bool elect_one() {
    unsigned p;
    asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\nselp.u32 %0, 1, 0, q;\n}" : "=r"(p));
    return p != 0;
}
// mbarrier wait with a unique label per expansion
// Bounded: a wait that outlives kS3SpinLimit polls (seconds) is a bug or a dead peer; trap instead of hanging the GPU.
constexpr unsigned long long kS3WaitLimitNs = 2000000000ull;
__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// what a wait that ran out of time leaves behind (cadl_results::_pad0 as three ints: code, warp, detail) before the
// thread exits; the host sees loss values that make no sense plus this record
__device__ __noinline__ void s3_bail(cadl_results* r, int code, int warp, int detail) {
    int* dbg = reinterpret_cast<int*>(r->_pad0);
    dbg[0] = code; dbg[1] = warp; dbg[2] = detail;
    __threadfence();
    asm volatile("exit;");
}
__device__ __forceinline__ void mbar_wait_parity(unsigned long long* bar, unsigned parity, cadl_results* r, int warp, int detail) {
    const unsigned addr = smem_u32(bar);
    unsigned done = 0, spins = 0;
    unsigned long long t0 = 0;
    do {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (!done && (++spins & 15u) == 0u) {
            const unsigned long long t = gtime_ns();
            if (t0 == 0) t0 = t;
            else if (t - t0 > kS3WaitLimitNs) s3_bail(r, 1, warp, detail);
        }
    } while (!done);
}

// one probe, no loop: lets the step ask for the NEXT row's slot at its top (the answer travels through the MIO queue like
// a shared-memory load) and only branch on it where the row is needed
__device__ __forceinline__ unsigned mbar_test_parity(unsigned long long* bar, unsigned parity) {
    unsigned done;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done;
}

// ---- the exact tier, out of line (rare: keeps the row loop short and its register budget for the common path) ----
// Signs (-1, 0, +1, NaN passed through) of the four x-edges to the right of the lane's pixels [1..4], of the edge to
// its left [0], and of the four y-edges C -> N, evaluated with the reference's operations (log(clamp(.)), differences
// in the reference's order: depth_loss.h:115-116, 140-163); with SMOOTH also the signs of the depth steps themselves.
struct ExactSigns {
    float sx[5], sy[4], tx[5], ty[4];
};
// cq / nq: this lane's quad of the current / next row inside their ring slots (pred at [0..3], the neighbours at [-1]
// and [4], gt one row pitch further).  The rows are read again from shared memory: passing the ten registers by value
// put a dozen argument moves in front of the branch of EVERY row.
template <bool SMOOTH>
__device__ __noinline__ ExactSigns exact_tier(const float* cq, const float* nq, float eps) {
    ExactSigns o;
    const float4 cp4 = *reinterpret_cast<const float4*>(cq), cg4 = *reinterpret_cast<const float4*>(cq + kS3RowFloats);
    const float4 np4 = *reinterpret_cast<const float4*>(nq), ng4 = *reinterpret_cast<const float4*>(nq + kS3RowFloats);
    const float pl = cq[-1], gl = cq[kS3RowFloats - 1], pr = cq[4], gr = cq[kS3RowFloats + 4];
    const float cp[4] = {cp4.x, cp4.y, cp4.z, cp4.w}, cg[4] = {cg4.x, cg4.y, cg4.z, cg4.w};
    const float np[4] = {np4.x, np4.y, np4.z, np4.w}, ng[4] = {ng4.x, ng4.y, ng4.z, ng4.w};
    float lp[5], lg[5];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 l = log_exact2(make_float2(clamp_nan(cp[k], eps, 1000.0f), clamp_nan(cg[k], eps, 1000.0f)));
        lp[k] = l.x; lg[k] = l.y;
        const float2 ln = log_exact2(make_float2(clamp_nan(np[k], eps, 1000.0f), clamp_nan(ng[k], eps, 1000.0f)));
        o.sy[k] = sgn3n((ln.x - l.x) - (ln.y - l.y));
        o.ty[k] = SMOOTH ? sgn3n(np[k] - cp[k]) : 0.f;
    }
    const float2 ll = log_exact2(make_float2(clamp_nan(pl, eps, 1000.0f), clamp_nan(gl, eps, 1000.0f)));
    const float2 lr = log_exact2(make_float2(clamp_nan(pr, eps, 1000.0f), clamp_nan(gr, eps, 1000.0f)));
    lp[4] = lr.x; lg[4] = lr.y;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        o.sx[j + 1] = sgn3n((lp[j + 1] - lp[j]) - (lg[j + 1] - lg[j]));
        o.tx[j + 1] = SMOOTH ? sgn3n((j == 3 ? pr : cp[j + 1]) - cp[j]) : 0.f;
    }
    o.sx[0] = sgn3n((lp[0] - ll.x) - (lg[0] - ll.y));
    o.tx[0] = SMOOTH ? sgn3n(cp[0] - pl) : 0.f;
    return o;
}

#ifdef CADL_S3_TRACE
// tuning builds only (profiles/r02_trace.py): per warp {start, main loop done, image ready seen, end} in globaltimer ns,
// rows whose slot was not ready when probed, and the time spent blocked on them
constexpr int kS3TraceWords = 8;
__device__ unsigned long long g_s3_trace[kS3TraceWords * 4096];
#endif

// Share `item` of the one-wave partition: image, 128-column strip, rows [ys, ye).  Items are numbered row range first
// (item = kk * columns + column): the four warps of a CTA work on neighbouring strips of the same rows.
// (Measured, not adopted: row ranges cut longer for the CTAs dispatched first to an SM, whose warps the schedulers
//  favour -- 44 / 46 / 51 us mean loop time by dispatch rank with equal ranges, 46 / 47 / 49 with ranges 1.07 : 1.01 :
//  0.92, but the slowest warp of every image still ends at 53-55 us and the kernel time does not move.)
struct Share3 { int b, strip, ys, ye; };
__device__ __forceinline__ int s3_row_bound(int k, int H, const Stream3Args& sa) {
    return (int)((long long)H * k / sa.kpi);
}
__device__ __forceinline__ Share3 s3_share(int item, int B, int H, const Stream3Args& sa) {
    const int ncols = B * sa.nstrip;
    const int kk = item / ncols, col = item - kk * ncols;
    Share3 s;
    s.b = col / sa.nstrip; s.strip = col - s.b * sa.nstrip;
    s.ys = s3_row_bound(kk, H, sa); s.ye = s3_row_bound(kk + 1, H, sa);
    return s;
}

// smooth_image_share (cadl_args.cuh) with the divisions by shape constants turned into multiplications: it runs on
// the chain between the last row and the end of the kernel (last warp of an image -> results block), where a double
// division is ~0.15 us of pure latency
__device__ __forceinline__ void smooth_share3(const PhaseBArgs& a, const Stream3Args& sa, int img, double sx, double sy,
                                              double& Lb, float& off) {
    const double mean = a.img_psum[img] * sa.r_hw;
    const float ab = 1.0f / ((float)mean + a.eps_smooth);   // depth_loss.h:193
    Lb = (double)ab * (sx * sa.rnx[0] + sy * sa.rny[0]);
    off = (float)((double)a.upstream * a.w_smooth * ab * Lb * sa.r_hw);
}

// One image row as a lane holds it in registers.
struct Row3 {
    float4 p, g;        // pred, gt (4 adjacent pixels)
    float d[4];         // lg2(clamp(pred)) - lg2(clamp(gt))     depth_loss.h:115-116 (log2 units)
    // (the rgb values are NOT kept: each use reads them from the row's ring slot again -- 24 registers less per thread,
    //  which is what lets a fourth CTA of 128 threads onto the SM)
};

template <int F, bool HAS_MASK>
__global__ void __launch_bounds__(kS3Threads, kS3MinBlocks)
stream3_kernel(const PhaseBArgs a, const Stream3Args sa, const __grid_constant__ CUtensorMap tm_pred,
               const __grid_constant__ CUtensorMap tm_gt, const __grid_constant__ CUtensorMap tm_rgb,
               const __grid_constant__ CUtensorMap tm_c1) {
    constexpr bool SMOOTH = (F & FB_SMOOTH) != 0;
    constexpr bool SI = (F & FB_SI) != 0;
    constexpr bool RP = (F & FB_RP) != 0;
    static_assert((F & FB_GRAD) != 0, "the streaming kernel is the gradient-matching path");
    constexpr int NI = SMOOTH ? 5 : 2;                       // image tensors streamed: pred, gt, 3 x rgb
    // floats per ring slot: pred, gt (pitch kS3RowFloats), the three rgb rows (one box of depth 3: dense, pitch kS3BoxW)
    // and the coarse-scale field C1 (half resolution); every array starts on TMA's 128-byte alignment
    constexpr int RGB0 = 2 * kS3RowFloats, RGBP = kS3RgbPitch;
    constexpr int C1OFF = RGB0 + (SMOOTH ? kS3RgbFloats : 0);
    constexpr int SLOT = C1OFF + kS3C1Floats;
    constexpr int D = kS3Depth;
    extern __shared__ __align__(128) unsigned char s3_smem[];
    __shared__ __align__(8) unsigned long long s_bar[kS3Threads / 32][D];

    const int tid = threadIdx.x, lane = tid & 31;
    const int wib = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform by construction (and known to be)
    const int H = a.H, W = a.W;
    const int plane = H * W;
    const float up = a.upstream;
    const float inx0 = sa.inx0, iny0 = sa.iny0;
    const float eps = a.eps_grad;                            // == eps_si == eps_rp on this path (host-checked)
    constexpr float kExpScale = -1.4426950408889634f / 3.0f; // exp(-mean_c|dI|) = 2^(kExpScale * sum_c|dI|)
    const bool want_grad = a.grad != nullptr;

    // this warp's ring: D slots x NT arrays of kS3RowFloats floats (element j <-> column x0 - 4 + j)
    float* ring = reinterpret_cast<float*>(s3_smem) + (size_t)wib * D * SLOT;
    unsigned long long* bars = s_bar[wib];
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < D; ++s) mbar_init(bars + s, 1);
    }
    __syncwarp();
    unsigned ph = 0u;                                        // phase parity per slot

    pdl_wait();        // C1 of pyr_coef_kernel, the statistics of phase A
#ifdef CADL_S3_TRACE
    unsigned long long tr_t0 = gtime_ns(), tr_main = 0, tr_ready = 0, tr_blocked = 0, tr_late = 0;
    unsigned long long tr_c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    // "image complete" is signalled as ready == epoch: nothing has to be reset while other warps may still poll it
    const unsigned epoch = __ldcg(sa.epoch) + 1u;
    // scalars derived from the phase-A statistics, weights and upstream folded in; d is in log2 units
    float c1n = 0.f, c2 = 0.f, rpn = 0.f;
    {
        const double n = a.stats[ST_SI_N], S = a.stats[ST_SI_S], nr = a.stats[ST_RP_N];
        if (SI && n > 0.0) {
            c1n = (float)(2.0 / n) * a.w_si * up * kLn2;
            c2 = (float)(-2.0 * (double)a.lambda * S / (n * n)) * a.w_si * up;
        }
        if (RP && nr > 0.0) rpn = (float)(1.0 / nr) * a.w_rp * up;
    }

    const int gwarp = blockIdx.x * (kS3Threads / 32) + wib;
    const int spi = sa.nstrip * sa.kpi;                      // shares per image
    const int nitems = a.B * spi;

    for (int item = gwarp; item < nitems; item += sa.nwarps) {
        const Share3 sh = s3_share(item, a.B, H, sa);
        const int b = sh.b, strip = sh.strip, ys = sh.ys, ye = sh.ye;                    // rows [ys, ye)

        const int img = b * plane;                           // B*H*W < 2^31 (checked on the host)
        float ab = 0.f;
        if (SMOOTH) ab = 1.0f / ((float)(a.img_psum[b] * sa.r_hw) + a.eps_smooth);              // a_b (:192-193)
        float fxe = 1.f, fye = 1.f, rfx = 1.f, rfy = 1.f, cxv = 0.f, cyv = 0.f;
        bool mk_ok = true;
        if constexpr (RP) {
            float fx, fy;
            load_K(a, b, fx, fy, cxv, cyv);
            fxe = fx + eps; fye = fy + eps;
            rfx = __frcp_rn(fxe); rfy = __frcp_rn(fye);
            mk_ok = markstein_safe(fxe) && markstein_safe(fye);
        }

        const int x0 = strip * 128;
        const int gx0 = x0 + 4 * lane;
        const bool lane_in = gx0 < W;                        // W % 4 == 0: a lane is fully inside or outside
        const bool border_r = gx0 + 4 >= W;                  // no edge to the right of this lane's last pixel
        const bool left_edge = gx0 >= 4 || (gx0 >= 1);       // an edge to the left of this lane's first pixel
        // at the image borders the neighbour pixel is the TMA zero fill and the edge's magnitude constants are zero
        const float inx3 = border_r ? 0.f : inx0;            // magnitude of this lane's 4th x-edge
        const float inxl = (lane == 0 && gx0 >= 1) ? inx0 : 0.f;    // lane 0's own left edge (other lanes: from the left lane)
        const float lsn3 = border_r ? -INFINITY : sa.lsnx;   // same for the smoothness weight: 2^-inf = 0
        const float lsnl = (lane == 0 && gx0 >= 1) ? sa.lsnx : -INFINITY;
        (void)left_edge;
        const float ufx0 = (float)gx0;                       // u of the lane's first pixel; u + k is exact

        float2 ax01 = make_float2(0.f, 0.f), ax23 = ax01;
        if constexpr (RP) {
            // (u - cx), u = float(column): depth_loss.h:283,299
            ax01 = make_float2(__fadd_rn(ufx0, -cxv), __fadd_rn(__fadd_rn(ufx0, 1.f), -cxv));
            ax23 = make_float2(__fadd_rn(__fadd_rn(ufx0, 2.f), -cxv), __fadd_rn(__fadd_rn(ufx0, 3.f), -cxv));
        }

        // ---- the row sequence of this share: [row above], rows ys .. ye-1, row below (the bottom image row again at
        //      the image border: that edge then has residual exactly 0) ----
        const int r_first = ys > 0 ? ys - 1 : 0;
        const int nseq = (ye - r_first) + 1;
        auto issue = [&](int i, int s) {       // one lane: fetch sequence element i into slot s (= i mod D)
            int y = r_first + i;
            y = y < H ? y : H - 1;
            float* dst = ring + (size_t)s * SLOT;
            mbar_expect_tx(bars + s, NI * kS3BoxBytes + kS3C1BoxBytes);
            tma_load_3d(dst + C1OFF, &tm_c1, (x0 >> 1) - 4, y >> 1, b, bars + s);
            tma_load_3d(dst, &tm_pred, x0 - 4, y, b, bars + s);
            tma_load_3d(dst + kS3RowFloats, &tm_gt, x0 - 4, y, b, bars + s);
            if constexpr (SMOOTH) {
                tma_load_3d(dst + RGB0, &tm_rgb, x0 - 4, y, 3 * b, bars + s);      // the three channel rows in one box
            }
        };
        auto wait_slot = [&](int s) {
            mbar_wait_parity(bars + s, (ph >> s) & 1u, a.results, gwarp, s);
            ph ^= 1u << s;
        };
        const float* myq = ring + 4 + 4 * lane;              // this lane's quad inside a row array
        // own quads of sequence element i -> registers, and its log differences
        auto read_row = [&](int s, Row3& R) {
            const float* q = myq + s * (SLOT);
            R.p = *reinterpret_cast<const float4*>(q);
            R.g = *reinterpret_cast<const float4*>(q + kS3RowFloats);
            const float pv[4] = {R.p.x, R.p.y, R.p.z, R.p.w}, gv[4] = {R.g.x, R.g.y, R.g.z, R.g.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                R.d[k] = lg2_approx(clamp_nan(pv[k], eps, 1000.0f)) - lg2_approx(clamp_nan(gv[k], eps, 1000.0f));
        };

        // the ring holds elements [ic, ic + D - 1] while step(ic) runs; the first step is ic = (ys > 0), and each step
        // issues element ic - 1 + D itself
        if (elect_one()) {
            const int n0 = D - 1 + (ys > 0 ? 1 : 0);
            for (int i = 0; i < n0 && i < nseq; ++i) issue(i, i);
        }

        float sg_gx = 0.f, sg_gx3 = 0.f, sg_gy = 0.f, sg_smx = 0.f;   // this share's sums
        float2 sg_smy = make_float2(0.f, 0.f), sg_rp = make_float2(0.f, 0.f);

        // ---- vertical edges (row C -> row N): approximate tier ----
        // sys: signed magnitude of d loss_0 / d e_y; tys: same for the smoothness term; flag: a residual inside the
        // band or a NaN -> the exact tier decides (a depth step that is exactly zero -- saturated predictions -- is
        // common and handled in line)
        auto yedges = [&](const Row3& C, const Row3& N, const float* cI, const float* nI, float (&sys)[4], float (&tys)[4], bool& flag) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float e = N.d[k] - C.d[k];                              // depth_loss.h:151-163
                sys[k] = with_sign(iny0, e);
                sg_gy += fabsf(e);
                flag |= !(fabsf(e) >= kBand);
            }
            if constexpr (SMOOTH) {
                const float2 lsy = make_float2(sa.lsny, sa.lsny), ks = make_float2(kExpScale, kExpScale);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float2 cp = h ? make_float2(C.p.z, C.p.w) : make_float2(C.p.x, C.p.y);
                    const float2 np = h ? make_float2(N.p.z, N.p.w) : make_float2(N.p.x, N.p.y);
                    float2 s = make_float2(0.f, 0.f);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float2 ci = *reinterpret_cast<const float2*>(cI + RGB0 + c * RGBP + 2 * h);
                        const float2 ni = *reinterpret_cast<const float2*>(nI + RGB0 + c * RGBP + 2 * h);
                        const float2 di = __fadd2_rn(ni, neg2(ci));
                        s = __fadd2_rn(s, make_float2(fabsf(di.x), fabsf(di.y)));     // depth_loss.h:218-227
                    }
                    const float2 arg = __ffma2_rn(s, ks, lsy);
                    const float2 wy = make_float2(ex2_approx(arg.x), ex2_approx(arg.y));   // (w up / N_y) exp(-mean|dI|)
                    const float2 dp = __fadd2_rn(np, neg2(cp));
                    sg_smy = __ffma2_rn(wy, make_float2(fabsf(dp.x), fabsf(dp.y)), sg_smy);
                    tys[2 * h] = with_sign0(wy.x, dp.x);
                    tys[2 * h + 1] = with_sign0(wy.y, dp.y);
                    flag |= (dp.x != dp.x) | (dp.y != dp.y);                  // NaN: the exact tier passes it on
                }
            }
        };

        // One row: C is the current row (sequence element ic, complete in registers), N receives element ic + 1.
        // up: signed terms of the edges to the row above (from the previous step); dn: those to the row below.
        // sc: ring slot of the current row (element ic); the slot before it is free and takes element ic - 1 + D
        // mk: std::true_type when Markstein's division covers both divisors (always, in practice): a compile-time tag, the
        // row loop exists twice
        auto step = [&](auto mk, int ic, int sc, int gy, float gyf, Row3& C, Row3& N, const float (&sy_up)[4], const float (&ty_up)[4],
                        float (&sy_dn)[4], float (&ty_dn)[4]) {
            constexpr bool MK = decltype(mk)::value;
            (void)MK;
            // 1. the slot of element ic - 1 was last read in the previous step: refill it
            __syncwarp();
            const int sp = sc == 0 ? D - 1 : sc - 1, sn = sc == D - 1 ? 0 : sc + 1;
            if (ic - 1 + D < nseq) {
                if (elect_one()) issue(ic - 1 + D, sp);
            }
            // the next row's slot: probe now, look at the answer in 4. (normally it completed several rows ago)
            const unsigned n_ready = mbar_test_parity(bars + sn, (ph >> sn) & 1u);
            uchar4 mk4 = make_uchar4(0, 0, 0, 0);
            const int gxc = lane_in ? gx0 : W - 4;
            (void)gxc;
            if constexpr (HAS_MASK) mk4 = __ldg(reinterpret_cast<const uchar4*>(a.mask + img + gy * W + gxc));
            const float cp[4] = {C.p.x, C.p.y, C.p.z, C.p.w}, cg[4] = {C.g.x, C.g.y, C.g.z, C.g.w};
            bool flag = false;

            // 2. horizontal edges of the current row: each lane evaluates the four edges to the right of its pixels;
            //    the edge to its left comes from the left lane (lane 0: evaluated from the halo pixel)
            const float* cq = myq + sc * (SLOT);                  // the current row's slot: neighbours
            // the coarse-scale field of this row's 2x2 cells: element j of its array <-> cell column x0/2 - 4 + j
            // (the box starts 16-byte aligned like the image boxes: a start at x0/2 - 2 raised an illegal-instruction fault)
            const float2 ccv = *reinterpret_cast<const float2*>(ring + sc * (SLOT) + C1OFF + 4 + 2 * lane);
            const float pr = cq[4], gr = cq[kS3RowFloats + 4];                 // right neighbour of the lane's last pixel
            const float pl = cq[-1], gl = cq[kS3RowFloats - 1];                // left neighbour of its first pixel
            float sx[5], tx[5];
            (void)tx;
            {
                const float dr = lg2_approx(clamp_nan(pr, eps, 1000.0f)) - lg2_approx(clamp_nan(gr, eps, 1000.0f));
                const float dx[5] = {C.d[0], C.d[1], C.d[2], C.d[3], dr};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float e = dx[j + 1] - dx[j];                        // depth_loss.h:140-148,162
                    sx[j + 1] = with_sign(j == 3 ? inx3 : inx0, e);
                    if (j == 3) sg_gx3 += fabsf(e); else sg_gx += fabsf(e);
                    flag |= !(fabsf(e) >= kBand);
                }
                const float dl = lg2_approx(clamp_nan(pl, eps, 1000.0f)) - lg2_approx(clamp_nan(gl, eps, 1000.0f));
                const float el = C.d[0] - dl;                                 // lane 0: edge to the left strip
                sx[0] = with_sign(inxl, el);
                flag |= (lane == 0) & !(fabsf(el) >= kBand);
            }
            if constexpr (SMOOTH) {
                const float px[5] = {cp[0], cp[1], cp[2], cp[3], pr};
                float Ix[3][5], Il[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float4 v = *reinterpret_cast<const float4*>(cq + RGB0 + c * RGBP);
                    Ix[c][0] = v.x; Ix[c][1] = v.y; Ix[c][2] = v.z; Ix[c][3] = v.w;
                    Ix[c][4] = cq[RGB0 + c * RGBP + 4];
                    Il[c] = cq[RGB0 + c * RGBP - 1];
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float s = fabsf(Ix[0][j + 1] - Ix[0][j]) + fabsf(Ix[1][j + 1] - Ix[1][j]) + fabsf(Ix[2][j + 1] - Ix[2][j]);
                    const float wx = ex2_approx(fmaf(s, kExpScale, j == 3 ? lsn3 : sa.lsnx));   // depth_loss.h:211-226
                    const float dpx = px[j + 1] - px[j];
                    tx[j + 1] = with_sign0(wx, dpx);
                    sg_smx = fmaf(wx, fabsf(dpx), sg_smx);
                    flag |= (dpx != dpx);
                }
                {   // lane 0: the edge to its left neighbour
                    const float s = fabsf(Ix[0][0] - Il[0]) + fabsf(Ix[1][0] - Il[1]) + fabsf(Ix[2][0] - Il[2]);
                    const float wl = ex2_approx(fmaf(s, kExpScale, lsnl));
                    const float dl = cp[0] - pl;
                    tx[0] = with_sign0(wl, dl);
                    flag |= (lane == 0) & (dl != dl);
                }
            }

            // the left lane's last edge is this lane's first: shuffled here so that its latency hides behind the pointwise
            // terms (the exact tier, rare, repeats it after its fix-up)
            auto left_from_lane = [&]() {
                const float sl = __shfl_up_sync(0xffffffffu, sx[4], 1);
                if (lane != 0) sx[0] = sl;
                if constexpr (SMOOTH) {
                    const float tl = __shfl_up_sync(0xffffffffu, tx[4], 1);
                    if (lane != 0) tx[0] = tl;
                }
            };
            left_from_lane();

            // 3. pointwise terms
            const bool um[4] = {mk4.x != 0, mk4.y != 0, mk4.z != 0, mk4.w != 0};
            float rg[4], pw[4];
            bool m[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                rg[k] = in_range_pos(cp[k], eps, 1000.0f) ? rcp_approx(cp[k]) : 0.f;     // clamp backward / p
                m[k] = (HAS_MASK ? um[k] : (cg[k] > eps)) && lane_in;
                float gsum = (k < 2 ? ccv.x : ccv.y);
                if constexpr (SI) {
                    if (m[k]) gsum = fmaf(fmaf(c1n, C.d[k], c2), rg[k], gsum);           // depth_loss.h:38-63
                }
                pw[k] = gsum;
            }
            if constexpr (RP) {
                const float ayv = gyf - cyv;                                             // (v - cy)
                const float2 ay2 = make_float2(ayv, ayv);
                const float2 rfx2 = make_float2(rfx, rfx), rfy2 = make_float2(rfy, rfy);
                const float2 fxe2 = make_float2(fxe, fxe), fye2 = make_float2(fye, fye);
                const float2 yh2 = __fmul2_rn(ay2, rfy2);                                // d pY / d p (tolerance path)
                const float2 rpn2 = make_float2(rpn, rpn), eps2 = make_float2(eps, eps);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float2 pp = h ? make_float2(C.p.z, C.p.w) : make_float2(C.p.x, C.p.y);
                    const float2 gg = h ? make_float2(C.g.z, C.g.w) : make_float2(C.g.x, C.g.y);
                    const float2 ax = h ? ax23 : ax01;
                    const float2 xh = __fmul2_rn(ax, rfx2);                              // d pX / d p: tolerance path
                    // same operations, same order as depth_loss.h:299-315: X = ((u - cx) * d) / (fx + eps)
                    float2 pX, gX, pY, gY;
                    const float2 tpx = __fmul2_rn(ax, pp), tgx = __fmul2_rn(ax, gg);
                    const float2 tpy = __fmul2_rn(ay2, pp), tgy = __fmul2_rn(ay2, gg);
                    if constexpr (MK) {
                        // Markstein: q0 = t * rb, rem = t - q0 * b (exact), q = q0 + rem * rb = RN(t / b)
                        float2 q0 = __fmul2_rn(tpx, rfx2); pX = __ffma2_rn(__ffma2_rn(neg2(q0), fxe2, tpx), rfx2, q0);
                        q0 = __fmul2_rn(tgx, rfx2);        gX = __ffma2_rn(__ffma2_rn(neg2(q0), fxe2, tgx), rfx2, q0);
                        q0 = __fmul2_rn(tpy, rfy2);        pY = __ffma2_rn(__ffma2_rn(neg2(q0), fye2, tpy), rfy2, q0);
                        q0 = __fmul2_rn(tgy, rfy2);        gY = __ffma2_rn(__ffma2_rn(neg2(q0), fye2, tgy), rfy2, q0);
                    } else {
                        // a divisor Markstein's scheme does not cover (all-ones significand; never in practice): the
                        // quotient through double precision, correctly rounded too (see div_via_double) and without
                        // a call -- a call on this cold path constrained the registers of the whole row
                        pX = make_float2(div_via_double(tpx.x, fxe, rfx), div_via_double(tpx.y, fxe, rfx));
                        gX = make_float2(div_via_double(tgx.x, fxe, rfx), div_via_double(tgx.y, fxe, rfx));
                        pY = make_float2(div_via_double(tpy.x, fye, rfy), div_via_double(tpy.y, fye, rfy));
                        gY = make_float2(div_via_double(tgy.x, fye, rfy), div_via_double(tgy.y, fye, rfy));
                    }
                    const float2 dX = __fadd2_rn(pX, neg2(gX)), dY = __fadd2_rn(pY, neg2(gY)), dZ = __fadd2_rn(pp, neg2(gg));
                    const float2 ss = __fadd2_rn(__ffma2_rn(dZ, dZ, __ffma2_rn(dY, dY, __fmul2_rn(dX, dX))), eps2);
                    float2 re = make_float2(rsqrt_approx(ss.x), rsqrt_approx(ss.y));
                    re.x = m[2 * h] ? re.x : 0.f;                                        // depth_loss.h:318-320
                    re.y = m[2 * h + 1] ? re.y : 0.f;
                    sg_rp = __ffma2_rn(ss, re, sg_rp);                                   // e = sqrt(ss)
                    const float2 t = __fmul2_rn(__ffma2_rn(dX, xh, __ffma2_rn(dY, yh2, dZ)), re);
                    const float2 o = __ffma2_rn(t, rpn2, make_float2(pw[2 * h], pw[2 * h + 1]));
                    pw[2 * h] = o.x; pw[2 * h + 1] = o.y;
                }
            }

            // 4. the next row: wait for its slot, its log differences, the vertical edges
            // (a branch over a cold block, not an if / else: the phase bit flips either way)
            const unsigned par_n = (ph >> sn) & 1u;
            ph ^= 1u << sn;
            if (__any_sync(0xffffffffu, !n_ready)) {
#ifdef CADL_S3_TRACE
                const unsigned long long w0 = gtime_ns();
                mbar_wait_parity(bars + sn, par_n, a.results, gwarp, sn);
                tr_blocked += gtime_ns() - w0; ++tr_late;
#else
                mbar_wait_parity(bars + sn, par_n, a.results, gwarp, sn);
#endif
            }
            read_row(sn, N);
            yedges(C, N, cq, myq + sn * (SLOT), sy_dn, ty_dn, flag);

            // 5. exact tier, warp-uniform and rare (about 1 % of the warp-rows on BASELINE's data)
            if (__any_sync(0xffffffffu, flag)) {
                const ExactSigns x = exact_tier<SMOOTH>(cq, myq + sn * (SLOT), eps);
                sx[0] = x.sx[0] * inxl;
#pragma unroll
                for (int j = 0; j < 4; ++j) sx[j + 1] = x.sx[j + 1] * (j == 3 ? inx3 : inx0);
#pragma unroll
                for (int k = 0; k < 4; ++k) sy_dn[k] = x.sy[k] * iny0;
                if constexpr (SMOOTH) {
#pragma unroll
                    for (int j = 0; j < 5; ++j) tx[j] = fabsf(tx[j]) * x.tx[j];
#pragma unroll
                    for (int k = 0; k < 4; ++k) ty_dn[k] = fabsf(ty_dn[k]) * x.ty[k];
                }
                left_from_lane();
            }

            // 6. assembly and the 128-bit store
            float out[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float gmk = (sx[k] - sx[k + 1]) + (sy_up[k] - sy_dn[k]);
                float gsum = fmaf(gmk, rg[k], pw[k]);
                if constexpr (SMOOTH) gsum = fmaf((tx[k] - tx[k + 1]) + (ty_up[k] - ty_dn[k]), ab, gsum);
                out[k] = gsum;
            }
            if (want_grad && lane_in)
                *reinterpret_cast<float4*>(a.grad + (img + gy * W + gx0)) = make_float4(out[0], out[1], out[2], out[3]);
        };

        // prologue: the row above this share only contributes its lower edges
        Row3 RA, RB;
        float u0s[4] = {0.f, 0.f, 0.f, 0.f}, u0t[4] = {0.f, 0.f, 0.f, 0.f}, u1s[4], u1t[4];
        int ic = 0, sc = 0;
        if (ys > 0) {
            wait_slot(0);
            read_row(0, RB);
            wait_slot(1);
            read_row(1, RA);
            bool flag = false;
            const float k_gy = sg_gy;
            const float2 k_smy = sg_smy;
            yedges(RB, RA, myq, myq + SLOT, u0s, u0t, flag);
            sg_gy = k_gy; sg_smy = k_smy;                    // that edge is counted by the share above
            if (__any_sync(0xffffffffu, flag)) {
                const ExactSigns x = exact_tier<SMOOTH>(myq, myq + SLOT, eps);      // (only its vertical signs are used)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    u0s[k] = x.sy[k] * iny0;
                    if constexpr (SMOOTH) u0t[k] = fabsf(u0t[k]) * x.ty[k];
                }
            }
            ic = 1; sc = 1;
        } else {
            wait_slot(0);
            read_row(0, RA);                                 // first image row: no edge above
        }
        // two rows per trip, the two row buffers and the two edge buffers trading places: no register copies.
        // (Measured, not adopted: the ring slot as a compile-time tag with the loop unrolled over a whole turn of the ring
        //  -- every shared-memory address an immediate -- 73.7 us with 4 slots / 4 rows per trip, 101 us with 5 / 10,
        //  against 69.9: the two-row body is already 17 KB of code and the instruction cache decides.  One row per trip
        //  with register copies: 72.2.)
        auto rows = [&](auto mk) {
            float gyf = (float)ys;
            for (int gy = ys; gy < ye; gy += 2) {
                step(mk, ic, sc, gy, gyf, RA, RB, u0s, u0t, u1s, u1t);
                if (gy + 1 >= ye) break;
                sc = sc == D - 1 ? 0 : sc + 1;
                step(mk, ic + 1, sc, gy + 1, gyf + 1.f, RB, RA, u1s, u1t, u0s, u0t);
                sc = sc == D - 1 ? 0 : sc + 1;
                gyf += 2.f;
                ic += 2;
            }
        };
        if (mk_ok) rows(std::true_type{});
        else rows(std::false_type{});

#ifdef CADL_S3_TRACE
        tr_main = gtime_ns();
#endif
        // this share's sums -> the image's fixed-point accumulators (integer atomics: any order, same result)
        {
            float it_gx = 0.f, it_gy = 0.f, it_smx = 0.f, it_smy = 0.f, it_rp = 0.f;
            if (lane_in) {
                it_gx = sg_gx + (border_r ? 0.f : sg_gx3);
                it_gy = sg_gy;
                it_smx = sg_smx;
                it_smy = sg_smy.x + sg_smy.y;
                it_rp = sg_rp.x + sg_rp.y;
            }
            float v[IQ_COUNT];
            v[IQ_GX0] = warp_sum(it_gx);
            v[IQ_GY0] = warp_sum(it_gy);
            v[IQ_SMX] = SMOOTH ? warp_sum(it_smx) : 0.f;
            v[IQ_SMY] = SMOOTH ? warp_sum(it_smy) : 0.f;
            v[IQ_RP] = RP ? warp_sum(it_rp) : 0.f;
            float mine = 0.f;
#pragma unroll
            for (int q = 0; q < IQ_COUNT; ++q) mine = (lane == q) ? v[q] : mine;
            ImgRec* rec = sa.img + b;
            if (lane < IQ_COUNT) {
                double dv = (double)mine;
                if (lane == IQ_GX0 || lane == IQ_GY0) dv *= 0.69314718055994531;    // log2 units -> natural log
                if (lane == IQ_SMX) dv *= sa.inv_snx;
                if (lane == IQ_SMY) dv *= sa.inv_sny;
                unsigned long long hi, lo;
                unsigned fl = 0u;
                fix_split(dv, hi, lo, fl, lane);
                if (hi) atomicAdd(&rec->hi[lane], hi);
                if (lo) atomicAdd(&rec->lo[lane], lo);
                if (fl) atomicOr(&rec->flags, fl);
            }
            __threadfence();
            __syncwarp();
#ifdef CADL_S3_TRACE
            tr_c[0] = gtime_ns();
#endif
            int last = 0;
            if (lane == 0) last = atomicAdd(&rec->cnt, 1u) == (unsigned)spi - 1u;
            last = __shfl_sync(0xffffffffu, last, 0);
#ifdef CADL_S3_TRACE
            tr_c[1] = gtime_ns();
#endif
            if (last && SMOOTH) {
                // the image is complete: its share of the smoothness loss and the gradient offset.
                // (acquire side only -- an L1 invalidate; a full fence here waits for the SM's write queue, which the
                //  other warps' offset reductions fill at this point of the kernel)
                asm volatile("fence.acquire.gpu;" ::: "memory");
                if (lane == 0) {
                    const unsigned fl = __ldcg(&rec->flags);
                    const double sx_ = fix_join(__ldcg(&rec->hi[IQ_SMX]), __ldcg(&rec->lo[IQ_SMX]), fl, IQ_SMX);
                    const double sy_ = fix_join(__ldcg(&rec->hi[IQ_SMY]), __ldcg(&rec->lo[IQ_SMY]), fl, IQ_SMY);
                    double Lb;
                    float off;
                    smooth_share3(a, sa, b, sx_, sy_, Lb, off);
                    rec->Lb = Lb;
                    rec->off = off;
                    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(&rec->ready), "r"(epoch) : "memory");
                }
            }
        }
    }   // items

    // ---- the last warp to finish writes the results and returns the workspace to zero ----
    // (before the offset pass: the results do not depend on it, and its reductions need not be waited for)
    if (gwarp >= sa.nwarps) return;
    __syncwarp();
#ifdef CADL_S3_TRACE
    tr_c[2] = gtime_ns();
#endif
    unsigned ticket = 0;
    if (lane == 0) ticket = atomicAdd(sa.done, 1u);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
#ifdef CADL_S3_TRACE
    tr_c[3] = gtime_ns();
#endif
    const bool fin = ticket == (unsigned)sa.nwarps - 1u;
    // the metric results need phase A's statistics only: the FIRST warp to get here writes them (it would otherwise
    // just wait for its image), not the last one
    if (ticket == 0u && a.metrics) write_metric_results(a.stats, a.metrics, *a.results, lane);
    if (fin) {
        asm volatile("fence.acquire.gpu;" ::: "memory");      // (every warp fenced its sums before its ticket)
        // Every load of this block is issued before the first use: the block runs while the other warps pour their
        // offset reductions into the L2, a round trip takes microseconds then, and a chain of dependent loads here
        // was a ~20 us tail on the whole kernel (profiles/r02_trace.py).
        const double stv = a.stats[lane];                    // ST_COUNT == 32: one statistic per lane
        double pqv = 0.0;
        {
            const unsigned pfl = __ldcg(reinterpret_cast<const unsigned*>(sa.pyr_rec + 12));
            const int q = lane < 6 ? lane : 0;
            pqv = fix_join(__ldcg(sa.pyr_rec + q), __ldcg(sa.pyr_rec + 6 + q), pfl, q);      // lane q: pooled-scale sum q
        }
        // fixed order: lane-strided over the images, fixed shuffle tree
        double tq[IQ_COUNT], tl = 0.0;
#pragma unroll
        for (int q = 0; q < IQ_COUNT; ++q) tq[q] = 0.0;
        for (int i = lane; i < a.B; i += 32) {
            ImgRec* rec = sa.img + i;
            const unsigned fl = __ldcg(&rec->flags);
#pragma unroll
            for (int q = 0; q < IQ_COUNT; ++q) tq[q] += fix_join(__ldcg(&rec->hi[q]), __ldcg(&rec->lo[q]), fl, q);
            if (SMOOTH) {
                double Lb = __ldcg(&rec->Lb);
                float off = __ldcg(&rec->off);
                if (!want_grad) {      // forward only: nobody computed the shares yet (ready is not used)
                    smooth_share3(a, sa, i, fix_join(__ldcg(&rec->hi[IQ_SMX]), __ldcg(&rec->lo[IQ_SMX]), fl, IQ_SMX),
                                       fix_join(__ldcg(&rec->hi[IQ_SMY]), __ldcg(&rec->lo[IQ_SMY]), fl, IQ_SMY), Lb, off);
                }
                a.img_sm[2 * i] = Lb;
                a.img_off[i] = off;
                tl += Lb;
            }
            // leave the record clean for the next call
#pragma unroll
            for (int q = 0; q < IQ_COUNT; ++q) { rec->hi[q] = 0ull; rec->lo[q] = 0ull; }
            rec->cnt = 0u; rec->flags = 0u;          // (off, Lb, ready are overwritten / epoch-valued)
        }
#ifdef CADL_S3_TRACE
        tr_c[4] = gtime_ns();
#endif
#pragma unroll
        for (int q = 0; q < IQ_COUNT; ++q) tq[q] = warp_sum(tq[q]);
        tl = warp_sum(tl);
#ifdef CADL_S3_TRACE
        tr_c[5] = gtime_ns();
#endif
        // loss sums of the pooled scales: pyr_coef_kernel's fixed-point totals (one record, no rows to fold here)
        double pq[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) pq[q] = __shfl_sync(0xffffffffu, pqv, q);
        static_assert(ST_COUNT == 32, "one statistic per lane");
        const double st_si_n = __shfl_sync(0xffffffffu, stv, ST_SI_N), st_si_s = __shfl_sync(0xffffffffu, stv, ST_SI_S);
        const double st_si_q = __shfl_sync(0xffffffffu, stv, ST_SI_Q), st_rp_n = __shfl_sync(0xffffffffu, stv, ST_RP_N);
        if (lane == 0) {
            cadl_results& r = *a.results;
            // gradient matching: scale 0 from this kernel, scales 1..3 from pyr_coef_kernel     depth_loss.h:162-163
            // (reciprocals of the shape constants from the host, one reciprocal per count: this block is the end of
            //  the kernel's critical chain and eight dependent double divisions were 1.2 us of it)
            double gm = 0.0;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const double sx_ = s == 0 ? tq[IQ_GX0] : pq[2 * (s - 1)];
                const double sy_ = s == 0 ? tq[IQ_GY0] : pq[2 * (s - 1) + 1];
                gm += sx_ * sa.rnx[s] + sy_ * sa.rny[s];                                      // NaN where a scale has no edges (0 * inf)
            }
            gm *= 0.25;
            double si = 0.0, rp = 0.0;
            if (SI) {                                                                         // depth_loss.h:58-63
                const double n = st_si_n;
                if (n > 0.0) {
                    const double rn = 1.0 / n, m = st_si_s * rn;
                    si = st_si_q * rn - (double)a.lambda * m * m;
                }
            }
            if (RP) {                                                                         // depth_loss.h:323-330
                const double n = st_rp_n;
                if (n > 0.0) rp = tq[IQ_RP] / n;
            }
            const double sm = SMOOTH ? tl : 0.0;
            r.n_si = SI ? (int64_t)st_si_n : 0;
            r.n_reproj = RP ? (int64_t)st_rp_n : 0;
            r.d_si = si; r.d_grad = gm; r.d_smooth = sm; r.d_reproj = rp;
            r.loss_si = (float)si; r.loss_grad = (float)gm; r.loss_smooth = (float)sm; r.loss_reproj = (float)rp;
            // depth_loss.h:427-430, in float like the reference's tensor arithmetic
            float tot = 0.f;
            if (SI) tot = a.w_si * r.loss_si;
            tot = tot + a.w_grad * r.loss_grad;
            if (SMOOTH) tot = tot + a.w_smooth * r.loss_smooth;
            if (RP) tot = tot + a.w_rp * r.loss_reproj;
            r.loss_total = tot;
            r.d_total = (double)a.w_si * si + (double)a.w_grad * gm + (double)a.w_smooth * sm + (double)a.w_rp * rp;
            *sa.done = 0u;
            *sa.epoch = epoch;
        }
    }

#ifdef CADL_S3_TRACE
    if (fin) {
        tr_c[6] = gtime_ns();
        if (lane == 0) {
            unsigned long long* t = g_s3_trace + (size_t)4095 * kS3TraceWords;
            for (int q = 0; q < 7; ++q) t[q] = tr_c[q] - tr_main;
            t[7] = tr_main - tr_t0;
        }
    }
#endif
    // ---- grad[rows this warp wrote] -= off[b], once the image's sums are complete (rows are L2-resident) ----
    if (SMOOTH && want_grad) {
        for (int item = gwarp; item < nitems; item += sa.nwarps) {
            const Share3 sh = s3_share(item, a.B, H, sa);
            const int b = sh.b, strip = sh.strip, ys = sh.ys, ye = sh.ye;
            const ImgRec* rec = sa.img + b;
            unsigned rdy = 0, spins = 0;
            unsigned long long t0 = 0;
            // relaxed polls with a pause (an acquire load invalidates the L1 on every probe and a busy loop takes issue
            // slots from the warps of this SM that are still working); one acquire fence once the flag is seen
            for (;;) {
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(rdy) : "l"(&rec->ready) : "memory");
                if (rdy == epoch) break;
                __nanosleep(256);
                if ((++spins & 63u) == 0u) {
                    const unsigned long long t = gtime_ns();
                    if (t0 == 0) t0 = t;
                    else if (t - t0 > kS3WaitLimitNs) s3_bail(a.results, 2, gwarp, b);
                }
            }
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
#ifdef CADL_S3_TRACE
            tr_ready = gtime_ns();
#endif
            float off;
            asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(off) : "l"(&rec->off) : "memory");
            // grad -= off as vector reductions executed at the L2 (same single fp32 rounding as a subtraction): nothing
            // comes back, so the warp does not wait for its rows a second time
            const float noff = -off;
            const int gx0 = strip * 128 + 4 * lane;
            if (gx0 < W) {
                // (independent address registers: a reduction holds its address register until it has left the SM, so
                //  a pointer that is bumped between two reductions stalls on the scoreboard; 43 rows -> 5 trips of 8)
                float* gp = a.grad + (b * plane + ys * W + gx0);
                int y = ys;
                const size_t w1 = (size_t)W;
                for (; y + 8 <= ye; y += 8) {
                    float* g0 = gp; float* g1 = gp + w1; float* g2 = gp + 2 * w1; float* g3 = gp + 3 * w1;
                    float* g4 = gp + 4 * w1; float* g5 = gp + 5 * w1; float* g6 = gp + 6 * w1; float* g7 = gp + 7 * w1;
                    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %1, %1, %1};" :: "l"(g0), "f"(noff) : "memory");
                    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %1, %1, %1};" :: "l"(g1), "f"(noff) : "memory");
                    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %1, %1, %1};" :: "l"(g2), "f"(noff) : "memory");
                    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %1, %1, %1};" :: "l"(g3), "f"(noff) : "memory");
                    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %1, %1, %1};" :: "l"(g4), "f"(noff) : "memory");
                    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %1, %1, %1};" :: "l"(g5), "f"(noff) : "memory");
                    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %1, %1, %1};" :: "l"(g6), "f"(noff) : "memory");
                    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %1, %1, %1};" :: "l"(g7), "f"(noff) : "memory");
                    gp += 8 * w1;
                }
                for (; y < ye; ++y) {
                    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %1, %1, %1};" :: "l"(gp), "f"(noff) : "memory");
                    gp += W;
                }
            }
        }
    }
#ifdef CADL_S3_TRACE
    if (lane == 0 && gwarp < 4096) {
        unsigned long long* t = g_s3_trace + (size_t)gwarp * kS3TraceWords;
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        t[0] = tr_t0; t[1] = tr_main; t[2] = tr_ready; t[3] = gtime_ns(); t[4] = tr_late; t[5] = tr_blocked; t[6] = smid; t[7] = 0;
    }
#endif

}

}  // namespace cadl
