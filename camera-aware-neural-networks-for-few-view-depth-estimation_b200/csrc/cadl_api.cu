// cadl C ABI (include/cadl.h): argument checks, workspace carving, kernel dispatch.
// Built only for sm_100a; there is no host/CPU implementation behind these entry points.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "cadl.h"
#include "cadl_common.cuh"
#include "cadl_phase_a.cuh"
#include "cadl_phase_b.cuh"
#include "cadl_phase_b_fast.cuh"
#include "cadl_phase_b_stream.cuh"
#include "cadl_stream3_host.h"
#include "cadl_rays.cuh"
#include "cadl_photometric.cuh"
#include "cadl_next.cuh"

using namespace cadl;

namespace cadl {
// Self-test kernels (tests/test_math_gpu.py): the replicas in cadl_math.cuh against the CUDA library forms.
__global__ void selftest_log_kernel(unsigned lo, unsigned hi, unsigned long long* mism) {
    unsigned long long bad = 0;
    for (unsigned long long u = (unsigned long long)lo + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; u <= hi;
         u += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned)u);
        const float ref = logf(x);
        const float a = log_exact(x);
        const float2 b = log_exact2(make_float2(x, __uint_as_float((unsigned)(hi - (u - lo)))));
        bad += (__float_as_uint(a) != __float_as_uint(ref)) + (__float_as_uint(b.x) != __float_as_uint(ref));
        bad += (__float_as_uint(b.y) != __float_as_uint(logf(__uint_as_float((unsigned)(hi - (u - lo))))));
    }
    if (bad) atomicAdd(mism, bad);
}
// which 2: inputs whose lg2.approx.ftz differs from log2 (fp64) by more than tol (absolute) -- the bound the two-tier
// sign logic of cadl_stream3.cuh and the threshold bands of phase A rest on
__global__ void selftest_lg2_kernel(unsigned lo, unsigned hi, float tol, unsigned long long* mism) {
    unsigned long long bad = 0;
    for (unsigned long long u = (unsigned long long)lo + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; u <= hi;
         u += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned)u);
        const double err = fabs((double)lg2_approx(x) - log2((double)x));
        bad += !(err <= (double)tol);
    }
    if (bad) atomicAdd(mism, bad);
}
__global__ void selftest_div_kernel(unsigned lo, unsigned hi, float b, unsigned long long* mism) {
    const float rb = __frcp_rn(b);
    unsigned long long bad = 0;
    for (unsigned long long u = (unsigned long long)lo + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; u <= hi;
         u += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned)u);
        bad += (__float_as_uint(div_by_const(x, b, rb)) != __float_as_uint(__fdiv_rn(x, b)));
        bad += (__float_as_uint(div_by_const(-x, b, rb)) != __float_as_uint(__fdiv_rn(-x, b)));
    }
    if (bad) atomicAdd(mism, bad);
}
// div_via_double against __fdiv_rn, for ANY divisor (the gradient pass takes it where Markstein's scheme is not safe)
__global__ void selftest_div2_kernel(unsigned lo, unsigned hi, float b, unsigned long long* mism) {
    const float rb = __frcp_rn(b);
    unsigned long long bad = 0;
    for (unsigned long long u = (unsigned long long)lo + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; u <= hi;
         u += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned)u);
        bad += (__float_as_uint(div_via_double(x, b, rb)) != __float_as_uint(__fdiv_rn(x, b)));
        bad += (__float_as_uint(div_via_double(-x, b, rb)) != __float_as_uint(__fdiv_rn(-x, b)));
    }
    if (bad) atomicAdd(mism, bad);
}
}  // namespace cadl

namespace {

inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? CADL_OK : CADL_ERR_CUDA + (int)e; }
inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

thread_local char g_err_detail[256];

constexpr int kMaxDevices = 64;
inline int current_device() {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
    return dev;
}

inline bool markstein_safe_host(float b) {
    uint32_t u;
    memcpy(&u, &b, 4);
    u &= 0x7fffffffu;
    return (u >= 0x00800000u) && (u < 0x7f000000u) && ((u & 0x007fffffu) != 0x007fffffu);
}

struct Ws {
    WsLayout L;
    char* base;
    WsHeader* hdr() const { return reinterpret_cast<WsHeader*>(base + L.header); }
    double* stats() const { return reinterpret_cast<double*>(base + L.stats); }
    double* img_psum() const { return reinterpret_cast<double*>(base + L.img_psum); }
    double* a_part() const { return reinterpret_cast<double*>(base + L.a_part); }
    double* b_part() const { return reinterpret_cast<double*>(base + L.b_part); }
    double* img_sm() const { return reinterpret_cast<double*>(base + L.img_sm); }
    float* img_off() const { return reinterpret_cast<float*>(base + L.img_off); }
    ImgRec* img_rec() const { return reinterpret_cast<ImgRec*>(base + L.img_rec); }
    unsigned long long* img_words() const { return reinterpret_cast<unsigned long long*>(base + L.img_words); }
    bool has_pyr() const { return L.pyr_blocks > 0; }
    PyrArrays pyr() const {
        PyrArrays p;
        for (int s = 0; s < 3; ++s) {
            p.lp[s] = reinterpret_cast<float*>(base + L.pyr_lp[s]);
            p.lg[s] = reinterpret_cast<float*>(base + L.pyr_lg[s]);
            p.rq[s] = reinterpret_cast<float*>(base + L.pyr_rq[s]);
        }
        p.c1 = reinterpret_cast<float*>(base + L.pyr_c1);
        return p;
    }
};

// Dispatch switches and per-launch timing exist only in the debug build (libcadl_dbg.so, -DCADL_DEBUG), which the
// tests use to compare the kernels of one term set against each other; the product library has neither the entry
// points nor the state.
#ifdef CADL_DEBUG
int g_force_generic = 0;    // bit 0: generic (any shape) phase-B kernel even for aligned shapes
int g_force_no_tma = 0;     // bit 1 (with bit 3): tile kernel staged with cp.async instead of TMA
int g_force_tile = 0;       // bit 3: the one-CTA-per-tile fast kernel instead of the streaming kernel
int g_no_pdl = 0;           // bit 4: plain stream-ordered launches (no programmatic dependent launch)
int g_no_overlap = 0;       // bit 5: pooled-pyramid kernels in line on the caller's stream instead of beside phase A
int g_no_coop = 0;          // bit 6: reprojection alone keeps the separate count kernel (no cooperative launch)
int g_no_poolstats = 0;     // bit 7: the loss statistics from phase A (the split API's kernels) instead of the pooled-sum pass
// cadl_debug_kernel_times: CUDA events between the launches of one cadl_stack_fwd_bwd call
struct KTimes {
    bool on = false;
    int n = 0, n_last = 0;
    cudaEvent_t ev[12] = {};
    const char* name[12] = {};
    const char* name_last[12] = {};
    float ms_last[12] = {};
};
KTimes g_kt;
void kt_mark(cudaStream_t st, const char* name) {
    if (!g_kt.on || g_kt.n >= 12) return;
    if (!g_kt.ev[g_kt.n]) cudaEventCreate(&g_kt.ev[g_kt.n]);
    cudaEventRecord(g_kt.ev[g_kt.n], st);
    g_kt.name[g_kt.n++] = name;
}
void kt_finish() {
    if (!g_kt.on || g_kt.n < 2) return;
    cudaEventSynchronize(g_kt.ev[g_kt.n - 1]);
    g_kt.n_last = g_kt.n - 1;
    for (int i = 1; i < g_kt.n; ++i) {
        cudaEventElapsedTime(&g_kt.ms_last[i - 1], g_kt.ev[i - 1], g_kt.ev[i]);
        g_kt.name_last[i - 1] = g_kt.name[i];
    }
    g_kt.n = 0;
}
inline bool kt_on() { return g_kt.on; }
#else
constexpr int g_force_generic = 0, g_force_no_tma = 0, g_force_tile = 0, g_no_pdl = 0, g_no_overlap = 0, g_no_coop = 0, g_no_poolstats = 0;
inline void kt_mark(cudaStream_t, const char*) {}
inline void kt_finish() {}
inline bool kt_on() { return false; }
#endif

WsLayout layout_for(int B, int H, int W) {
    WsLayout L = ws_layout(B, H, W);
    return L;
}

int check_common(int B, int H, int W, const void* ws, size_t ws_bytes) {
    if (B < 1 || H < 1 || W < 1) return CADL_ERR_SHAPE;
    if ((long long)B * H * W > 0x7fffffffLL) return CADL_ERR_SHAPE;
    if (!ws) return CADL_ERR_NULL;
    if (!aligned(ws, 256)) return CADL_ERR_WORKSPACE;
    if (ws_bytes < cadl_workspace_bytes(B, H, W)) return CADL_ERR_WORKSPACE;
    return CADL_OK;
}

uint32_t phase_a_flags(const cadl_params& p) {
    uint32_t f = 0;
    if (p.terms & CADL_TERM_SI) f |= FA_SI;
    if (p.terms & CADL_TERM_REPROJ) f |= FA_RP;
    if (p.terms & CADL_TERM_SMOOTH) f |= FA_PSUM;
    if (p.metrics & CADL_METRICS_EVAL) f |= FA_EV;
    if (p.metrics & CADL_METRICS_TRAIN) f |= FA_TR;
    return f;
}

template <int F>
cudaError_t launch_a(const PhaseAArgs& a, dim3 grid, cudaStream_t st) {
    if (a.mask) phase_a_kernel<F, true><<<grid, kThreadsA, 0, st>>>(a);
    else phase_a_kernel<F, false><<<grid, kThreadsA, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t dispatch_a(uint32_t f, const PhaseAArgs& a, dim3 grid, cudaStream_t st) {
    switch (f) {
#define CADL_A(n) case n: return launch_a<n>(a, grid, st);
        CADL_A(1) CADL_A(2) CADL_A(3) CADL_A(4) CADL_A(5) CADL_A(6) CADL_A(7) CADL_A(8) CADL_A(9)
        CADL_A(10) CADL_A(11) CADL_A(12) CADL_A(13) CADL_A(14) CADL_A(15) CADL_A(16) CADL_A(17)
        CADL_A(18) CADL_A(19) CADL_A(20) CADL_A(21) CADL_A(22) CADL_A(23) CADL_A(24) CADL_A(25)
        CADL_A(26) CADL_A(27) CADL_A(28) CADL_A(29) CADL_A(30) CADL_A(31)
#undef CADL_A
        default: return cudaErrorInvalidValue;
    }
}

template <int F>
cudaError_t launch_tile(const PhaseBArgs& a, cudaStream_t st) {
    static bool configured[kMaxDevices] = {};   // the attribute applies to the current device only
    const int dev = current_device();
    if (dev < 0) return cudaErrorInvalidDevice;
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(phase_b_tile_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kTileSmemBytes);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    phase_b_tile_kernel<F><<<a.b_rows, kThreadsB, kTileSmemBytes, st>>>(a);
    return cudaGetLastError();
}

// ---- TMA descriptors (cuTensorMapEncodeTiled through the runtime's driver entry point: no libcuda link) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)f;
    }
    return fn;
}
// (W, H, B) fp32 tensor, box = (FRW, FRH, 1), zero fill outside
bool make_tile_map(CUtensorMap* m, const float* base, int B, int H, int W) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)FRW, (cuuint32_t)FRH, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
struct TileMaps {
    CUtensorMap pred, gt;
    const float *pp = nullptr, *pg = nullptr;
    int B = 0, H = 0, W = 0;
    bool ok = false;
};
thread_local TileMaps g_maps;   // re-encoded only when a pointer or the shape changes

template <int F, bool M>
cudaError_t launch_fast_m(PhaseBArgs& a, cudaStream_t st) {
    static bool configured[kMaxDevices] = {};
    const int dev = current_device();
    if (dev < 0) return cudaErrorInvalidDevice;
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(phase_b_fast_kernel<F, M>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kFastSmemBytes);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    a.use_tma = 0;
    if ((F & FB_GRAD) && !g_force_no_tma) {
        TileMaps& tm = g_maps;
        if (!(tm.ok && tm.pp == a.pred && tm.pg == a.gt && tm.B == a.B && tm.H == a.H && tm.W == a.W)) {
            tm.ok = make_tile_map(&tm.pred, a.pred, a.B, a.H, a.W) && make_tile_map(&tm.gt, a.gt, a.B, a.H, a.W);
            tm.pp = a.pred; tm.pg = a.gt; tm.B = a.B; tm.H = a.H; tm.W = a.W;
        }
        a.use_tma = tm.ok ? 1 : 0;
    }
    phase_b_fast_kernel<F, M><<<a.b_rows, kThreadsB, kFastSmemBytes, st>>>(a, g_maps.pred, g_maps.gt);
    return cudaGetLastError();
}
template <int F>
cudaError_t launch_fast(PhaseBArgs& a, cudaStream_t st) {
    return a.mask ? launch_fast_m<F, true>(a, st) : launch_fast_m<F, false>(a, st);
}

// Launch with programmatic stream serialization: the kernel may start while its predecessor in the stream drains;
// it calls pdl_wait() (cadl_math.cuh) before it touches anything the predecessor wrote.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

int num_sms_cached() {         // of the CURRENT device (a process may drive several)
    static int num_sms[kMaxDevices] = {};
    const int dev = current_device();
    if (dev < 0) return 1;
    if (!num_sms[dev]) cudaDeviceGetAttribute(&num_sms[dev], cudaDevAttrMultiProcessorCount, dev);
    return num_sms[dev] > 0 ? num_sms[dev] : 1;
}

// How one cadl_stack_fwd_bwd call is laid out over launches (decided once, used by the reduce and the gradient part).
struct StepPlan {
    int pool_stats = 0;             // PS_* bits: the pooled-sum kernel also produces the loss statistics (no phase A for them)
    bool pyr_prelaunched = false;   // the pooled-pyramid kernels already run on the auxiliary stream, beside phase A
    int pyr_grid = 0;               // CTAs (= partial rows) of pyr_coef_kernel; 0 = one thread per 8x8 block
    int a_blocks_per_img = 0;       // phase A grid override (0 = the workspace layout's)
};

// The pooled-pyramid kernels need pred/gt only, phase A needs pred/gt only, and neither needs the other: they run
// SIDE BY SIDE -- the pyramid as one persistent CTA per SM on an auxiliary stream (memory-latency bound, few issue
// slots), phase A on the caller's stream with its grid shrunk so that both fit in one resident wave.
struct AuxStream {
    cudaStream_t s2 = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    int dev = -1;
};
thread_local AuxStream g_aux[16];
AuxStream* aux_for_current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    AuxStream& x = g_aux[dev];
    if (x.dev != dev) {
        if (cudaStreamCreateWithFlags(&x.s2, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&x.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&x.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        x.dev = dev;
    }
    return &x;
}

cudaError_t launch_pyramid(const PhaseBArgs& a, const Ws& ws, cudaStream_t st, int grid, bool pdl_pool, bool pdl, int pool_stats = 0) {
    const PyrArrays py = ws.pyr();
    PoolStatsArgs ps{};
    cudaError_t e;
    unsigned int* recs = reinterpret_cast<unsigned int*>(ws.img_rec());
    if (pool_stats) {
        // the statistics of the loss terms ride on the pooled-sum pass (eps_si == eps_rp == eps_grad on this path)
        ps.mask = a.mask; ps.rec = reinterpret_cast<PoolStatsRec*>(ws.hdr()->pool_rec); ps.img_words = ws.img_words();
        ps.stats = ws.stats(); ps.img_psum = ws.img_psum(); ps.want_rp = (a.terms & CADL_TERM_REPROJ) ? 1 : 0;
        if (a.mask)
            e = launch_pdl(pyr_pool_kernel<PS_SI | PS_PSUM, true>, dim3(grid), dim3(256), st, pdl_pool, a.pred, a.gt, a.B, a.H, a.W, a.eps_grad, py, recs, ps);
        else
            e = launch_pdl(pyr_pool_kernel<PS_SI | PS_PSUM, false>, dim3(grid), dim3(256), st, pdl_pool, a.pred, a.gt, a.B, a.H, a.W, a.eps_grad, py, recs, ps);
    } else {
        e = launch_pdl(pyr_pool_kernel<0, false>, dim3(grid), dim3(256), st, pdl_pool, a.pred, a.gt, a.B, a.H, a.W, a.eps_grad, py, recs, ps);
    }
    if (e != cudaSuccess) return e;
    kt_mark(st, "pyr_pool_kernel");
    PyrCoefArgs ca{};
    ca.py = py; ca.B = a.B; ca.H = a.H; ca.W = a.W;
    for (int s = 0; s < 4; ++s) { ca.inv_nx[s] = a.inv_nx[s]; ca.inv_ny[s] = a.inv_ny[s]; }
    ca.wg = 0.25f * a.w_grad * a.upstream;          // 1/num_scales * weight * upstream
    ca.rec = reinterpret_cast<unsigned long long*>(ws.img_rec() + a.B);      // the record behind the B image records
    e = launch_pdl(pyr_coef_kernel, dim3(grid), dim3(256), st, pdl, ca);
    if (e != cudaSuccess) return e;
    kt_mark(st, "pyr_coef_kernel");
    return cudaSuccess;
}

// streaming split of the fast path: pooled pyramid -> coarse coefficients (cadl_phase_b_stream.cuh) -> full-resolution pass
// (cadl_stream3.cuh).  cudaErrorNotSupported: the caller takes the one-tile-per-CTA kernel instead.
template <int F>
cudaError_t launch_stream(PhaseBArgs& a, const Ws& ws, cudaStream_t st, bool* offset_done, const StepPlan& plan) {
    const PyrArrays py = ws.pyr();
    const int nblk = plan.pyr_grid > 0 ? plan.pyr_grid : ws.L.pyr_blocks;
    // (event records between the launches would serialise them anyway; after a cross-stream join the streaming
    //  kernel has two predecessors and is launched plainly)
    const bool pdl = !g_no_pdl && !kt_on();
    cudaError_t e = cudaSuccess;
    if (!(a.eps_grad >= 1e-30f && (!(F & FB_SI) || a.eps_si == a.eps_grad) && (!(F & FB_RP) || a.eps_rp == a.eps_grad)))
        return cudaErrorNotSupported;
    if (!plan.pyr_prelaunched) {
        e = launch_pyramid(a, ws, st, nblk, pdl, pdl, plan.pool_stats);
        if (e != cudaSuccess) return e;
    }
    // TMA row ring, two-tier logs, packed reprojection, smoothness offset applied in-kernel (cadl_stream3.cuh)
    Stream3Args s3{};
    s3.c1 = py.c1; s3.nstrip = (a.W + 127) / 128;
    s3.pyr_rec = reinterpret_cast<const unsigned long long*>(ws.img_rec() + a.B);
    s3.img = ws.img_rec(); s3.done = &ws.hdr()->ticket_b; s3.epoch = &ws.hdr()->pad[0];
    if ((F & FB_SMOOTH) && !stream3_fill_smooth(a, s3)) return cudaErrorNotSupported;
    e = launch_stream3(F, a, s3, st);
    if (e != cudaSuccess) return e;
    kt_mark(st, "stream3_kernel");
    *offset_done = true;
    return cudaSuccess;
}

template <int F>
cudaError_t launch_point_fast(PhaseBArgs& a, cudaStream_t st) {
    // (blocks per image, B): ONE resident wave -- the kernel is capped at 64 registers, 4 CTAs of 256 threads per SM
    // (the earlier sizing assumed 8 per SM and ran 1.6 waves); its warps deal the image's 128-pixel row segments
    // round-robin.  b_rows = grid size (partial rows, ticket).
    const int wpb = kThreadsB / 32;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, a.mask ? (const void*)phase_b_point_fast_kernel<F, true>
                                                                       : (const void*)phase_b_point_fast_kernel<F, false>,
                                                      kThreadsB, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    int bpi = (per_sm * num_sms_cached()) / a.B;
    const int max_bpi = (a.H * ((a.W + 127) / 128) + wpb - 1) / wpb;        // at least one segment per warp
    if (bpi > max_bpi) bpi = max_bpi;
    while (bpi > 1 && bpi * a.B > kPointBlocks) --bpi;
    if (bpi < 1) bpi = 1;
    dim3 grid(bpi, a.B);
    a.b_rows = bpi * a.B;
    const bool pdl = !g_no_pdl && !kt_on();
    if (a.mask) return launch_pdl(phase_b_point_fast_kernel<F, true>, grid, dim3(kThreadsB), st, pdl, a);
    return launch_pdl(phase_b_point_fast_kernel<F, false>, grid, dim3(kThreadsB), st, pdl, a);
}

// Reprojection alone, no metrics: count + gradient in ONE cooperative launch (phase_b_point_fast_kernel<.., COUNT>).
// Returns cudaErrorNotSupported when the grid cannot be co-resident, so the caller takes the two-launch path.
cudaError_t launch_point_count(PhaseBArgs& a, cudaStream_t st) {
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, current_device() < 0 ? 0 : current_device());
    if (!coop) return cudaErrorNotSupported;
    const int wpb = kThreadsB / 32;
    void* fn = a.mask ? (void*)phase_b_point_fast_kernel<FB_RP, true, true> : (void*)phase_b_point_fast_kernel<FB_RP, false, true>;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreadsB, 0);
    if (e != cudaSuccess) return e;
    int bpi = (per_sm * num_sms_cached()) / a.B;
    const int max_bpi = (a.H * ((a.W + 127) / 128) + wpb - 1) / wpb;
    if (bpi > max_bpi) bpi = max_bpi;
    if (bpi < 1) return cudaErrorNotSupported;                    // more images than resident CTAs
    if ((long long)bpi * a.B > (long long)per_sm * num_sms_cached()) return cudaErrorNotSupported;
    dim3 grid(bpi, a.B);
    a.b_rows = bpi * a.B;
    void* args[] = {(void*)&a};
    return cudaLaunchCooperativeKernel(fn, grid, dim3(kThreadsB), args, 0, st);
}

template <int F>
cudaError_t launch_point(const PhaseBArgs& a, cudaStream_t st) {
    phase_b_point_kernel<F><<<a.b_rows, kThreadsB, 0, st>>>(a);
    return cudaGetLastError();
}

// Grid split when the pyramid kernels run beside phase A, in CTAs of 256 threads out of the 4 per SM that fit.
// Phase A with the metric variants is the longer of the two: 2 pyramid CTAs per SM, the rest for phase A (config 3,
// us/step: 222 pyramid CTAs 207.4, 260 202.0, 280 195.2, 296 196.2, 330 198.9, 370 214.8; in line 208.7).  Without
// metrics (the trainers' step) phase A is short and the pyramid gets 3 per SM (296: 177.1, 370: 169.3, 444: 166.1,
// 518: 189.9; in line 174.8).  Everything must fit ONE wave: more, shorter phase-A blocks starve the pyramid (13
// blocks per image 207.0 us/step, 18: 204.8, 36: 208.7 against 9: 196.5).
// CTAs of 256 threads per SM of the pooled-sum / coefficient kernels in the step without a phase A (cadl_stack_fwd_bwd;
// 2: 116 us/step, 3, 4, 5: 108-109)
#ifndef CADL_PLAN2_PYR_N
#define CADL_PLAN2_PYR_N 4
#endif
StepPlan concurrent_plan(const Ws& ws, int B, bool metrics) {
    StepPlan plan;
    const int sms = num_sms_cached();
#ifndef CADL_PLAN_M
#define CADL_PLAN_M 2
#endif
#ifndef CADL_PLAN_N
#define CADL_PLAN_N 3
#endif
    const int pyr_ctas = (metrics ? CADL_PLAN_M : CADL_PLAN_N) * sms;
    plan.pyr_grid = ws.L.pyr_blocks < pyr_ctas ? ws.L.pyr_blocks : pyr_ctas;
    plan.a_blocks_per_img = (4 * sms - plan.pyr_grid) / B;       // phase A: 4 CTAs of 256 threads per SM
    if (plan.a_blocks_per_img < 1) plan.a_blocks_per_img = 1;
    plan.pyr_prelaunched = true;
    return plan;
}

cudaError_t prelaunch_pyramid(const PhaseBArgs& a, const Ws& ws, cudaStream_t st, AuxStream* aux, const StepPlan& plan) {
    cudaError_t e = cudaEventRecord(aux->fork, st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(aux->s2, aux->fork, 0);
    if (e == cudaSuccess) e = launch_pyramid(a, ws, aux->s2, plan.pyr_grid, false, !g_no_pdl);
    if (e == cudaSuccess) e = cudaEventRecord(aux->join, aux->s2);
    return e;
}

int run_reduce(const float* pred, const float* gt, const uint8_t* mask, int B, int H, int W,
               const cadl_params& p, const Ws& ws, cudaStream_t st, const StepPlan& plan = StepPlan()) {
    uint32_t f = phase_a_flags(p);
    if (f == 0) return CADL_OK;
    const bool need_p = f & (FA_SI | FA_PSUM | FA_EV | FA_TR);
    const bool need_g = f & (FA_SI | FA_RP | FA_EV | FA_TR);
    if ((need_p && !pred) || (need_g && !gt)) return CADL_ERR_NULL;
    PhaseAArgs a{};
    a.pred = pred; a.gt = gt; a.mask = mask;
    a.B = B; a.HW = H * W;
    a.blocks_per_img = ws.L.a_blocks_per_img;
    if (plan.a_blocks_per_img > 0 && plan.a_blocks_per_img < a.blocks_per_img) a.blocks_per_img = plan.a_blocks_per_img;
    a.vec_ok = ((H * W) % 4 == 0) && (!pred || aligned(pred, 16)) && (!gt || aligned(gt, 16)) &&
               (!mask || aligned(mask, 4));
    a.eps_si = p.eps_si; a.eps_rp = p.eps_reproj; a.min_d = p.min_depth; a.max_d = p.max_depth;
    {
        // guard bands of the delta thresholds in the log2 domain: two lg2.approx (about one ulp of the result each:
        // < 6.8e-7 for depths in [0.05, 20], growing with |log2 depth|) and the fp32 rounding of the difference
        float big = fabsf(log2f(p.max_depth > 0.f ? p.max_depth : 1.f));
        const float lo = fabsf(log2f(p.min_depth > 1e-30f ? p.min_depth : 1e-30f));
        if (lo > big) big = lo;
        if (!(big > 4.f)) big = 4.f;
        const float bw = 3.8e-6f * (big / 4.f);
        for (int k = 0; k < 3; ++k) {
            const float t = 0.32192809488736235f * (float)(k + 1), l = t - bw, h = t + bw;
            uint32_t ul, uh;
            memcpy(&ul, &l, 4); memcpy(&uh, &h, 4);
            a.near_lo[k] = ul; a.near_span[k] = uh - ul;
        }
    }
    a.hdr = ws.hdr(); a.stats = ws.stats(); a.img_psum = ws.img_psum(); a.a_part = ws.a_part();
    dim3 grid(a.blocks_per_img, B);
    return cuda_rc(dispatch_a(f, a, grid, st));
}

// Validation + the argument block shared by every phase-B kernel.  Returns CADL_OK with *nothing_to_do set when only
// metrics (or nothing) were asked for.
int fill_b_args(PhaseBArgs& a, const float* pred, const float* gt, const float* rgb, const float* K, const uint8_t* mask,
                int B, int H, int W, const cadl_params& p, float* grad, cadl_results* results, const Ws& ws,
                bool* nothing_to_do) {
    *nothing_to_do = false;
    if (!results) return CADL_ERR_NULL;
    const uint32_t t = p.terms & CADL_TERM_ALL;
    if (t == 0) {
        *nothing_to_do = true;
        return CADL_OK;
    }
    if (!pred) return CADL_ERR_NULL;
    if ((t & (CADL_TERM_SI | CADL_TERM_GRAD | CADL_TERM_REPROJ)) && !gt) return CADL_ERR_NULL;
    if ((t & CADL_TERM_SMOOTH) && !rgb) return CADL_ERR_NULL;
    if ((t & CADL_TERM_REPROJ) && !K) return CADL_ERR_NULL;
    if (t & CADL_TERM_GRAD) {
        if (p.num_scales < 1 || p.num_scales > CADL_MAX_SCALES) return CADL_ERR_UNSUPPORTED;
        // avg_pool2d needs at least one output cell at the coarsest scale (torch raises otherwise)
        if ((H >> (p.num_scales - 1)) < 1 || (W >> (p.num_scales - 1)) < 1) return CADL_ERR_SHAPE;
    }
    a = PhaseBArgs{};
    a.pred = pred; a.gt = gt; a.rgb = rgb; a.K = K; a.mask = mask; a.grad = grad;
    a.B = B; a.H = H; a.W = W;
    a.tiles_x = (W + TW - 1) / TW; a.tiles_y = (H + TH - 1) / TH;
    a.vec_ok = (W % 4 == 0) && aligned(pred, 16) && (!gt || aligned(gt, 16)) && (!rgb || aligned(rgb, 16)) &&
               (!grad || aligned(grad, 16)) && (!mask || aligned(mask, 4));
    a.num_scales = p.num_scales; a.k_batched = p.k_batched;
    a.global_B = p.global_B > 0 ? p.global_B : B;
    a.terms = t; a.metrics = p.metrics;
    a.w_si = p.w_si; a.w_grad = p.w_grad; a.w_smooth = p.w_smooth; a.w_rp = p.w_reproj;
    a.lambda = p.si_lambda;
    a.eps_si = p.eps_si; a.eps_grad = p.eps_grad; a.eps_smooth = p.eps_smooth; a.eps_rp = p.eps_reproj;
    a.upstream = p.upstream;
    a.stats = ws.stats(); a.img_psum = ws.img_psum(); a.b_part = ws.b_part();
    a.hdr = ws.hdr(); a.img_sm = ws.img_sm(); a.img_off = ws.img_off();
    a.results = results;
    for (int s = 0; s < 4; ++s) {
        const int Hs = H >> s, Ws = W >> s;
        const double nx = (double)a.global_B * Hs * (Ws - 1), ny = (double)a.global_B * (Hs - 1) * Ws;
        a.inv_nx[s] = nx > 0.0 ? (float)(1.0 / nx) : 0.f;
        a.inv_ny[s] = ny > 0.0 ? (float)(1.0 / ny) : 0.f;
    }
    a.sm_nx = a.inv_nx[0];
    a.sm_ny = a.inv_ny[0];
    return CADL_OK;
}

// fast path: aligned shapes (every BASELINE configuration); same values, ~4x fewer instructions
bool fast_path_ok(const PhaseBArgs& a, const cadl_params& p) {
    const uint32_t t = a.terms;
    return a.vec_ok && (a.H % 8 == 0) && (a.W % 8 == 0) && (!(t & CADL_TERM_GRAD) || p.num_scales == 4) &&
           (!((t & CADL_TERM_SI) && (t & CADL_TERM_GRAD)) || p.eps_si == p.eps_grad) && p.eps_si > 0.f &&
           p.eps_grad > 0.f && p.eps_si <= 1000.f && p.eps_grad <= 1000.f && !g_force_generic;
}
bool stream_path_ok(const PhaseBArgs& a, const cadl_params& p, const Ws& ws) {
    return fast_path_ok(a, p) && (a.terms & CADL_TERM_GRAD) && ws.has_pyr() && !g_force_tile;
}

int run_grad(const float* pred, const float* gt, const float* rgb, const float* K, const uint8_t* mask,
             int B, int H, int W, const cadl_params& p, float* grad, cadl_results* results, const Ws& ws,
             cudaStream_t st, const StepPlan& plan = StepPlan()) {
    PhaseBArgs a;
    bool nothing = false;
    int rc0 = fill_b_args(a, pred, gt, rgb, K, mask, B, H, W, p, grad, results, ws, &nothing);
    if (rc0) return rc0;
    if (nothing) {
        if (p.metrics) {
            metrics_finalize_kernel<<<1, 32, 0, st>>>(ws.stats(), p.metrics, results);
            return cuda_rc(cudaGetLastError());
        }
        return CADL_OK;
    }
    const uint32_t t = a.terms;

    cudaError_t e = cudaSuccess;
    bool offset_done = false;
    if ((t & (CADL_TERM_GRAD | CADL_TERM_SMOOTH)) == 0) {
        a.b_rows = kPointBlocks;
        const bool pfast = a.vec_ok && p.eps_si > 0.f && p.eps_si <= 1000.f && !g_force_generic;
        if (pfast) {
            switch (t) {
                case CADL_TERM_SI: e = launch_point_fast<FB_SI>(a, st); break;
                case CADL_TERM_REPROJ: e = launch_point_fast<FB_RP>(a, st); break;
                case CADL_TERM_SI | CADL_TERM_REPROJ: e = launch_point_fast<FB_SI | FB_RP>(a, st); break;
                default: return CADL_ERR_UNSUPPORTED;
            }
            return cuda_rc(e);
        }
        switch (t) {
            case CADL_TERM_SI: e = launch_point<FB_SI>(a, st); break;
            case CADL_TERM_REPROJ: e = launch_point<FB_RP>(a, st); break;
            case CADL_TERM_SI | CADL_TERM_REPROJ: e = launch_point<FB_SI | FB_RP>(a, st); break;
            default: return CADL_ERR_UNSUPPORTED;
        }
        return cuda_rc(e);
    }
    const bool fast = fast_path_ok(a, p);
    if (fast) {
        a.tiles_x = (W + FTW - 1) / FTW; a.tiles_y = (H + FTH - 1) / FTH;
        a.b_rows = a.tiles_x * a.tiles_y * B;
        bool stream = stream_path_ok(a, p, ws);
        if (stream) {
            PhaseBArgs as = a;             // (launch_stream re-purposes the tile fields)
            switch (t) {
                case CADL_TERM_ALL: e = launch_stream<15>(as, ws, st, &offset_done, plan); break;
                case CADL_TERM_SI | CADL_TERM_GRAD | CADL_TERM_SMOOTH: e = launch_stream<7>(as, ws, st, &offset_done, plan); break;
                case CADL_TERM_GRAD: e = launch_stream<FB_GRAD>(as, ws, st, &offset_done, plan); break;
                default: e = cudaErrorNotSupported; break;
            }
            if (e == cudaErrorNotSupported || e == cudaErrorCooperativeLaunchTooLarge) {
                cudaGetLastError();
                stream = false;            // e.g. unequal eps, a device without cooperative launch: the tile kernel
            }
        }
        if (!stream) {
            switch (t) {
                case CADL_TERM_ALL: e = launch_fast<15>(a, st); break;
                case CADL_TERM_SI | CADL_TERM_GRAD | CADL_TERM_SMOOTH: e = launch_fast<7>(a, st); break;
                case CADL_TERM_GRAD: e = launch_fast<FB_GRAD>(a, st); break;
                case CADL_TERM_SMOOTH: e = launch_fast<FB_SMOOTH>(a, st); break;
                default: return CADL_ERR_UNSUPPORTED;
            }
        }
    } else {
    a.b_rows = ws.L.b_tiles;
    switch (t) {
        case CADL_TERM_ALL: e = launch_tile<15>(a, st); break;                           // forwardWithIntrinsics
        case CADL_TERM_SI | CADL_TERM_GRAD | CADL_TERM_SMOOTH: e = launch_tile<7>(a, st); break;  // forward
        case CADL_TERM_GRAD: e = launch_tile<FB_GRAD>(a, st); break;
        case CADL_TERM_SMOOTH: e = launch_tile<FB_SMOOTH>(a, st); break;
        default: return CADL_ERR_UNSUPPORTED;
    }
    }
    if (e != cudaSuccess) return cuda_rc(e);
    if ((t & CADL_TERM_SMOOTH) && grad && !offset_done) {
        const int HW = H * W;
        const int vec = (HW % 4 == 0) && aligned(grad, 16);
        int bx = (HW / 4 + 255) / 256;
        int cap = (kGridCap + B - 1) / B;
        if (bx > cap) bx = cap;
        if (bx < 1) bx = 1;
        smooth_offset_kernel<<<dim3(bx, B), 256, 0, st>>>(grad, ws.img_off(), HW, vec);
        kt_mark(st, "smooth_offset_kernel");
        e = cudaGetLastError();
    }
    return cuda_rc(e);
}

}  // namespace

extern "C" {

void cadl_default_params(cadl_params* p) {
    memset(p, 0, sizeof(*p));
    p->terms = CADL_TERM_ALL;
    p->metrics = 0;
    p->w_si = 1.0f; p->w_grad = 0.1f; p->w_smooth = 0.001f; p->w_reproj = 0.01f;   // depth_loss.h:368-371
    p->si_lambda = 0.5f;                                                            // :22
    p->eps_si = p->eps_grad = p->eps_smooth = p->eps_reproj = 1e-6f;                // :22,84,180,257
    p->num_scales = 4;                                                              // :84
    p->k_batched = 1;
    p->min_depth = 0.1f; p->max_depth = 10.0f;                                      // depth_metrics.h:44-45
    p->upstream = 1.0f;
    p->global_B = 0;
}

int cadl_version(void) { return CADL_VERSION; }
#ifdef CADL_DEBUG
int cadl_debug_kernel_times(int enable, float* ms_out, const char** names_out, int cap) {
    g_kt.on = enable != 0;
    g_kt.n = 0;
    int n = g_kt.n_last < cap ? g_kt.n_last : cap;
    for (int i = 0; i < n; ++i) {
        if (ms_out) ms_out[i] = g_kt.ms_last[i];
        if (names_out) names_out[i] = g_kt.name_last[i];
    }
    return n;
}

void cadl_debug_force_generic(int on) {
    g_force_generic = on & 1;
    g_force_no_tma = (on >> 1) & 1;
    g_force_tile = (on >> 3) & 1;
    g_no_pdl = (on >> 4) & 1;
    g_no_overlap = (on >> 5) & 1;
    g_no_coop = (on >> 6) & 1;
    g_no_poolstats = (on >> 7) & 1;
}
#endif
size_t cadl_sizeof_params(void) { return sizeof(cadl_params); }
size_t cadl_sizeof_results(void) { return sizeof(cadl_results); }

const char* cadl_error_string(int code) {
    switch (code) {
        case CADL_OK: return "ok";
        case CADL_ERR_NULL: return "cadl: a required pointer is NULL";
        case CADL_ERR_SHAPE: return "cadl: B/H/W out of range for the requested terms";
        case CADL_ERR_WORKSPACE: return "cadl: workspace too small or not 256-byte aligned";
        case CADL_ERR_UNSUPPORTED: return "cadl: unsupported term combination or num_scales";
        case CADL_ERR_ALIGN: return "cadl: data pointer not 4-byte aligned";
        default: break;
    }
    if (code >= CADL_ERR_CUDA) {
        snprintf(g_err_detail, sizeof(g_err_detail), "cadl: CUDA error %d: %s", code - CADL_ERR_CUDA,
                 cudaGetErrorString((cudaError_t)(code - CADL_ERR_CUDA)));
        return g_err_detail;
    }
    return "cadl: unknown error";
}

size_t cadl_workspace_bytes(int B, int H, int W) {
    if (B < 1 || H < 1 || W < 1) return 0;
    WsLayout L = layout_for(B, H, W);
    return L.total + 256;
}

static Ws make_ws(void* workspace, int B, int H, int W) {
    Ws ws;
    ws.L = layout_for(B, H, W);
    ws.base = static_cast<char*>(workspace);
    return ws;
}

int cadl_workspace_init(void* workspace, size_t bytes, cadl_stream_t stream) {
    if (!workspace) return CADL_ERR_NULL;
    return cuda_rc(cudaMemsetAsync(workspace, 0, bytes, (cudaStream_t)stream));
}

size_t cadl_stats_offset(void) { return align_up(sizeof(WsHeader), 256); }
int cadl_stats_count(void) { return ST_COUNT; }

int cadl_stack_reduce(const float* pred, const float* gt, const uint8_t* mask, int B, int H, int W,
                      const cadl_params* params, void* workspace, size_t workspace_bytes, cadl_stream_t stream) {
    if (!params) return CADL_ERR_NULL;
    int rc = check_common(B, H, W, workspace, workspace_bytes);
    if (rc) return rc;
    Ws ws = make_ws(workspace, B, H, W);
    StepPlan plan;
    if (params->pyramid_prepared) { plan = concurrent_plan(ws, B, params->metrics != 0); plan.pyr_prelaunched = false; }   // (only the grid split matters here)
    return run_reduce(pred, gt, mask, B, H, W, *params, ws, (cudaStream_t)stream, plan);
}

int cadl_stack_prepare(const float* pred, const float* gt, int B, int H, int W, const cadl_params* params,
                       void* workspace, size_t workspace_bytes, cadl_stream_t stream) {
    if (!params || !pred || !gt) return CADL_ERR_NULL;
    int rc = check_common(B, H, W, workspace, workspace_bytes);
    if (rc) return rc;
    Ws ws = make_ws(workspace, B, H, W);
    const cadl_params& p = *params;
    // the conditions of the streaming path that can be known from pred/gt alone (stream_path_ok checks the rest)
    const bool ok = (p.terms & CADL_TERM_GRAD) && ws.has_pyr() && !g_force_tile && !g_force_generic && p.num_scales == 4 &&
                    (W % 4 == 0) && aligned(pred, 16) && aligned(gt, 16) && p.eps_grad > 0.f && p.eps_grad <= 1000.f &&
                    (!(p.terms & CADL_TERM_SI) || p.eps_si == p.eps_grad);
    if (!ok) return CADL_ERR_UNSUPPORTED;
    AuxStream* aux = aux_for_current_device();
    if (!aux) return CADL_ERR_UNSUPPORTED;
    PhaseBArgs a{};
    a.pred = pred; a.gt = gt; a.B = B; a.H = H; a.W = W;
    a.global_B = p.global_B > 0 ? p.global_B : B;
    a.w_grad = p.w_grad; a.eps_grad = p.eps_grad; a.upstream = p.upstream;
    a.b_part = ws.b_part();
    for (int s = 0; s < 4; ++s) {
        const int Hs = H >> s, Ws_ = W >> s;
        const double nx = (double)a.global_B * Hs * (Ws_ - 1), ny = (double)a.global_B * (Hs - 1) * Ws_;
        a.inv_nx[s] = nx > 0.0 ? (float)(1.0 / nx) : 0.f;
        a.inv_ny[s] = ny > 0.0 ? (float)(1.0 / ny) : 0.f;
    }
    return cuda_rc(prelaunch_pyramid(a, ws, (cudaStream_t)stream, aux, concurrent_plan(ws, B, params->metrics != 0)));
}

int cadl_stack_grad(const float* pred, const float* gt, const float* rgb, const float* K, const uint8_t* mask,
                    int B, int H, int W, const cadl_params* params, float* grad_pred, cadl_results* results,
                    void* workspace, size_t workspace_bytes, cadl_stream_t stream) {
    if (!params) return CADL_ERR_NULL;
    int rc = check_common(B, H, W, workspace, workspace_bytes);
    if (rc) return rc;
    Ws ws = make_ws(workspace, B, H, W);
    StepPlan plan;
    if (params->pyramid_prepared) {
        AuxStream* aux = aux_for_current_device();
        if (!aux) return CADL_ERR_UNSUPPORTED;
        plan = concurrent_plan(ws, B, params->metrics != 0);
        cudaError_t e = cudaStreamWaitEvent((cudaStream_t)stream, aux->join, 0);
        if (e != cudaSuccess) return cuda_rc(e);
        // (if the gradient part then takes another kernel -- e.g. an unaligned gradient buffer -- the prepared
        //  pyramid is simply not used)
    }
    return run_grad(pred, gt, rgb, K, mask, B, H, W, *params, grad_pred, results, ws, (cudaStream_t)stream, plan);
}

int cadl_stack_fwd_bwd(const float* pred, const float* gt, const float* rgb, const float* K, const uint8_t* mask,
                       int B, int H, int W, const cadl_params* params, float* grad_pred, cadl_results* results,
                       void* workspace, size_t workspace_bytes, cadl_stream_t stream) {
    if (!params) return CADL_ERR_NULL;
    int rc = check_common(B, H, W, workspace, workspace_bytes);
    if (rc) return rc;
    Ws ws = make_ws(workspace, B, H, W);
    cudaStream_t st = (cudaStream_t)stream;
    kt_mark(st, "start");
    // Streaming path with a phase A: the pooled-pyramid kernels go FIRST, onto the auxiliary stream, as one
    // persistent CTA per SM; phase A follows on the caller's stream with a grid that leaves them room.
    StepPlan plan;
    AuxStream* aux = nullptr;
    // Streaming path with the SI and smoothness terms and no metric variant (the trainers' training step): NO phase A.
    // The pooled-sum kernel reads every pred/gt value anyway and produces SI n / sum d / sum d^2 and the per-image
    // sum(pred) on its way: pool + statistics -> coefficients -> gradient pass, three launches on the caller's stream
    // (config 3 without metrics: 108 us/step against 120 with phase A beside the pyramid kernels).  With metric variants
    // phase A has to run anyway and carries the loss statistics at no extra cost, so that step keeps the layout below
    // (measured: a metrics-only phase A beside the new chain 130-150 us/step depending on how the two streams' CTAs
    // interleave, behind the gradient pass with programmatic serialization 153, the layout below 128).
    if (!g_no_overlap && !g_no_poolstats && !kt_on() && results && params->metrics == 0) {
        PhaseBArgs a;
        bool nothing = false;
        const uint32_t t = params->terms & CADL_TERM_ALL;
        if ((t == CADL_TERM_ALL || t == (CADL_TERM_SI | CADL_TERM_GRAD | CADL_TERM_SMOOTH)) &&
            fill_b_args(a, pred, gt, rgb, K, mask, B, H, W, *params, grad_pred, results, ws, &nothing) == CADL_OK && !nothing &&
            stream_path_ok(a, *params, ws) && params->eps_si == params->eps_grad &&
            (!(t & CADL_TERM_REPROJ) || params->eps_reproj == params->eps_grad) && (!mask || aligned(mask, 8))) {
            plan.pool_stats = PS_SI | PS_PSUM;
            plan.pyr_grid = CADL_PLAN2_PYR_N * num_sms_cached();
            if (plan.pyr_grid > ws.L.pyr_blocks) plan.pyr_grid = ws.L.pyr_blocks;
            return run_grad(pred, gt, rgb, K, mask, B, H, W, *params, grad_pred, results, ws, st, plan);
        }
    }
    if (!g_no_overlap && !kt_on() && phase_a_flags(*params) != 0 && results) {
        PhaseBArgs a;
        bool nothing = false;
        if (fill_b_args(a, pred, gt, rgb, K, mask, B, H, W, *params, grad_pred, results, ws, &nothing) == CADL_OK && !nothing &&
            stream_path_ok(a, *params, ws) && (aux = aux_for_current_device()) != nullptr) {
            plan = concurrent_plan(ws, B, params->metrics != 0);
            plan.pyr_prelaunched = false;
            cudaError_t e = prelaunch_pyramid(a, ws, st, aux, plan);
            if (e != cudaSuccess) return cuda_rc(e);
            plan.pyr_prelaunched = true;
        }
    }
    // Reprojection alone, no metrics (BASELINE config 2): the gradient kernel counts the valid pixels itself
    if ((params->terms & CADL_TERM_ALL) == CADL_TERM_REPROJ && params->metrics == 0 && !g_force_generic && !g_no_coop &&
        !kt_on() && results) {
        PhaseBArgs a;
        bool nothing = false;
        if (fill_b_args(a, pred, gt, rgb, K, mask, B, H, W, *params, grad_pred, results, ws, &nothing) == CADL_OK && !nothing &&
            a.vec_ok) {
            cudaError_t e = launch_point_count(a, st);
            if (e == cudaSuccess) return CADL_OK;
            if (e != cudaErrorNotSupported && e != cudaErrorCooperativeLaunchTooLarge) return cuda_rc(e);
            cudaGetLastError();            // not co-resident on this device / this shape: two launches
        }
    }
    // (Measured, not adopted: the metrics-only reduce pass BESIDE the gradient pass on the auxiliary stream -- 168 vs
    //  158 us/step: the cooperative gradient kernel fills every SM's register file, so the two cannot share the SMs.)
    rc = run_reduce(pred, gt, mask, B, H, W, *params, ws, st, plan);
    if (rc) return rc;
    kt_mark(st, "phase_a_kernel");
    if (plan.pyr_prelaunched) {
        cudaError_t e = cudaStreamWaitEvent(st, aux->join, 0);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    rc = run_grad(pred, gt, rgb, K, mask, B, H, W, *params, grad_pred, results, ws, st, plan);
    kt_mark((cudaStream_t)stream, "end");
    kt_finish();
    return rc;
}

int cadl_si_fwd_bwd(const float* pred, const float* gt, const uint8_t* mask, int B, int H, int W, float lambda,
                    float eps, float upstream, float* grad_pred, cadl_results* results, void* workspace,
                    size_t workspace_bytes, cadl_stream_t stream) {
    cadl_params p;
    cadl_default_params(&p);
    p.terms = CADL_TERM_SI; p.w_si = 1.0f; p.si_lambda = lambda; p.eps_si = eps; p.upstream = upstream;
    return cadl_stack_fwd_bwd(pred, gt, nullptr, nullptr, mask, B, H, W, &p, grad_pred, results, workspace,
                              workspace_bytes, stream);
}

int cadl_gradmatch_fwd_bwd(const float* pred, const float* gt, int B, int H, int W, int num_scales, float eps,
                           float upstream, float* grad_pred, cadl_results* results, void* workspace,
                           size_t workspace_bytes, cadl_stream_t stream) {
    cadl_params p;
    cadl_default_params(&p);
    p.terms = CADL_TERM_GRAD; p.w_grad = 1.0f; p.num_scales = num_scales; p.eps_grad = eps; p.upstream = upstream;
    return cadl_stack_fwd_bwd(pred, gt, nullptr, nullptr, nullptr, B, H, W, &p, grad_pred, results, workspace,
                              workspace_bytes, stream);
}

int cadl_smooth_fwd_bwd(const float* pred, const float* rgb, int B, int H, int W, float eps, float upstream,
                        float* grad_pred, cadl_results* results, void* workspace, size_t workspace_bytes,
                        cadl_stream_t stream) {
    cadl_params p;
    cadl_default_params(&p);
    p.terms = CADL_TERM_SMOOTH; p.w_smooth = 1.0f; p.eps_smooth = eps; p.upstream = upstream;
    return cadl_stack_fwd_bwd(pred, nullptr, rgb, nullptr, nullptr, B, H, W, &p, grad_pred, results, workspace,
                              workspace_bytes, stream);
}

int cadl_reproj_fwd_bwd(const float* pred, const float* gt, const float* K, int k_batched, const uint8_t* mask,
                        int B, int H, int W, float eps, float upstream, float* grad_pred, cadl_results* results,
                        void* workspace, size_t workspace_bytes, cadl_stream_t stream) {
    cadl_params p;
    cadl_default_params(&p);
    p.terms = CADL_TERM_REPROJ; p.w_reproj = 1.0f; p.eps_reproj = eps; p.k_batched = k_batched;
    p.upstream = upstream;
    return cadl_stack_fwd_bwd(pred, gt, nullptr, K, mask, B, H, W, &p, grad_pred, results, workspace,
                              workspace_bytes, stream);
}

int cadl_scale_grad(const float* grad_in, const float* upstream_dev, float* grad_out, size_t n,
                    cadl_stream_t stream) {
    if (!grad_in || !upstream_dev || !grad_out) return CADL_ERR_NULL;
    if (n == 0) return CADL_OK;
    const int vec = aligned(grad_in, 16) && aligned(grad_out, 16);
    size_t blocks = (n / 4 + 255) / 256;
    if (blocks > kGridCap) blocks = kGridCap;
    if (blocks < 1) blocks = 1;
    scale_grad_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(grad_in, upstream_dev, grad_out, n, vec);
    return cuda_rc(cudaGetLastError());
}

int cadl_metrics(const float* pred, const float* gt, const uint8_t* mask, size_t n, uint32_t which,
                 float min_depth, float max_depth, cadl_results* results, void* workspace,
                 size_t workspace_bytes, cadl_stream_t stream) {
    if (!pred || !gt || !results) return CADL_ERR_NULL;
    if (n == 0 || n > 0x7fffffffULL) return CADL_ERR_SHAPE;
    if ((which & (CADL_METRICS_EVAL | CADL_METRICS_TRAIN)) == 0) return CADL_ERR_UNSUPPORTED;
    // treated as one "image" of n values: H = 1, W = n
    cadl_params p;
    cadl_default_params(&p);
    p.terms = 0; p.metrics = which; p.min_depth = min_depth; p.max_depth = max_depth;
    return cadl_stack_fwd_bwd(pred, gt, nullptr, nullptr, mask, 1, 1, (int)n, &p, nullptr, results, workspace,
                              workspace_bytes, stream);
}

int cadl_selftest(int which, uint32_t lo_bits, uint32_t hi_bits, float param, unsigned long long* mismatches_dev,
                  cadl_stream_t stream) {
    if (!mismatches_dev) return CADL_ERR_NULL;
    if (hi_bits < lo_bits) return CADL_ERR_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    if (which == 0) selftest_log_kernel<<<kGridCap, 256, 0, st>>>(lo_bits, hi_bits, mismatches_dev);
    else if (which == 1) {
        if (!markstein_safe_host(param)) return CADL_ERR_UNSUPPORTED;
        selftest_div_kernel<<<kGridCap, 256, 0, st>>>(lo_bits, hi_bits, param, mismatches_dev);
    } else if (which == 2) {
        selftest_lg2_kernel<<<kGridCap, 256, 0, st>>>(lo_bits, hi_bits, param, mismatches_dev);
    } else if (which == 3) {
        selftest_div2_kernel<<<kGridCap, 256, 0, st>>>(lo_bits, hi_bits, param, mismatches_dev);
    } else return CADL_ERR_UNSUPPORTED;
    return cuda_rc(cudaGetLastError());
}

int cadl_rays_from_K(const float* K, int k_batched, const float* pose, int B, int H, int W, int layout,
                     float* out, cadl_stream_t stream) {
    if (!K || !out) return CADL_ERR_NULL;
    if (B < 1 || H < 1 || W < 1 || (layout != 0 && layout != 1)) return CADL_ERR_SHAPE;
    return cuda_rc(launch_rays(K, k_batched, pose, B, H, W, layout, out, (cudaStream_t)stream));
}

int cadl_batch_prep(const float* rgb_in, const float* depth_in, const float* K_in, int B, int h, int w, int H, int W,
                    float* rgb_out, float* depth_out, float* K_out, cadl_stream_t stream) {
    if (!rgb_in || !depth_in || !K_in || !rgb_out || !depth_out || !K_out) return CADL_ERR_NULL;
    if (B < 1 || h < 1 || w < 1 || H < 1 || W < 1 || H > 65535 || B > 65535) return CADL_ERR_SHAPE;
    PrepArgs a{rgb_in, depth_in, K_in, rgb_out, depth_out, K_out, B, h, w, H, W, nullptr};
    return cuda_rc(launch_batch_prep(a, (cudaStream_t)stream));
}

int cadl_batch_augment(const float* rgb_in, const float* depth_in, const float* K_in, const float* aug_dev, int B, int h,
                       int w, int H, int W, float* rgb_out, float* depth_out, float* K_out, cadl_stream_t stream) {
    if (!rgb_in || !depth_in || !K_in || !rgb_out || !depth_out || !K_out || !aug_dev) return CADL_ERR_NULL;
    if (B < 1 || h < 1 || w < 1 || H < 1 || W < 1 || H > 65535 || B > 65535) return CADL_ERR_SHAPE;
    PrepArgs a{rgb_in, depth_in, K_in, rgb_out, depth_out, K_out, B, h, w, H, W, aug_dev};
    return cuda_rc(launch_batch_prep(a, (cudaStream_t)stream));
}

size_t cadl_p2p_inbox_bytes(int world) {
    if (world < 1 || world > kP2PMaxWorld) return 0;
    return align_up(sizeof(double) * 2 * (size_t)world * kP2PSlotDoubles + sizeof(int) * 4, 256);
}

int cadl_p2p_alloc(int world, void** inbox_dev, unsigned char handle_out[64]) {
    if (!inbox_dev || !handle_out) return CADL_ERR_NULL;
    const size_t bytes = cadl_p2p_inbox_bytes(world);
    if (!bytes) return CADL_ERR_SHAPE;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { if (p) cudaFree(p); return cuda_rc(e); }
    memcpy(handle_out, &h, 64);
    *inbox_dev = p;
    return CADL_OK;
}

int cadl_p2p_open(const unsigned char handle[64], void** inbox_dev) {
    if (!handle || !inbox_dev) return CADL_ERR_NULL;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    return cuda_rc(cudaIpcOpenMemHandle(inbox_dev, h, cudaIpcMemLazyEnablePeerAccess));
}

int cadl_p2p_close(void* inbox_dev, int own) {
    if (!inbox_dev) return CADL_ERR_NULL;
    return cuda_rc(own ? cudaFree(inbox_dev) : cudaIpcCloseMemHandle(inbox_dev));
}

int cadl_stats_exchange(void* workspace, void* const* inboxes_host, int rank, int world, unsigned long long epoch,
                        double timeout_s, cadl_stream_t stream) {
    if (!workspace || !inboxes_host) return CADL_ERR_NULL;
    if (world < 1 || world > kP2PMaxWorld || rank < 0 || rank >= world || epoch == 0) return CADL_ERR_SHAPE;
    P2PArgs a{};
    a.stats = reinterpret_cast<double*>(static_cast<char*>(workspace) + cadl_stats_offset());
    for (int r = 0; r < world; ++r) {
        if (!inboxes_host[r]) return CADL_ERR_NULL;
        a.inbox[r] = static_cast<double*>(inboxes_host[r]);
    }
    a.rank = rank; a.world = world; a.epoch = epoch;
    a.timeout_ns = (unsigned long long)((timeout_s > 0.0 ? timeout_s : 600.0) * 1e9);
    a.error = reinterpret_cast<int*>(a.inbox[rank] + 2 * (size_t)world * kP2PSlotDoubles);
    return cuda_rc(launch_pdl(stats_exchange_kernel, dim3(1), dim3(64), (cudaStream_t)stream, !g_no_pdl, a));
}

int cadl_p2p_error(void* own_inbox_dev, int world, int* error_host, int clear) {
    if (!own_inbox_dev || !error_host) return CADL_ERR_NULL;
    char* p = static_cast<char*>(own_inbox_dev) + sizeof(double) * 2 * (size_t)world * kP2PSlotDoubles;
    cudaError_t e = cudaMemcpy(error_host, p, sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && clear && *error_host) e = cudaMemset(p, 0, sizeof(int));
    return cuda_rc(e);
}

int cadl_accumulate(const float* values_dev, int n, double weight, double* acc_dev, cadl_stream_t stream) {
    if (!values_dev || !acc_dev) return CADL_ERR_NULL;
    if (n < 1) return CADL_ERR_SHAPE;
    accumulate_kernel<<<(n + 1 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(values_dev, n, weight, acc_dev);
    return cuda_rc(cudaGetLastError());
}

size_t cadl_clip_workspace_bytes(void) { return 256 + sizeof(double) * kGridCap; }

int cadl_clip_grad_norm(float* const* grad_ptrs_dev, const long long* sizes_dev, const long long* chunk_prefix_dev,
                        int count, long long total_chunks, float max_norm, float* out2_dev, void* workspace,
                        size_t workspace_bytes, int do_clip, cadl_stream_t stream) {
    if (!grad_ptrs_dev || !sizes_dev || !chunk_prefix_dev || !out2_dev || !workspace) return CADL_ERR_NULL;
    if (count < 1 || total_chunks < 1) return CADL_ERR_SHAPE;
    if (!aligned(workspace, 256) || workspace_bytes < cadl_clip_workspace_bytes()) return CADL_ERR_WORKSPACE;
    ClipArgs a{};
    a.ptrs = grad_ptrs_dev; a.sizes = sizes_dev; a.chunk_prefix = chunk_prefix_dev; a.count = count;
    a.max_norm = max_norm; a.out = out2_dev;
    a.hdr = reinterpret_cast<WsHeader*>(workspace);
    a.part = reinterpret_cast<double*>(static_cast<char*>(workspace) + 256);
    int grid = kGridCap;
    if ((long long)grid > total_chunks) grid = (int)total_chunks;
    a.part_rows = grid;
    cudaStream_t st = (cudaStream_t)stream;
    gradnorm_kernel<<<grid, 256, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || !do_clip) return cuda_rc(e);
    gradscale_kernel<<<grid, 256, 0, st>>>(a);
    return cuda_rc(cudaGetLastError());
}

int cadl_photometric_fwd_bwd(const float* pred, const float* K, int k_batched, const float* T,
                             const float* source, const float* target, int B, int H, int W, float eps,
                             float upstream, float* grad_pred, cadl_results* results, void* workspace,
                             size_t workspace_bytes, cadl_stream_t stream) {
    if (!pred || !K || !T || !source || !target || !results) return CADL_ERR_NULL;
    int rc = check_common(B, H, W, workspace, workspace_bytes);
    if (rc) return rc;
    if (W % 4 != 0 || !aligned(pred, 16) || !aligned(target, 16) || (grad_pred && !aligned(grad_pred, 16)))
        return CADL_ERR_UNSUPPORTED;              // 128-bit rows (every BASELINE shape)
    Ws ws = make_ws(workspace, B, H, W);
    return cuda_rc(launch_photometric(pred, K, k_batched, T, source, target, B, H, W, eps, upstream, grad_pred,
                                      results, ws.hdr(), ws.b_part(), kPointBlocks, ws.img_off(),
                                      (cudaStream_t)stream));
}

}  // extern "C"
