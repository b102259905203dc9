// Translation unit of the round-2 streaming gradient kernel (cadl_stream3.cuh): its instantiations, the TMA
// descriptors and the host-side launch geometry.  Built only for sm_100a.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>

#include "cadl_stream3.cuh"

namespace cadl {

namespace {
// ---- TMA descriptors (cuTensorMapEncodeTiled through the runtime's driver entry point: no libcuda link) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
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
// (W, H, planes) fp32 tensor, box = (kS3BoxW, 1, 1), zero fill outside
bool make_row_map(CUtensorMap* m, const float* base, int planes, int H, int W, int box_w = kS3BoxW, int box_d = 1) {
    EncodeTiledFn fn = encode_fn();
    if (!fn || !base) return false;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)box_w, 1, (cuuint32_t)box_d};
    cuuint32_t estr[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
struct RowMaps {
    CUtensorMap pred, gt, rgb, c1;
    const float *pp = nullptr, *pg = nullptr, *pi = nullptr, *pc = nullptr;
    int B = 0, H = 0, W = 0;
    bool ok = false;
};
thread_local RowMaps g_maps;   // re-encoded only when a pointer or the shape changes

struct Geometry {
    int per_sm = 0;      // resident CTAs per SM of this instantiation (0 = not yet queried on this device)
};
Geometry g_geo[16][16][2];   // [device][F][mask]

template <int F, bool M>
cudaError_t launch_one(const PhaseBArgs& a, Stream3Args& sa, cudaStream_t st) {
    constexpr bool SMOOTH = (F & FB_SMOOTH) != 0;
    constexpr int SLOT = 2 * kS3RowFloats + (SMOOTH ? kS3RgbFloats : 0) + kS3C1Floats;
    constexpr int wpc = kS3Threads / 32;
    constexpr size_t smem = (size_t)wpc * kS3Depth * SLOT * sizeof(float);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 16) return cudaErrorNotSupported;
    int sms = 0, coop = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop || sms < 1) return cudaErrorNotSupported;
    Geometry& geo = g_geo[dev][F][M ? 1 : 0];
    if (!geo.per_sm) {
        e = cudaFuncSetAttribute(stream3_kernel<F, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&geo.per_sm, stream3_kernel<F, M>, kS3Threads, smem);
        if (e != cudaSuccess) return e;
        if (geo.per_sm < 1) return cudaErrorNotSupported;
    }
    RowMaps& tm = g_maps;
    if (!(tm.ok && tm.pp == a.pred && tm.pg == a.gt && tm.pi == (SMOOTH ? a.rgb : nullptr) && tm.pc == sa.c1 && tm.B == a.B && tm.H == a.H && tm.W == a.W)) {
        tm.ok = make_row_map(&tm.pred, a.pred, a.B, a.H, a.W) && make_row_map(&tm.gt, a.gt, a.B, a.H, a.W);
        if (SMOOTH) tm.ok = tm.ok && make_row_map(&tm.rgb, a.rgb, 3 * a.B, a.H, a.W, kS3BoxW, kS3RgbDepth);
        else tm.rgb = tm.pred;
        tm.ok = tm.ok && make_row_map(&tm.c1, sa.c1, a.B, a.H / 2, a.W / 2, kS3C1BoxW);
        tm.pp = a.pred; tm.pg = a.gt; tm.pi = SMOOTH ? a.rgb : nullptr; tm.pc = sa.c1; tm.B = a.B; tm.H = a.H; tm.W = a.W;
    }
    if (!tm.ok) return cudaErrorNotSupported;

    // shares: (image, strip, row range); one per warp of the resident wave where the problem is large enough
    const int wave = geo.per_sm * sms * wpc;                 // co-resident warps
    const int cols = a.B * sa.nstrip;                        // image strips
    int kpi = wave / cols;
    if (kpi < 1) kpi = 1;
    const int cap = a.H / 4 > 0 ? a.H / 4 : 1;               // at least ~4 rows per share (one extra row is the halo)
    if (kpi > cap) kpi = cap;
    sa.kpi = kpi;
    const long long items = (long long)cols * kpi;
    sa.nwarps = items < wave ? (int)items : wave;
    const int grid = (sa.nwarps + wpc - 1) / wpc;
    void* args[] = {(void*)&a, (void*)&sa, (void*)&tm.pred, (void*)&tm.gt, (void*)&tm.rgb, (void*)&tm.c1};
    return cudaLaunchCooperativeKernel((void*)stream3_kernel<F, M>, dim3(grid), dim3(kS3Threads), args, smem, st);
}
}  // namespace

cudaError_t launch_stream3(int F, const PhaseBArgs& a, Stream3Args& sa, cudaStream_t st) {
    sa.inx0 = a.inv_nx[0] * 0.25f * a.w_grad * a.upstream;
    sa.iny0 = a.inv_ny[0] * 0.25f * a.w_grad * a.upstream;
    for (int s = 0; s < 4; ++s) {
        const int Hs = a.H >> s, Ws = a.W >> s;
        // (x / 0 of the reference's empty mean: the sum is 0 there, and 0 * inf is the same NaN)
        sa.rnx[s] = 1.0 / ((double)a.global_B * Hs * (Ws - 1));
        sa.rny[s] = 1.0 / ((double)a.global_B * (Hs - 1) * Ws);
    }
    sa.r_hw = 1.0 / ((double)a.H * a.W);
    const bool m = a.mask != nullptr;
    switch (F) {
        case 15: return m ? launch_one<15, true>(a, sa, st) : launch_one<15, false>(a, sa, st);
        case 7: return m ? launch_one<7, true>(a, sa, st) : launch_one<7, false>(a, sa, st);
        case FB_GRAD: return m ? launch_one<FB_GRAD, true>(a, sa, st) : launch_one<FB_GRAD, false>(a, sa, st);
        default: return cudaErrorNotSupported;
    }
}

// The smoothness weights are folded into the exponent (2^(k s + log2 c)); that needs a positive normal c.
bool stream3_fill_smooth(const PhaseBArgs& a, Stream3Args& sa) {
    const double cx = (double)a.sm_nx * a.w_smooth * a.upstream, cy = (double)a.sm_ny * a.w_smooth * a.upstream;
    if (!(cx > 1e-30 && cx < 1e30 && cy > 1e-30 && cy < 1e30)) return false;
    sa.lsnx = (float)log2(cx); sa.lsny = (float)log2(cy);
    // the kernel multiplies by 2^lsn (the rounded logarithm): undo exactly that factor
    sa.inv_snx = exp2(-(double)sa.lsnx); sa.inv_sny = exp2(-(double)sa.lsny);
    return true;
}

}  // namespace cadl

#ifdef CADL_S3_TRACE
extern "C" int cadl_debug_s3_trace(unsigned long long* out_host, int warps) {
    if (warps > 4096) warps = 4096;
    return (int)cudaMemcpyFromSymbol(out_host, cadl::g_s3_trace, sizeof(unsigned long long) * cadl::kS3TraceWords * (size_t)warps);
}
#endif
