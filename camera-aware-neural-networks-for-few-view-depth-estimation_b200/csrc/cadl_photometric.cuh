// cadl photometric reprojection -- BUILDER EXTENSION, not in the reference
// (ReprojectionLoss::forwardPhotometric is a stub returning zeros: depth_loss.h:343-351).
// Formulas: documents/algorithms_and_theory.md:22-25,62-77; sampling convention:
// grid_sample bilinear / zeros / align_corners=false as used by src/layers/pcl_layer.h:104-108.
//
//   X_t = d K^-1 [u v 1]^T ;  X_s = R X_t + t ;  (u', v') = (fx Xs/Zs + cx, fy Ys/Zs + cy)
//   residual = mean_c | bilinear(source, u', v') - target |  over pixels with Zs > eps and (u',v') inside
//   loss = mean over those pixels;   backward to d through the bilinear weights (explicit).
//
// Gather-bound: reads 4 B/px depth + 12 B/px target + the source image once (12 B/px through L2), writes 4 B/px:
// 32 algorithmic bytes per pixel.
#pragma once
#include "cadl_common.cuh"
#include "cadl_phase_a.cuh"
#include "cadl_phase_b.cuh"   // scale_grad_kernel

namespace cadl {

struct PhotoArgs {
    const float* pred; const float* K; const float* T; const float* src; const float* tgt;
    float* grad; cadl_results* results;
    int B, H, W, k_batched, rows;
    float eps, upstream;
    WsHeader* hdr; double* part; float* scale_out;
};

// One thread = 4 adjacent target pixels (128-bit loads of depth and of the three target channels, 128-bit gradient
// store); the per-image camera (K, R|t) is read once per row segment.  The source image is gathered through L1/L2:
// its footprint under a target tile depends on depth and pose (a 10-degree tilt moves the sample point by ~90 pixels
// at 480x640), so a fixed shared-memory halo cannot hold it; one image's three channels (3.7 MB) stay L2-resident
// while the image is being processed, and neighbouring threads hit the same 128-byte lines.
// Sampling follows ATen's grid_sampler (bilinear, zeros padding, align_corners = false) operation for operation --
// pixel -> normalised grid -> source index -- so that it can be checked against F.grid_sample in fp32.
__global__ void __launch_bounds__(256) photometric_kernel(const PhotoArgs a) {
    __shared__ double s_d[8];
    __shared__ float s_sum[8];
    __shared__ unsigned s_cnt[8];
    __shared__ int s_last;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = a.H, W = a.W;
    const int hw = H * W;
    const int W4 = W >> 2;                               // W % 4 == 0 on this path (host-checked)
    float acc = 0.f;
    unsigned cnt = 0;
    const long long quads = (long long)a.B * H * W4;
    const float Wf = (float)W, Hf = (float)H;
    for (long long q = (long long)blockIdx.x * blockDim.x + tid; q < quads; q += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(q / ((long long)H * W4));
        const int rem4 = (int)(q - (long long)b * H * W4);
        const int v = rem4 / W4, u0 = (rem4 - v * W4) << 2;
        const float* Kb = a.K + (a.k_batched ? (size_t)b * 9 : 0);
        const float fx = __ldg(Kb), cx = __ldg(Kb + 2), fy = __ldg(Kb + 4), cy = __ldg(Kb + 5);
        const float* Tb = a.T + (size_t)b * 16;
        const float r00 = __ldg(Tb), r01 = __ldg(Tb + 1), r02 = __ldg(Tb + 2), t0 = __ldg(Tb + 3);
        const float r10 = __ldg(Tb + 4), r11 = __ldg(Tb + 5), r12 = __ldg(Tb + 6), t1 = __ldg(Tb + 7);
        const float r20 = __ldg(Tb + 8), r21 = __ldg(Tb + 9), r22 = __ldg(Tb + 10), t2 = __ldg(Tb + 11);
        const size_t pix = (size_t)b * hw + (size_t)v * W + u0;
        const float4 d4 = __ldg(reinterpret_cast<const float4*>(a.pred + pix));
        const float* sb = a.src + (size_t)b * 3 * hw;
        const float* tb = a.tgt + (size_t)b * 3 * hw + (size_t)v * W + u0;
        const float4 tg0 = ldg_stream(reinterpret_cast<const float4*>(tb));
        const float4 tg1 = ldg_stream(reinterpret_cast<const float4*>(tb + hw));
        const float4 tg2 = ldg_stream(reinterpret_cast<const float4*>(tb + 2 * hw));
        const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
        const float tg[3][4] = {{tg0.x, tg0.y, tg0.z, tg0.w}, {tg1.x, tg1.y, tg1.z, tg1.w}, {tg2.x, tg2.y, tg2.z, tg2.w}};
        const float yh = ((float)v - cy) / fy;
        float gout[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float d = dd[k];
            const float xh = ((float)(u0 + k) - cx) / fx;
            // X_t = d (xh, yh, 1);  X_s = R X_t + t
            const float X = xh * d, Y = yh * d;
            const float Xs = r00 * X + r01 * Y + r02 * d + t0;
            const float Ys = r10 * X + r11 * Y + r12 * d + t1;
            const float Zs = r20 * X + r21 * Y + r22 * d + t2;
            float g = 0.f;
            if (Zs > a.eps) {
                const float us = fx * Xs / Zs + cx, vs = fy * Ys / Zs + cy;
                if (us >= 0.f && us <= Wf - 1.f && vs >= 0.f && vs <= Hf - 1.f) {
                    // ATen: grid = (2 u + 1) / W - 1;  ix = ((grid + 1) * W - 1) / 2
                    const float gx = (2.0f * us + 1.0f) / Wf - 1.0f, gy = (2.0f * vs + 1.0f) / Hf - 1.0f;
                    const float ix = ((gx + 1.f) * Wf - 1.f) * 0.5f, iy = ((gy + 1.f) * Hf - 1.f) * 0.5f;
                    const float fu = floorf(ix), fv = floorf(iy);
                    const int iu = (int)fu, iv = (int)fv;
                    const float tx = ix - fu, ty = iy - fv;
                    // corner weights as ATen forms them: (ix_se - ix) * (iy_se - iy) ...
                    const float wnw = (1.f - tx) * (1.f - ty), wne = tx * (1.f - ty), wsw = (1.f - tx) * ty, wse = tx * ty;
                    const bool w_ok = iu >= 0 && iu < W, e_ok = iu + 1 >= 0 && iu + 1 < W;
                    const bool n_ok = iv >= 0 && iv < H, s_ok = iv + 1 >= 0 && iv + 1 < H;
                    float resid = 0.f, dwdu = 0.f, dwdv = 0.f;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float* sc = sb + (size_t)c * hw;
                        const float nw = (w_ok && n_ok) ? __ldg(sc + (size_t)iv * W + iu) : 0.f;
                        const float ne = (e_ok && n_ok) ? __ldg(sc + (size_t)iv * W + iu + 1) : 0.f;
                        const float sw = (w_ok && s_ok) ? __ldg(sc + (size_t)(iv + 1) * W + iu) : 0.f;
                        const float se = (e_ok && s_ok) ? __ldg(sc + (size_t)(iv + 1) * W + iu + 1) : 0.f;
                        const float wv = nw * wnw + ne * wne + sw * wsw + se * wse;
                        const float r = wv - tg[c][k];
                        resid += fabsf(r);
                        const float sg = sgnf(r);
                        dwdu += sg * ((ne - nw) * (1.f - ty) + (se - sw) * ty);
                        dwdv += sg * ((sw - nw) * (1.f - tx) + (se - ne) * tx);
                    }
                    acc += resid * (1.0f / 3.0f);
                    cnt += 1u;
                    // d us / d d, d vs / d d  with  d X_s / d d = R (xh, yh, 1)
                    const float A0 = r00 * xh + r01 * yh + r02, A1 = r10 * xh + r11 * yh + r12, A2 = r20 * xh + r21 * yh + r22;
                    const float iz = 1.0f / Zs;
                    const float dus = fx * (A0 * Zs - Xs * A2) * iz * iz;
                    const float dvs = fy * (A1 * Zs - Ys * A2) * iz * iz;
                    g = (dwdu * dus + dwdv * dvs) * (1.0f / 3.0f);
                }
            }
            gout[k] = g;       // un-normalised; scaled by upstream / n afterwards
        }
        if (a.grad) *reinterpret_cast<float4*>(a.grad + pix) = make_float4(gout[0], gout[1], gout[2], gout[3]);
    }
    acc = warp_sum(acc);
    cnt = warp_sum(cnt);
    if (lane == 0) { s_sum[warp] = acc; s_cnt[warp] = cnt; }
    __syncthreads();
    if (tid == 0) {
        double s = 0.0, c = 0.0;
        for (int w = 0; w < 8; ++w) { s += (double)s_sum[w]; c += (double)s_cnt[w]; }
        a.part[2 * (size_t)blockIdx.x] = s;
        a.part[2 * (size_t)blockIdx.x + 1] = c;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned t = atomicAdd(&a.hdr->ticket_b, 1u);
        s_last = (t == gridDim.x - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const volatile double* part = a.part;
    double s = 0.0, c = 0.0;
    for (int i = tid; i < (int)gridDim.x; i += blockDim.x) { s += part[2 * i]; c += part[2 * i + 1]; }
    s = block_sum_double(s, s_d);
    c = block_sum_double(c, s_d);
    if (tid == 0) {
        cadl_results& r = *a.results;
        const double loss = c > 0.0 ? s / c : 0.0;
        r.d_reproj = loss; r.loss_reproj = (float)loss; r.n_reproj = (int64_t)c;
        r.d_total = loss; r.loss_total = (float)loss;
        *a.scale_out = c > 0.0 ? (float)((double)a.upstream / c) : 0.f;
        a.hdr->ticket_b = 0u;
    }
}

inline cudaError_t launch_photometric(const float* pred, const float* K, int k_batched, const float* T,
                                      const float* src, const float* tgt, int B, int H, int W, float eps,
                                      float upstream, float* grad, cadl_results* results, WsHeader* hdr,
                                      double* part, int rows, float* scale_out, cudaStream_t st) {
    PhotoArgs a{};
    a.pred = pred; a.K = K; a.T = T; a.src = src; a.tgt = tgt; a.grad = grad; a.results = results;
    a.B = B; a.H = H; a.W = W; a.k_batched = k_batched; a.rows = rows;
    a.eps = eps; a.upstream = upstream; a.hdr = hdr; a.part = part; a.scale_out = scale_out;
    long long total = (long long)B * H * W;
    if (W % 4 != 0) return cudaErrorNotSupported;
    int blocks = (int)((total / 4 + 255) / 256);
    if (blocks > rows) blocks = rows;
    photometric_kernel<<<blocks, 256, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || !grad) return e;
    const size_t n = (size_t)total;
    const int vec = (reinterpret_cast<uintptr_t>(grad) % 16) == 0;
    size_t sb = (n / 4 + 255) / 256;
    if (sb > (size_t)kGridCap) sb = kGridCap;
    if (sb < 1) sb = 1;
    // grad *= upstream / n  (same kernel as the autograd scale; scale_out lives in the workspace)
    scale_grad_kernel<<<(unsigned)sb, 256, 0, st>>>(grad, scale_out, grad, n, vec);
    return cudaGetLastError();
}

}  // namespace cadl
