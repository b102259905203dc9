// cadl photometric reprojection -- BUILDER EXTENSION, not in the reference
// (ReprojectionLoss::forwardPhotometric is a stub returning zeros: depth_loss.h:343-351).
// Formulas: documents/algorithms_and_theory.md:22-25,62-77; sampling convention:
// grid_sample bilinear / zeros / align_corners=false as used by src/layers/pcl_layer.h:104-108.
//
//   X_t = d K^-1 [u v 1]^T ;  X_s = R X_t + t ;  (u', v') = (fx Xs/Zs + cx, fy Ys/Zs + cy)
//   residual = mean_c | bilinear(source, u', v') - target |  over pixels with Zs > eps and (u',v') inside
//   loss = mean over those pixels;   backward to d through the bilinear weights (explicit).
//
// Gather-bound: reads 4 B/px depth + 12 B/px target + ~4 source texels x 3 channels (L1/L2 resident for
// small motions), writes 4 B/px.
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

__global__ void __launch_bounds__(256) photometric_kernel(const PhotoArgs a) {
    __shared__ double s_d[8];
    __shared__ float s_sum[8];
    __shared__ unsigned s_cnt[8];
    __shared__ int s_last;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = a.H, W = a.W;
    const size_t hw = (size_t)H * W;
    float acc = 0.f;
    unsigned cnt = 0;
    const long long total = (long long)a.B * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / (long long)hw);
        const int rem = (int)(i - (long long)b * hw);
        const int v = rem / W, u = rem - v * W;
        const float* Kb = a.K + (a.k_batched ? (size_t)b * 9 : 0);
        const float fx = __ldg(Kb), cx = __ldg(Kb + 2), fy = __ldg(Kb + 4), cy = __ldg(Kb + 5);
        const float* Tb = a.T + (size_t)b * 16;
        const float d = __ldg(a.pred + i);
        const float xh = ((float)u - cx) / fx, yh = ((float)v - cy) / fy;
        // X_s = (R [xh yh 1]^T) d + t
        const float A0 = Tb[0] * xh + Tb[1] * yh + Tb[2];
        const float A1 = Tb[4] * xh + Tb[5] * yh + Tb[6];
        const float A2 = Tb[8] * xh + Tb[9] * yh + Tb[10];
        const float Xs = A0 * d + Tb[3], Ys = A1 * d + Tb[7], Zs = A2 * d + Tb[11];
        float g = 0.f;
        if (Zs > a.eps) {
            const float iz = 1.0f / Zs;
            const float us = fx * Xs * iz + cx, vs = fy * Ys * iz + cy;
            if (us >= 0.f && us <= (float)(W - 1) && vs >= 0.f && vs <= (float)(H - 1)) {
                const float fu = floorf(us), fv = floorf(vs);
                const int iu = (int)fu, iv = (int)fv;
                const float tx = us - fu, ty = vs - fv;
                const bool e_ok = iu + 1 < W, s_ok = iv + 1 < H;
                const float* sb = a.src + (size_t)b * 3 * hw;
                const float* tb = a.tgt + (size_t)b * 3 * hw;
                float resid = 0.f, dwdu = 0.f, dwdv = 0.f;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float* sc = sb + c * hw;
                    const float nw = __ldg(sc + (size_t)iv * W + iu);
                    const float ne = e_ok ? __ldg(sc + (size_t)iv * W + iu + 1) : 0.f;
                    const float sw = s_ok ? __ldg(sc + (size_t)(iv + 1) * W + iu) : 0.f;
                    const float se = (e_ok && s_ok) ? __ldg(sc + (size_t)(iv + 1) * W + iu + 1) : 0.f;
                    const float wv = nw * (1.f - tx) * (1.f - ty) + ne * tx * (1.f - ty) + sw * (1.f - tx) * ty + se * tx * ty;
                    const float r = wv - __ldg(tb + c * hw + rem);
                    resid += fabsf(r);
                    const float sg = sgnf(r);
                    dwdu += sg * ((ne - nw) * (1.f - ty) + (se - sw) * ty);
                    dwdv += sg * ((sw - nw) * (1.f - tx) + (se - ne) * tx);
                }
                acc += resid * (1.0f / 3.0f);
                cnt += 1u;
                const float dus = fx * (A0 * Zs - Xs * A2) * iz * iz;
                const float dvs = fy * (A1 * Zs - Ys * A2) * iz * iz;
                g = (dwdu * dus + dwdv * dvs) * (1.0f / 3.0f);
            }
        }
        if (a.grad) a.grad[i] = g;   // un-normalised; scaled by upstream/n afterwards
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
    int blocks = (int)((total + 255) / 256);
    if (blocks > rows) blocks = rows;
    photometric_kernel<<<blocks, 256, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || !grad) return e;
    const size_t n = (size_t)total;
    const int vec = (reinterpret_cast<uintptr_t>(grad) % 16) == 0;
    size_t sb = (n / 4 + 255) / 256;
    if (sb > 148 * 8) sb = 148 * 8;
    if (sb < 1) sb = 1;
    // grad *= upstream / n  (same kernel as the autograd scale; scale_out lives in the workspace)
    scale_grad_kernel<<<(unsigned)sb, 256, 0, st>>>(grad, scale_out, grad, n, vec);
    return cudaGetLastError();
}

}  // namespace cadl
