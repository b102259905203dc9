// cadl phase B: argument block and the small device helpers shared by every gradient-pass kernel
// (split out of cadl_phase_b.cuh so that several translation units can include them).
#pragma once
#include "cadl_common.cuh"
#include "cadl_math.cuh"

namespace cadl {

constexpr int FB_SI = 1, FB_GRAD = 2, FB_SMOOTH = 4, FB_RP = 8;

struct PhaseBArgs {
    const float* pred;
    const float* gt;
    const float* rgb;
    const float* K;
    const uint8_t* mask;
    float* grad;  // may be null: forward only
    int B, H, W, tiles_x, tiles_y;
    int vec_ok;
    int num_scales, k_batched, global_B;
    uint32_t terms, metrics;
    float w_si, w_grad, w_smooth, w_rp, lambda;
    float eps_si, eps_grad, eps_smooth, eps_rp, upstream;
    const double* stats;
    const double* img_psum;
    double* b_part;
    int b_rows;  // rows of b_part written by this launch
    WsHeader* hdr;
    double* img_sm;
    float* img_off;
    cadl_results* results;
    // host-computed 1/N of the means (depth_loss.h:162-163, :230-231); 0 where a mean has no element
    float inv_nx[4], inv_ny[4], sm_nx, sm_ny;
    int use_tma;   // fast kernel: stage pred/gt with TMA box loads (tensor maps passed next to this struct)
};

// Scalars every CTA derives from the phase-A statistics.
struct Derived {
    float si_c1, si_c2;  // d(SI)/dd_i = c1*d_i + c2        
    float rp_inv_n;      // 1/n                              (a4)
    bool si_on, rp_on;   // n > 0 (depth_loss.h:53-55, :325-327)
};

__device__ __forceinline__ Derived derive(const PhaseBArgs& a) {
    Derived d;
    double n = a.stats[ST_SI_N], S = a.stats[ST_SI_S];
    d.si_on = n > 0.0;
    d.si_c1 = d.si_on ? (float)(2.0 / n) : 0.f;
    d.si_c2 = d.si_on ? (float)(-2.0 * (double)a.lambda * S / (n * n)) : 0.f;
    double nr = a.stats[ST_RP_N];
    d.rp_on = nr > 0.0;
    d.rp_inv_n = d.rp_on ? (float)(1.0 / nr) : 0.f;
    return d;
}

// 1/N for the two means of one gradient-matching scale (depth_loss.h:162-163); 0/0 -> NaN like
// torch's mean of an empty tensor.
__device__ __forceinline__ void scale_dims(const PhaseBArgs& a, int s, int& Hs, int& Ws, float& inv_nx,
                                           float& inv_ny) {
    Hs = a.H >> s;
    Ws = a.W >> s;
    double nx = (double)a.global_B * Hs * (Ws - 1);
    double ny = (double)a.global_B * (Hs - 1) * Ws;
    // no edges at all (Ws == 1 / Hs == 1): the reference's mean over an empty tensor is NaN for the LOSS
    // (finalize_results keeps that: 0/0) but contributes no gradient
    inv_nx = nx > 0.0 ? (float)(1.0 / nx) : 0.f;
    inv_ny = ny > 0.0 ? (float)(1.0 / ny) : 0.f;
}

// Image b's share of the smoothness loss and the constant its gradient is shifted by:
//   L_b = a_b * (sum_x / n_x + sum_y / n_y),   a_b = 1 / (mean(pred_b) + eps)          depth_loss.h:192-193, :230-231
//   d/dp_j of the mean-normalisation = -a_b * L_b / (H*W)
__device__ __forceinline__ void smooth_image_share(const PhaseBArgs& a, int img, double sx, double sy, double& Lb, float& off) {
    const double HW = (double)a.H * a.W;
    const double nx = (double)a.global_B * a.H * (a.W - 1);
    const double ny = (double)a.global_B * (a.H - 1) * a.W;
    const double mean = a.img_psum[img] / HW;
    const float ab = 1.0f / ((float)mean + a.eps_smooth);   // depth_loss.h:193
    Lb = (double)ab * (sx / nx + sy / ny);
    off = (float)((double)a.upstream * a.w_smooth * ab * Lb / HW);
}

__device__ __forceinline__ void load_K(const PhaseBArgs& a, int b, float& fx, float& fy, float& cx,
                                       float& cy) {
    const float* Kb = a.K + (a.k_batched ? (size_t)b * 9 : 0);
    fx = __ldg(Kb + 0);   // depth_loss.h:290-293
    cx = __ldg(Kb + 2);
    fy = __ldg(Kb + 4);
    cy = __ldg(Kb + 5);
}

}  // namespace cadl
