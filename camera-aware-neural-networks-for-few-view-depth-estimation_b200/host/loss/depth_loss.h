// Drop-in replacement for the reference's src/loss/depth_loss.h.
//
// Same namespace, class names, constructor defaults, method signatures and return ranks
// (reference src/loss/depth_loss.h:20-479), so the trainers compile against it unchanged:
//   loss_fn_->forwardWithIntrinsics(pred, gt, rgb, K)    production_trainer.h:203
//   loss.backward()                                      production_trainer.h:206
// Behind the signatures there is no ATen op chain: each forward() is ONE autograd node that runs
// the fused sm_100a kernels of libcadl.so (include/cadl.h) -- forward and backward in the same
// pass -- on the current CUDA stream.  Differentiable w.r.t. pred only, like the trainers use it.
//
// Deliberate differences:
//   * CUDA tensors only.  A CPU tensor raises c10::Error; there is no fallback.
//   * ScaleInvariantLoss / ReprojectionLoss return a 0-dim tensor also when no pixel is valid (the
//     reference returns zeros(1) there, depth_loss.h:53-55,325-327, at the price of a host sync);
//     referenceEmptyRank(true) on either object opts into that sync and the rank-1 zeros.
#ifndef DEPTH_LOSS_H
#define DEPTH_LOSS_H

#include <torch/torch.h>

#include <map>
#include <string>

#include "../cadl_torch.h"

namespace camera_aware_depth {

/// Scale-invariant log loss, L = mean(d^2) - lambda * mean(d)^2, d = log pred - log gt over valid pixels.
/// reference: depth_loss.h:20-69
class ScaleInvariantLoss {
public:
    ScaleInvariantLoss(float lambda = 0.5f, float eps = 1e-6f) : lambda_(lambda), eps_(eps) {}

    torch::Tensor forward(torch::Tensor pred_depth, torch::Tensor gt_depth,
                          torch::optional<torch::Tensor> valid_mask = torch::nullopt) {
        using namespace cadl_detail;
        auto pred = as_input(pred_depth, "pred_depth", 1);
        auto gt = as_input(gt_depth, "gt_depth", 1);
        TORCH_CHECK(gt.sizes() == pred.sizes(), "cadl: gt_depth and pred_depth differ in shape");
        auto mask = as_mask(valid_mask, pred);
        cadl_params p;
        cadl_default_params(&p);
        p.terms = CADL_TERM_SI; p.w_si = 1.0f; p.si_lambda = lambda_; p.eps_si = eps_;
        auto loss = apply_fused(pred, gt, torch::Tensor(), torch::Tensor(), mask, p, offsetof(cadl_results, loss_si));
        return sync_empty_ ? with_reference_empty_rank(loss, offsetof(cadl_results, n_si), pred) : loss;
    }

    /// opt in to the reference's zeros(1) when no pixel is valid (depth_loss.h:53-55): costs one host sync per call
    ScaleInvariantLoss& referenceEmptyRank(bool on) { sync_empty_ = on; return *this; }

    float lambda() const { return lambda_; }
    float eps() const { return eps_; }

private:
    float lambda_;
    float eps_;
    bool sync_empty_ = false;
};

/// Multi-scale gradient matching in log depth (avg-pool pyramid, forward differences, L1).
/// reference: depth_loss.h:82-167.  Returns shape [1] like the reference (it starts from zeros(1), :99).
class GradientMatchingLoss {
public:
    GradientMatchingLoss(int num_scales = 4, float eps = 1e-6f) : num_scales_(num_scales), eps_(eps) {}

    torch::Tensor forward(torch::Tensor pred_depth, torch::Tensor gt_depth,
                          torch::optional<torch::Tensor> valid_mask = torch::nullopt) {
        using namespace cadl_detail;
        (void)valid_mask;   // accepted and unused, exactly like the reference (depth_loss.h:137)
        auto pred = as_input(pred_depth, "pred_depth", 1);
        auto gt = as_input(gt_depth, "gt_depth", 1);
        TORCH_CHECK(gt.sizes() == pred.sizes(), "cadl: gt_depth and pred_depth differ in shape");
        TORCH_CHECK(num_scales_ >= 1 && num_scales_ <= CADL_MAX_SCALES, "cadl: num_scales must be in [1,",
                    CADL_MAX_SCALES, "]");
        cadl_params p;
        cadl_default_params(&p);
        p.terms = CADL_TERM_GRAD; p.w_grad = 1.0f; p.num_scales = num_scales_; p.eps_grad = eps_;
        return apply_fused(pred, gt, torch::Tensor(), torch::Tensor(), torch::Tensor(), p,
                           offsetof(cadl_results, loss_grad)).reshape({1});
    }

    int num_scales() const { return num_scales_; }
    float eps() const { return eps_; }

private:
    int num_scales_;
    float eps_;
};

/// Edge-aware smoothness on mean-normalised depth.  reference: depth_loss.h:178-238
class SmoothnessLoss {
public:
    SmoothnessLoss(float eps = 1e-6f) : eps_(eps) {}

    torch::Tensor forward(torch::Tensor pred_depth, torch::Tensor image) {
        using namespace cadl_detail;
        auto pred = as_input(pred_depth, "pred_depth", 1);
        auto rgb = as_input(image, "image", 3);
        TORCH_CHECK(rgb.size(0) == pred.size(0) && rgb.size(2) == pred.size(2) && rgb.size(3) == pred.size(3),
                    "cadl: image and pred_depth differ in shape");
        TORCH_CHECK(rgb.device() == pred.device(), "cadl: image must be on the same device as pred");
        cadl_params p;
        cadl_default_params(&p);
        p.terms = CADL_TERM_SMOOTH; p.w_smooth = 1.0f; p.eps_smooth = eps_;
        return apply_fused(pred, torch::Tensor(), rgb, torch::Tensor(), torch::Tensor(), p,
                           offsetof(cadl_results, loss_smooth));
    }

    float eps() const { return eps_; }

private:
    float eps_;
};

/// 3-D point error between pred and gt depth back-projected with K.  reference: depth_loss.h:255-355
class ReprojectionLoss {
public:
    ReprojectionLoss(float eps = 1e-6f) : eps_(eps) {}

    torch::Tensor forward(torch::Tensor pred_depth, torch::Tensor gt_depth, torch::Tensor intrinsics,
                          torch::optional<torch::Tensor> valid_mask = torch::nullopt) {
        using namespace cadl_detail;
        auto pred = as_input(pred_depth, "pred_depth", 1);
        auto gt = as_input(gt_depth, "gt_depth", 1);
        TORCH_CHECK(gt.sizes() == pred.sizes(), "cadl: gt_depth and pred_depth differ in shape");
        int batched = 1;
        auto K = as_intrinsics(intrinsics, pred.size(0), pred.device(), batched);
        auto mask = as_mask(valid_mask, pred);
        cadl_params p;
        cadl_default_params(&p);
        p.terms = CADL_TERM_REPROJ; p.w_reproj = 1.0f; p.eps_reproj = eps_; p.k_batched = batched;
        auto loss = apply_fused(pred, gt, torch::Tensor(), K, mask, p, offsetof(cadl_results, loss_reproj));
        return sync_empty_ ? with_reference_empty_rank(loss, offsetof(cadl_results, n_reproj), pred) : loss;
    }

    /// opt in to the reference's zeros(1) when no pixel is valid (depth_loss.h:325-327): costs one host sync per call
    ReprojectionLoss& referenceEmptyRank(bool on) { sync_empty_ = on; return *this; }

    /// The reference's stub (depth_loss.h:343-351): zeros(1).  Kept verbatim in behaviour so callers
    /// see no change; the real warp is the opt-in forwardPhotometricWarp below.
    torch::Tensor forwardPhotometric(torch::Tensor pred_depth, torch::Tensor gt_depth, torch::Tensor intrinsics,
                                     torch::Tensor source_image, torch::Tensor target_image) {
        (void)gt_depth; (void)intrinsics; (void)source_image; (void)target_image;
        return torch::zeros(1, pred_depth.options());
    }

    /// OPT-IN EXTENSION (not in the reference, whose forwardPhotometric is the stub above): the photometric
    /// reprojection documents/algorithms_and_theory.md:18-77 describes -- back-project target pixels with pred depth
    /// and K, move them with T = [R|t] (B,4,4 row-major, target -> source), project with K, sample source_image
    /// bilinearly (grid_sample convention of src/layers/pcl_layer.h:104-108: zeros padding, align_corners = false)
    /// and average the channel-mean L1 residual over the pixels that land inside the source image with Z > eps.
    /// Differentiable w.r.t. pred_depth (explicit backward kernel).  Returns a 0-dim tensor; zeros when no pixel is valid.
    torch::Tensor forwardPhotometricWarp(torch::Tensor pred_depth, torch::Tensor intrinsics, torch::Tensor pose,
                                         torch::Tensor source_image, torch::Tensor target_image) {
        using namespace cadl_detail;
        auto pred = as_input(pred_depth, "pred_depth", 1);
        auto src = as_input(source_image, "source_image", 3);
        auto tgt = as_input(target_image, "target_image", 3);
        TORCH_CHECK(src.sizes() == tgt.sizes() && src.size(0) == pred.size(0) && src.size(2) == pred.size(2) &&
                    src.size(3) == pred.size(3), "cadl: images and depth differ in shape");
        int batched = 1;
        auto K = as_intrinsics(intrinsics, pred.size(0), pred.device(), batched);
        TORCH_CHECK(pose.defined() && pose.scalar_type() == torch::kFloat32 && pose.device() == pred.device() &&
                    pose.dim() == 3 && pose.size(0) == pred.size(0) && pose.size(1) == 4 && pose.size(2) == 4,
                    "cadl: pose must be a float32 (B,4,4) tensor on pred's device");
        const bool want_grad = pred.requires_grad() && torch::GradMode::is_enabled();
        return PhotometricFunction::apply(pred, K, pose.contiguous(), src, tgt, (int64_t)batched, (double)eps_, want_grad);
    }

    float eps() const { return eps_; }

private:
    float eps_;
    bool sync_empty_ = false;
};

/// Weighted sum of the four terms.  reference: depth_loss.h:366-479
class CombinedDepthLoss {
public:
    CombinedDepthLoss(float si_weight = 1.0f, float grad_weight = 0.1f, float smooth_weight = 0.001f,
                      float reproj_weight = 0.01f)
        : si_weight_(si_weight), grad_weight_(grad_weight), smooth_weight_(smooth_weight),
          reproj_weight_(reproj_weight), si_loss_(), grad_loss_(), smooth_loss_(), reproj_loss_() {}

    /// SI + gradient matching + smoothness (depth_loss.h:390-404); shape [1]
    torch::Tensor forward(torch::Tensor pred_depth, torch::Tensor gt_depth, torch::Tensor image,
                          torch::optional<torch::Tensor> valid_mask = torch::nullopt) {
        return run(pred_depth, gt_depth, image, torch::Tensor(), valid_mask, /*with_reproj=*/false);
    }

    /// all four terms (depth_loss.h:416-433); shape [1].  The trainers' per-batch call.
    torch::Tensor forwardWithIntrinsics(torch::Tensor pred_depth, torch::Tensor gt_depth, torch::Tensor image,
                                        torch::Tensor intrinsics,
                                        torch::optional<torch::Tensor> valid_mask = torch::nullopt) {
        return run(pred_depth, gt_depth, image, intrinsics, valid_mask, /*with_reproj=*/true);
    }

    /// depth_loss.h:438-449: {"si_loss","grad_loss","smooth_loss"} -- one fused forward, one host sync
    std::map<std::string, float> getComponents(torch::Tensor pred_depth, torch::Tensor gt_depth, torch::Tensor image,
                                               torch::optional<torch::Tensor> valid_mask = torch::nullopt) {
        auto r = components(pred_depth, gt_depth, image, torch::Tensor(), valid_mask, false);
        return {{"si_loss", r.loss_si}, {"grad_loss", r.loss_grad}, {"smooth_loss", r.loss_smooth}};
    }

    /// depth_loss.h:454-467: adds "reproj_loss"
    std::map<std::string, float> getComponentsWithIntrinsics(torch::Tensor pred_depth, torch::Tensor gt_depth,
                                                             torch::Tensor image, torch::Tensor intrinsics,
                                                             torch::optional<torch::Tensor> valid_mask = torch::nullopt) {
        auto r = components(pred_depth, gt_depth, image, intrinsics, valid_mask, true);
        return {{"si_loss", r.loss_si}, {"grad_loss", r.loss_grad}, {"smooth_loss", r.loss_smooth},
                {"reproj_loss", r.loss_reproj}};
    }

private:
    struct Prepared {
        cadl_detail::StackInputs in;
        cadl_params p;
    };

    Prepared prepare(torch::Tensor pred_depth, torch::Tensor gt_depth, torch::Tensor image, torch::Tensor intrinsics,
                     torch::optional<torch::Tensor> valid_mask, bool with_reproj) {
        using namespace cadl_detail;
        Prepared q;
        q.in.pred = as_input(pred_depth, "pred_depth", 1);
        q.in.gt = as_input(gt_depth, "gt_depth", 1);
        q.in.rgb = as_input(image, "image", 3);
        TORCH_CHECK(q.in.gt.sizes() == q.in.pred.sizes(), "cadl: gt_depth and pred_depth differ in shape");
        TORCH_CHECK(q.in.rgb.size(0) == q.in.pred.size(0) && q.in.rgb.size(2) == q.in.pred.size(2) &&
                        q.in.rgb.size(3) == q.in.pred.size(3), "cadl: image and pred_depth differ in shape");
        TORCH_CHECK(q.in.gt.device() == q.in.pred.device() && q.in.rgb.device() == q.in.pred.device(),
                    "cadl: all tensors must be on one device");   // production_trainer.h:192-194
        cadl_default_params(&q.p);
        q.p.terms = CADL_TERM_SI | CADL_TERM_GRAD | CADL_TERM_SMOOTH;
        q.p.w_si = si_weight_; q.p.w_grad = grad_weight_; q.p.w_smooth = smooth_weight_; q.p.w_reproj = reproj_weight_;
        q.p.si_lambda = si_loss_.lambda(); q.p.eps_si = si_loss_.eps();
        q.p.num_scales = grad_loss_.num_scales(); q.p.eps_grad = grad_loss_.eps();
        q.p.eps_smooth = smooth_loss_.eps(); q.p.eps_reproj = reproj_loss_.eps();
        if (with_reproj) {
            int batched = 1;
            q.in.K = as_intrinsics(intrinsics, q.in.pred.size(0), q.in.pred.device(), batched);
            q.p.k_batched = batched;
            q.p.terms |= CADL_TERM_REPROJ;
        }
        q.in.mask = as_mask(valid_mask, q.in.pred);
        q.in.B = (int)q.in.pred.size(0); q.in.H = (int)q.in.pred.size(2); q.in.W = (int)q.in.pred.size(3);
        return q;
    }

    torch::Tensor run(torch::Tensor pred_depth, torch::Tensor gt_depth, torch::Tensor image, torch::Tensor intrinsics,
                      torch::optional<torch::Tensor> valid_mask, bool with_reproj) {
        auto q = prepare(pred_depth, gt_depth, image, intrinsics, valid_mask, with_reproj);
        return cadl_detail::apply_fused(q.in.pred, q.in.gt, q.in.rgb, q.in.K, q.in.mask, q.p,
                                        offsetof(cadl_results, loss_total)).reshape({1});
    }

    cadl_results components(torch::Tensor pred_depth, torch::Tensor gt_depth, torch::Tensor image,
                            torch::Tensor intrinsics, torch::optional<torch::Tensor> valid_mask, bool with_reproj) {
        auto q = prepare(pred_depth, gt_depth, image, intrinsics, valid_mask, with_reproj);
        torch::NoGradGuard ng;
        auto res = cadl_detail::run_stack(q.in, q.p, torch::Tensor());
        return cadl_detail::results_to_host(res);
    }

    float si_weight_;
    float grad_weight_;
    float smooth_weight_;
    float reproj_weight_;

    ScaleInvariantLoss si_loss_;
    GradientMatchingLoss grad_loss_;
    SmoothnessLoss smooth_loss_;
    ReprojectionLoss reproj_loss_;
};

}  // namespace camera_aware_depth

#endif  // DEPTH_LOSS_H
