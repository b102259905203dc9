// TEST INFRASTRUCTURE (oracle/): compiled only into oracle/_ref/libcadl_refharness.so, the build of the step harness
// against the UNMODIFIED reference headers.  Two things the reference's public headers do not give the harness:
//   * integer delta counts with the reference's own ops -- the float mean DepthMetrics::compute reports cannot hold
//     counts above 2^24 exactly, so count parity is defined against (ratio < thr).sum() of the same tensors;
//   * the trainers' private computeDepthMetrics (src/training/tensorboard_trainer_enhanced.h:400-439; duplicate at
//     tensorboard_trainer.h:348-387), restated op for op: the trainer headers cannot be included (they pull in OpenCV).
// Nothing on the product path includes this file.
#pragma once
#include <torch/torch.h>

namespace {
// Integer delta counts with the reference's own ops (depth_metrics.h:154-161,57-66,219-229):
// the float mean the reference reports cannot hold counts above 2^24 exactly, so parity on
// counts is defined against (ratio < thr).sum() of the same tensors.
void ref_ops_eval_counts(torch::Tensor pred, torch::Tensor gt, torch::optional<torch::Tensor> user,
                         float min_d, float max_d, int64_t counts[4]) {
    if (pred.dim() == 3) pred = pred.unsqueeze(1);
    if (gt.dim() == 3) gt = gt.unsqueeze(1);
    auto mask = (gt > min_d) & (gt < max_d);
    if (user.has_value()) {
        auto um = user.value();
        if (um.dim() == 3) um = um.unsqueeze(1);
        mask = mask & um.to(torch::kBool);
    }
    auto p = pred.masked_select(mask);
    auto g = gt.masked_select(mask);
    counts[0] = p.numel();
    counts[1] = counts[2] = counts[3] = 0;
    if (counts[0] == 0) return;
    p = torch::clamp(p, min_d, max_d);
    auto ratio = torch::max(p / g, g / p);
    float thr[3] = {1.25f, 1.25f * 1.25f, 1.25f * 1.25f * 1.25f};
    for (int i = 0; i < 3; ++i) counts[1 + i] = (ratio < thr[i]).sum().item<int64_t>();
}

// Restatement, op for op, of the trainers' private computeDepthMetrics
// (src/training/tensorboard_trainer_enhanced.h:400-439; duplicate at tensorboard_trainer.h:348-387).
// The trainer headers cannot be included here: they pull in OpenCV.
struct ValidationMetrics {
    float loss = 0.0f, abs_rel = 0.0f, sq_rel = 0.0f, rmse = 0.0f, rmse_log = 0.0f;
    float a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
};
ValidationMetrics computeDepthMetrics(const torch::Tensor& pred, const torch::Tensor& gt) {
    ValidationMetrics metrics;
    auto pred_flat = pred.view({-1});
    auto gt_flat = gt.view({-1});
    auto valid_mask = gt_flat > 0.0f;
    auto pred_valid = pred_flat.masked_select(valid_mask);
    auto gt_valid = gt_flat.masked_select(valid_mask);
    if (pred_valid.numel() == 0) return metrics;
    auto abs_diff = torch::abs(pred_valid - gt_valid);
    metrics.abs_rel = (abs_diff / gt_valid).mean().item<float>();
    metrics.sq_rel = ((abs_diff * abs_diff) / gt_valid).mean().item<float>();
    metrics.rmse = torch::sqrt((abs_diff * abs_diff).mean()).item<float>();
    auto log_diff = torch::abs(torch::log(pred_valid + 1e-8) - torch::log(gt_valid + 1e-8));
    metrics.rmse_log = torch::sqrt((log_diff * log_diff).mean()).item<float>();
    auto ratio = torch::max(pred_valid / gt_valid, gt_valid / pred_valid);
    metrics.a1 = (ratio < 1.25f).to(torch::kFloat32).mean().item<float>();
    metrics.a2 = (ratio < 1.5625f).to(torch::kFloat32).mean().item<float>();
    metrics.a3 = (ratio < 1.953125f).to(torch::kFloat32).mean().item<float>();
    return metrics;
}

void ref_ops_train_counts(const torch::Tensor& pred, const torch::Tensor& gt, int64_t counts[4]) {
    auto pf = pred.reshape({-1});
    auto gf = gt.reshape({-1});
    auto m = gf > 0.0f;
    auto p = pf.masked_select(m);
    auto g = gf.masked_select(m);
    counts[0] = p.numel();
    counts[1] = counts[2] = counts[3] = 0;
    if (counts[0] == 0) return;
    auto ratio = torch::max(p / g, g / p);
    counts[1] = (ratio < 1.25f).sum().item<int64_t>();
    counts[2] = (ratio < 1.5625f).sum().item<int64_t>();
    counts[3] = (ratio < 1.953125f).sum().item<int64_t>();
}

}  // namespace
