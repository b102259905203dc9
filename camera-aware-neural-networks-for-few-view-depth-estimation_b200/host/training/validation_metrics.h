// The trainers' validation metrics as a standalone header.
//
// In the reference this is a private member of two trainer classes
// (src/training/tensorboard_trainer_enhanced.h:65-74,400-439; duplicate at tensorboard_trainer.h:348-387)
// built from 2 masked_select compactions and 7 .item() syncs per validation sample.  Here it is one
// pass of the cadl metrics kernel and one device->host read.  A trainer switches over by replacing
// the body of its computeDepthMetrics with `return camera_aware_depth::computeDepthMetrics(pred, gt);`
// (INTEGRATION.md).
#ifndef CADL_VALIDATION_METRICS_H
#define CADL_VALIDATION_METRICS_H

#include <torch/torch.h>

#include "../cadl_torch.h"

namespace camera_aware_depth {

/// Same fields and defaults as TensorBoardTrainerEnhanced::ValidationMetrics
/// (tensorboard_trainer_enhanced.h:65-74); `loss` is filled by the caller (:363).
struct ValidationMetrics {
    float loss = 0.0f;
    float abs_rel = 0.0f;
    float sq_rel = 0.0f;
    float rmse = 0.0f;
    float rmse_log = 0.0f;
    float a1 = 0.0f;   // delta < 1.25
    float a2 = 0.0f;   // delta < 1.25^2
    float a3 = 0.0f;   // delta < 1.25^3
};

namespace cadl_detail {
inline cadl_results train_metrics_raw(const torch::Tensor& pred_in, const torch::Tensor& gt_in) {
    TORCH_CHECK(pred_in.is_cuda() && gt_in.is_cuda(), "cadl: metrics need CUDA tensors (this build has no CPU path)");
    TORCH_CHECK(pred_in.scalar_type() == torch::kFloat32 && gt_in.scalar_type() == torch::kFloat32,
                "cadl: metrics need float32 tensors");
    TORCH_CHECK(pred_in.numel() == gt_in.numel() && pred_in.numel() > 0, "cadl: pred and gt differ in size");
    auto pred = pred_in.detach().contiguous();
    auto gt = gt_in.detach().contiguous();
    const auto dev = pred.device();
    c10::cuda::CUDAGuard guard(dev);
    const int64_t n = pred.numel();
    size_t ws_bytes = 0;
    auto ws = workspace_for(dev, 1, 1, (int)n, &ws_bytes);
    auto results = new_results(dev);
    int rc = cadl_metrics(pred.data_ptr<float>(), gt.data_ptr<float>(), nullptr, (size_t)n, CADL_METRICS_TRAIN,
                          0.1f, 10.0f, reinterpret_cast<cadl_results*>(results.data_ptr<uint8_t>()),
                          ws.data_ptr<uint8_t>(), ws_bytes, current_stream(dev));
    check_rc(rc, "cadl_metrics");
    return results_to_host(results);
}
}  // namespace cadl_detail

/// abs_rel, sq_rel, rmse, rmse_log, a1..a3 over pixels with gt > 0; no clamping; logs of x + 1e-8.
/// reference: tensorboard_trainer_enhanced.h:400-439
inline ValidationMetrics computeDepthMetrics(const torch::Tensor& pred, const torch::Tensor& gt) {
    cadl_results r = cadl_detail::train_metrics_raw(pred, gt);
    ValidationMetrics m;
    m.abs_rel = r.train[0]; m.sq_rel = r.train[1]; m.rmse = r.train[2]; m.rmse_log = r.train[3];
    m.a1 = r.train[4]; m.a2 = r.train[5]; m.a3 = r.train[6];
    return m;
}

/// Extension: exact integer counts {n_valid, #(ratio<1.25), #(<1.5625), #(<1.953125)}.
inline void computeDepthMetricsCounts(const torch::Tensor& pred, const torch::Tensor& gt, int64_t counts[4]) {
    cadl_results r = cadl_detail::train_metrics_raw(pred, gt);
    for (int i = 0; i < 4; ++i) counts[i] = r.train_counts[i];
}

}  // namespace camera_aware_depth

#endif  // CADL_VALIDATION_METRICS_H
