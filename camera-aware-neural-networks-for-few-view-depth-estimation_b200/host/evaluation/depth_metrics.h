// Drop-in replacement for the reference's src/evaluation/depth_metrics.h
// (DepthMetrics: depth_metrics.h:28-254, MetricsAccumulator: :259-304, formatMetrics: :309-333).
//
// DepthMetrics::compute is one pass of the cadl metrics kernel over pred/gt (8 B/px) and one
// device->host read of the result block, instead of two masked_select compactions, ~25 elementwise
// ops and 11 .item() syncs.  delta thresholds are counted in integers on the device.
#ifndef DEPTH_METRICS_H
#define DEPTH_METRICS_H

#include <torch/torch.h>

#include <array>
#include <cmath>
#include <iomanip>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../cadl_torch.h"

namespace camera_aware_depth {

class DepthMetrics {
public:
    /// 12 metrics over valid pixels (min_depth < gt < max_depth [& valid_mask]); pred is clamped to
    /// [min_depth, max_depth] after masking.  reference: depth_metrics.h:40-88
    static std::map<std::string, float> compute(torch::Tensor pred_depth, torch::Tensor gt_depth,
                                                torch::optional<torch::Tensor> valid_mask = torch::nullopt,
                                                float min_depth = 0.1f, float max_depth = 10.0f) {
        cadl_results r = evaluate(pred_depth, gt_depth, valid_mask, min_depth, max_depth);
        std::map<std::string, float> m;
        static const char* keys[12] = {"abs_rel", "sq_rel", "rmse", "rmse_log", "mae", "log10",
                                       "delta_1.25", "delta_1.25^2", "delta_1.25^3",
                                       "num_valid_pixels", "mean_pred_depth", "mean_gt_depth"};
        for (int i = 0; i < 12; ++i) m[keys[i]] = r.eval[i];
        return m;
    }

    /// Extension: the exact integer counts behind the delta fractions
    /// {n_valid, #(ratio<1.25), #(ratio<1.25^2), #(ratio<1.25^3)}; the float means above cannot
    /// represent them beyond 2^24 pixels.
    static void computeCounts(torch::Tensor pred_depth, torch::Tensor gt_depth,
                              torch::optional<torch::Tensor> valid_mask, float min_depth, float max_depth,
                              int64_t counts[4]) {
        cadl_results r = evaluate(pred_depth, gt_depth, valid_mask, min_depth, max_depth);
        for (int i = 0; i < 4; ++i) counts[i] = r.eval_counts[i];
    }

    /// one map per batch element.  reference: depth_metrics.h:93-117
    static std::vector<std::map<std::string, float>> computePerSample(
        torch::Tensor pred_depth, torch::Tensor gt_depth, torch::optional<torch::Tensor> valid_mask = torch::nullopt,
        float min_depth = 0.1f, float max_depth = 10.0f) {
        const int64_t n = pred_depth.size(0);
        std::vector<std::map<std::string, float>> out;
        out.reserve((size_t)n);
        for (int64_t i = 0; i < n; ++i) {
            torch::optional<torch::Tensor> mi = torch::nullopt;
            if (valid_mask.has_value()) mi = valid_mask.value()[i].unsqueeze(0);
            out.push_back(compute(pred_depth[i].unsqueeze(0), gt_depth[i].unsqueeze(0), mi, min_depth, max_depth));
        }
        return out;
    }

    /// key-wise mean of a list of metric maps (unweighted).  reference: depth_metrics.h:122-141
    static std::map<std::string, float> average(const std::vector<std::map<std::string, float>>& metrics_list) {
        if (metrics_list.empty()) return zeroMetrics();
        std::map<std::string, float> avg;
        for (const auto& kv : metrics_list.front()) {
            float total = 0.0f;
            for (const auto& m : metrics_list) total += m.at(kv.first);
            avg[kv.first] = total / metrics_list.size();
        }
        return avg;
    }

private:
    static std::map<std::string, float> zeroMetrics() {   // depth_metrics.h:238-253
        std::map<std::string, float> z;
        for (const char* k : {"abs_rel", "sq_rel", "rmse", "rmse_log", "mae", "log10", "delta_1.25", "delta_1.25^2",
                              "delta_1.25^3", "num_valid_pixels", "mean_pred_depth", "mean_gt_depth"})
            z[k] = 0.0f;
        return z;
    }

    static cadl_results evaluate(torch::Tensor pred, torch::Tensor gt, torch::optional<torch::Tensor> valid_mask,
                                 float min_depth, float max_depth) {
        using namespace cadl_detail;
        if (pred.dim() == 3) pred = pred.unsqueeze(1);   // depth_metrics.h:50-51
        if (gt.dim() == 3) gt = gt.unsqueeze(1);
        TORCH_CHECK(pred.is_cuda() && gt.is_cuda(), "cadl: metrics need CUDA tensors (this build has no CPU path)");
        TORCH_CHECK(pred.scalar_type() == torch::kFloat32 && gt.scalar_type() == torch::kFloat32,
                    "cadl: metrics need float32 tensors");
        TORCH_CHECK(pred.numel() == gt.numel() && pred.numel() > 0, "cadl: pred and gt differ in size");
        pred = pred.detach().contiguous();
        gt = gt.detach().contiguous();
        auto mask = as_mask(valid_mask, pred);
        const auto dev = pred.device();
        c10::cuda::CUDAGuard guard(dev);
        const int64_t n = pred.numel();
        size_t ws_bytes = 0;
        auto ws = workspace_for(dev, 1, 1, (int)n, &ws_bytes);
        auto results = new_results(dev);
        int rc = cadl_metrics(pred.data_ptr<float>(), gt.data_ptr<float>(),
                              mask.defined() ? reinterpret_cast<const uint8_t*>(mask.data_ptr<bool>()) : nullptr,
                              (size_t)n, CADL_METRICS_EVAL, min_depth, max_depth,
                              reinterpret_cast<cadl_results*>(results.data_ptr<uint8_t>()), ws.data_ptr<uint8_t>(),
                              ws_bytes, current_stream(dev));
        check_rc(rc, "cadl_metrics");
        return results_to_host(results);
    }
};

/// Running mean of metric maps over batches (mean of per-batch means, not pixel-weighted).
/// reference: depth_metrics.h:259-304
class MetricsAccumulator {
public:
    MetricsAccumulator() : count_(0) {}

    void update(const std::map<std::string, float>& metrics) {
        for (const auto& kv : metrics) sums_[kv.first] += kv.second;
        ++count_;
    }

    std::map<std::string, float> average() const {
        std::map<std::string, float> avg;
        if (count_ == 0) return avg;
        for (const auto& kv : sums_) avg[kv.first] = kv.second / count_;
        return avg;
    }

    void reset() {
        sums_.clear();
        count_ = 0;
    }

    int64_t count() const { return count_; }

private:
    std::map<std::string, float> sums_;
    int64_t count_;
};

/// Human-readable block; same fields, order and 4-digit precision as depth_metrics.h:309-333.
inline std::string formatMetrics(const std::map<std::string, float>& metrics) {
    std::ostringstream os;
    os << std::fixed << std::setprecision(4);
    os << "Error Metrics:\n";
    os << "  AbsRel:  " << metrics.at("abs_rel") << "\n";
    os << "  RMSE:    " << metrics.at("rmse") << "\n";
    os << "  RMSElog: " << metrics.at("rmse_log") << "\n";
    os << "  MAE:     " << metrics.at("mae") << "\n";
    os << "\nAccuracy Metrics (%):\n";
    os << "  \xCE\xB4 < 1.25:    " << (metrics.at("delta_1.25") * 100.0f) << "%\n";
    os << "  \xCE\xB4 < 1.25\xC2\xB2:   " << (metrics.at("delta_1.25^2") * 100.0f) << "%\n";
    os << "  \xCE\xB4 < 1.25\xC2\xB3:   " << (metrics.at("delta_1.25^3") * 100.0f) << "%\n";
    os << "\nStatistics:\n";
    os << "  Valid pixels: " << static_cast<int>(metrics.at("num_valid_pixels")) << "\n";
    os << "  Mean pred:    " << metrics.at("mean_pred_depth") << "m\n";
    os << "  Mean GT:      " << metrics.at("mean_gt_depth") << "m\n";
    return os.str();
}

}  // namespace camera_aware_depth

#endif  // DEPTH_METRICS_H
