// Fused gradient-norm + clipping for the trainers' step.
//
// Replaces, in src/training/tensorboard_trainer_enhanced.h:
//   :300-302  torch::nn::utils::clip_grad_norm_(model_->parameters(), config_.grad_clip_value);
//   :560-571  computeGradientNorm(): one norm().item<float>() host sync PER PARAMETER TENSOR
// by two multi-tensor launches over all gradients with the total norm left on the device.
#ifndef CADL_GRAD_CLIP_H
#define CADL_GRAD_CLIP_H

#include <torch/torch.h>

#include <vector>

#include "../cadl_torch.h"

namespace camera_aware_depth {

class FusedGradClipper {
public:
    /// clip_grad_norm_(params, max_norm): returns a device tensor {total_norm, clip_coef}; no host sync.
    /// With clip = false only the norm is computed (computeGradientNorm).
    torch::Tensor run(const std::vector<torch::Tensor>& params, float max_norm, bool clip = true) {
        using namespace cadl_detail;
        std::vector<int64_t> meta;   // ptrs | sizes | chunk prefix, one H2D copy
        std::vector<torch::Tensor> grads;
        for (const auto& p : params) {
            if (!p.grad().defined()) continue;
            auto g = p.grad();
            TORCH_CHECK(g.is_cuda() && g.scalar_type() == torch::kFloat32 && g.is_contiguous(),
                        "cadl: gradients must be contiguous float32 CUDA tensors");
            grads.push_back(g);
        }
        TORCH_CHECK(!grads.empty(), "cadl: no gradients to clip");
        const auto dev = grads[0].device();
        c10::cuda::CUDAGuard guard(dev);
        const int64_t n = (int64_t)grads.size();
        meta.resize(3 * n + 1);
        int64_t chunks = 0;
        for (int64_t i = 0; i < n; ++i) {
            meta[i] = reinterpret_cast<int64_t>(grads[i].data_ptr<float>());
            meta[n + i] = grads[i].numel();
            meta[2 * n + i] = chunks;
            chunks += (grads[i].numel() + 4095) / 4096;
        }
        meta[3 * n] = chunks;
        auto host = torch::from_blob(meta.data(), {(int64_t)meta.size()}, torch::kInt64).pin_memory();
        auto devmeta = host.to(dev, /*non_blocking=*/true);
        if (!ws_.defined() || ws_.device() != dev) {
            ws_ = torch::zeros({(int64_t)cadl_clip_workspace_bytes()}, torch::TensorOptions().dtype(torch::kUInt8).device(dev));
        }
        auto out = torch::empty({2}, torch::TensorOptions().dtype(torch::kFloat32).device(dev));
        const int64_t* m = devmeta.data_ptr<int64_t>();
        int rc = cadl_clip_grad_norm(reinterpret_cast<float* const*>(m), reinterpret_cast<const long long*>(m + n),
                                     reinterpret_cast<const long long*>(m + 2 * n), (int)n, (long long)chunks, max_norm,
                                     out.data_ptr<float>(), ws_.data_ptr<uint8_t>(), (size_t)ws_.numel(), clip ? 1 : 0,
                                     current_stream(dev));
        check_rc(rc, "cadl_clip_grad_norm");
        keep_ = devmeta;   // keep the metadata alive until the stream has consumed it
        return out;
    }

private:
    torch::Tensor ws_, keep_;
};

}  // namespace camera_aware_depth

#endif  // CADL_GRAD_CLIP_H
