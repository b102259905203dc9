// Device-resident running sums of per-batch scalars: replaces
//   float loss_val = loss.item<float>(); metrics.loss += loss_val * actual_batch_size;      (production_trainer.h:213-216)
//   metrics.loss += loss.item<float>();                                                     (tensorboard_trainer_enhanced.h:363)
// and the host sync they force on every batch; the sums are read once, when the epoch's averages are printed.
#ifndef CADL_LOSS_ACCUMULATOR_H
#define CADL_LOSS_ACCUMULATOR_H

#include <torch/torch.h>

#include <vector>

#include "../cadl_torch.h"

namespace camera_aware_depth {

class DeviceAccumulator {
public:
    /// n scalars per batch (e.g. 1 for the loss; 5 for total + the four terms)
    explicit DeviceAccumulator(int64_t n) : n_(n) {}

    /// sums[i] += weight * values[i]; weight_sum += weight.  values: float32 CUDA tensor with n elements (any shape).
    void add(const torch::Tensor& values, double weight = 1.0) {
        using namespace cadl_detail;
        TORCH_CHECK(values.is_cuda() && values.scalar_type() == torch::kFloat32 && values.numel() == n_,
                    "cadl: DeviceAccumulator::add expects ", n_, " float32 values on a CUDA device");
        auto v = values.detach().contiguous();
        const auto dev = v.device();
        c10::cuda::CUDAGuard guard(dev);
        if (!acc_.defined()) acc_ = torch::zeros({n_ + 1}, torch::TensorOptions().dtype(torch::kFloat64).device(dev));
        int rc = cadl_accumulate(v.data_ptr<float>(), (int)n_, weight, acc_.data_ptr<double>(), current_stream(dev));
        check_rc(rc, "cadl_accumulate");
    }

    /// weighted means, one device->host copy (the only sync); zeros when nothing was added
    std::vector<double> mean() const {
        std::vector<double> out((size_t)n_, 0.0);
        if (!acc_.defined()) return out;
        auto h = acc_.to(torch::kCPU);
        const double* p = h.data_ptr<double>();
        if (p[n_] > 0.0)
            for (int64_t i = 0; i < n_; ++i) out[(size_t)i] = p[i] / p[n_];
        return out;
    }

    void reset() {
        if (acc_.defined()) acc_.zero_();
    }

private:
    int64_t n_;
    torch::Tensor acc_;
};

}  // namespace camera_aware_depth

#endif  // CADL_LOSS_ACCUMULATOR_H
