// LibTorch <-> cadl C-ABI glue shared by the drop-in headers (loss/depth_loss.h,
// evaluation/depth_metrics.h, training/validation_metrics.h).
//
// What lives here: input validation (the reference's tensor conventions), the
// per-(device, stream) workspace cache, and ONE torch::autograd::Function that runs the fused
// forward+backward kernels in forward() and hands the stored gradient back in backward().
// There is no CPU path: a non-CUDA tensor is an error, as is a missing libcadl.so at link time.
#ifndef CADL_TORCH_H
#define CADL_TORCH_H

#include <torch/torch.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <cstddef>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>

#include "cadl.h"

namespace camera_aware_depth {
namespace cadl_detail {

inline void check_rc(int rc, const char* what) {
    // C-ABI status -> c10::Error (a std::exception): lands in train_main.cpp:503-506's catch
    TORCH_CHECK(rc == CADL_OK, what, ": ", cadl_error_string(rc));
}

inline torch::Tensor as_input(const torch::Tensor& t, const char* name, int64_t channels) {
    TORCH_CHECK(t.defined(), "cadl: ", name, " is undefined");
    TORCH_CHECK(t.is_cuda(), "cadl: ", name, " must be a CUDA tensor (this build has no CPU path)");
    TORCH_CHECK(t.scalar_type() == torch::kFloat32, "cadl: ", name, " must be float32");
    TORCH_CHECK(t.dim() == 4 && t.size(1) == channels, "cadl: ", name, " must be (B,", channels, ",H,W)");
    return t.contiguous();
}

inline torch::Tensor as_intrinsics(const torch::Tensor& K, int64_t B, const torch::Device& dev, int& batched) {
    TORCH_CHECK(K.defined() && K.scalar_type() == torch::kFloat32, "cadl: intrinsics must be float32");
    TORCH_CHECK(K.device() == dev, "cadl: intrinsics must be on the same device as pred");
    if (K.dim() == 2) {   // (3,3) broadcast: depth_loss.h:278-280
        TORCH_CHECK(K.size(0) == 3 && K.size(1) == 3, "cadl: intrinsics must be (3,3) or (B,3,3)");
        batched = 0;
    } else {
        TORCH_CHECK(K.dim() == 3 && K.size(0) == B && K.size(1) == 3 && K.size(2) == 3,
                    "cadl: intrinsics must be (3,3) or (B,3,3)");
        batched = 1;
    }
    return K.contiguous();
}

inline torch::Tensor as_mask(const torch::optional<torch::Tensor>& m, const torch::Tensor& like) {
    if (!m.has_value() || !m.value().defined()) return torch::Tensor();
    auto t = m.value();
    TORCH_CHECK(t.device() == like.device(), "cadl: valid_mask must be on the same device as pred");
    if (t.dim() == 3) t = t.unsqueeze(1);
    TORCH_CHECK(t.numel() == like.numel(), "cadl: valid_mask must have as many elements as pred");
    return t.to(torch::kBool).contiguous();
}

inline cadl_stream_t current_stream(const torch::Device& dev) {
    return static_cast<cadl_stream_t>(c10::cuda::getCurrentCUDAStream(dev.index()).stream());
}

// One zero-initialised workspace per (device, stream); grown on demand.  Calls on one stream are
// ordered, so sharing is safe; the kernels leave the workspace clean (include/cadl.h).
inline torch::Tensor workspace_for(const torch::Device& dev, int B, int H, int W, size_t* bytes_out) {
    static std::mutex mu;
    static std::map<std::tuple<int, void*>, torch::Tensor> cache;
    const size_t need = cadl_workspace_bytes(B, H, W);
    TORCH_CHECK(need > 0, "cadl: bad problem size");
    void* st = current_stream(dev);
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_tuple((int)dev.index(), st);
    auto it = cache.find(key);
    if (it == cache.end() || (size_t)it->second.numel() < need) {
        auto t = torch::zeros({(int64_t)need}, torch::TensorOptions().dtype(torch::kUInt8).device(dev));
        cache[key] = t;
        it = cache.find(key);
    }
    *bytes_out = (size_t)it->second.numel();
    return it->second;
}

inline torch::Tensor new_results(const torch::Device& dev) {
    return torch::empty({(int64_t)sizeof(cadl_results)}, torch::TensorOptions().dtype(torch::kUInt8).device(dev));
}

// float32 0-dim view of one member of the on-device cadl_results
inline torch::Tensor result_scalar(const torch::Tensor& results, size_t offset) {
    return results.narrow(0, (int64_t)offset, 4).view(torch::kFloat32).reshape({});
}

inline cadl_results results_to_host(const torch::Tensor& results) {
    auto h = results.to(torch::kCPU);   // the one device->host sync of the getComponents*/metrics calls
    cadl_results r;
    std::memcpy(&r, h.data_ptr<uint8_t>(), sizeof(r));
    return r;
}

// the on-device cadl_results of this thread's most recent fused call (what the opt-in host reads look at)
inline torch::Tensor& last_results() {
    thread_local torch::Tensor t;
    return t;
}

struct StackInputs {
    torch::Tensor pred, gt, rgb, K, mask;
    int B = 0, H = 0, W = 0;
};

// Runs the fused kernels.  grad may be undefined (forward only).  Returns the on-device results.
inline torch::Tensor run_stack(const StackInputs& in, cadl_params p, torch::Tensor grad) {
    const auto dev = in.pred.device();
    c10::cuda::CUDAGuard guard(dev);
    size_t ws_bytes = 0;
    auto ws = workspace_for(dev, in.B, in.H, in.W, &ws_bytes);
    auto results = new_results(dev);
    int rc = cadl_stack_fwd_bwd(
        in.pred.data_ptr<float>(), in.gt.defined() ? in.gt.data_ptr<float>() : nullptr,
        in.rgb.defined() ? in.rgb.data_ptr<float>() : nullptr, in.K.defined() ? in.K.data_ptr<float>() : nullptr,
        in.mask.defined() ? reinterpret_cast<const uint8_t*>(in.mask.data_ptr<bool>()) : nullptr, in.B, in.H, in.W,
        &p, grad.defined() ? grad.data_ptr<float>() : nullptr,
        reinterpret_cast<cadl_results*>(results.data_ptr<uint8_t>()), ws.data_ptr<uint8_t>(), ws_bytes,
        current_stream(dev));
    check_rc(rc, "cadl_stack_fwd_bwd");
    last_results() = results;
    return results;
}

// The autograd node of every loss class.  forward() launches forward AND backward kernels (one fused
// pass; the un-scaled dL/dpred is kept), backward() multiplies by the incoming gradient on the
// device -- a no-op when it is exactly 1 (loss.backward()).
struct FusedLossFunction : public torch::autograd::Function<FusedLossFunction> {
    static torch::Tensor forward(torch::autograd::AutogradContext* ctx, torch::Tensor pred, torch::Tensor gt,
                                 torch::Tensor rgb, torch::Tensor K, torch::Tensor mask, int64_t B, int64_t H,
                                 int64_t W, std::string params_blob, int64_t result_offset, bool want_grad) {
        cadl_params p;
        std::memcpy(&p, params_blob.data(), sizeof(p));
        // absent inputs arrive as 0-element placeholders (Function::apply rejects undefined tensors)
        auto opt = [](const torch::Tensor& t) { return t.numel() == 0 ? torch::Tensor() : t; };
        StackInputs in;
        in.pred = pred; in.gt = opt(gt); in.rgb = opt(rgb); in.K = opt(K); in.mask = opt(mask);
        in.B = (int)B; in.H = (int)H; in.W = (int)W;
        torch::Tensor grad;
        if (want_grad) grad = torch::empty_like(pred);
        auto results = run_stack(in, p, grad);
        if (want_grad) {
            ctx->save_for_backward({grad});
            ctx->saved_data["used"] = false;
        }
        return result_scalar(results, (size_t)result_offset);
    }

    static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx,
                                                   torch::autograd::variable_list grad_outputs) {
        auto saved = ctx->get_saved_variables();
        TORCH_CHECK(saved.size() == 1, "cadl: backward without a stored gradient");
        auto grad = saved[0];
        auto go = grad_outputs[0];
        TORCH_CHECK(go.defined() && go.numel() == 1, "cadl: the loss is a scalar");
        const auto dev = grad.device();
        c10::cuda::CUDAGuard guard(dev);
        go = go.to(dev, torch::kFloat32).contiguous();
        torch::Tensor out = grad;
        if (ctx->saved_data["used"].toBool()) {
            // second backward through the same node (retain_graph): the stored gradient was already
            // scaled in place, so scale a copy of it by go / previous instead.
            auto prev = ctx->saved_data["prev"].toTensor();
            go = (go.reshape({1}) / prev.reshape({1})).contiguous();
            out = grad.clone();
        }
        int rc = cadl_scale_grad(out.data_ptr<float>(), go.data_ptr<float>(), out.data_ptr<float>(),
                                 (size_t)out.numel(), current_stream(dev));
        check_rc(rc, "cadl_scale_grad");
        if (!ctx->saved_data["used"].toBool()) {
            ctx->saved_data["used"] = true;
            ctx->saved_data["prev"] = grad_outputs[0].detach().to(dev, torch::kFloat32).clone();
        }
        torch::Tensor none;
        return {out, none, none, none, none, none, none, none, none, none, none};
    }
};

// Autograd node of the opt-in photometric warp (ReprojectionLoss::forwardPhotometricWarp): same shape as
// FusedLossFunction -- forward() runs forward + backward kernels, backward() scales the stored gradient.
struct PhotometricFunction : public torch::autograd::Function<PhotometricFunction> {
    static torch::Tensor forward(torch::autograd::AutogradContext* ctx, torch::Tensor pred, torch::Tensor K, torch::Tensor T,
                                 torch::Tensor src, torch::Tensor tgt, int64_t batched, double eps, bool want_grad) {
        const auto dev = pred.device();
        c10::cuda::CUDAGuard guard(dev);
        const int B = (int)pred.size(0), H = (int)pred.size(2), W = (int)pred.size(3);
        size_t ws_bytes = 0;
        auto ws = workspace_for(dev, B, H, W, &ws_bytes);
        auto results = new_results(dev);
        torch::Tensor grad;
        if (want_grad) grad = torch::empty_like(pred);
        int rc = cadl_photometric_fwd_bwd(pred.data_ptr<float>(), K.data_ptr<float>(), (int)batched, T.data_ptr<float>(),
                                          src.data_ptr<float>(), tgt.data_ptr<float>(), B, H, W, (float)eps, 1.0f,
                                          grad.defined() ? grad.data_ptr<float>() : nullptr,
                                          reinterpret_cast<cadl_results*>(results.data_ptr<uint8_t>()), ws.data_ptr<uint8_t>(),
                                          ws_bytes, current_stream(dev));
        check_rc(rc, "cadl_photometric_fwd_bwd");
        if (want_grad) ctx->save_for_backward({grad});
        return result_scalar(results, offsetof(cadl_results, loss_reproj));
    }
    static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx,
                                                   torch::autograd::variable_list grad_outputs) {
        auto saved = ctx->get_saved_variables();
        TORCH_CHECK(saved.size() == 1, "cadl: backward without a stored gradient");
        auto go = grad_outputs[0];
        TORCH_CHECK(go.defined() && go.numel() == 1, "cadl: the loss is a scalar");
        const auto dev = saved[0].device();
        c10::cuda::CUDAGuard guard(dev);
        go = go.to(dev, torch::kFloat32).contiguous();
        auto out = saved[0].clone();              // (a copy: the node may be run again with retain_graph)
        int rc = cadl_scale_grad(out.data_ptr<float>(), go.data_ptr<float>(), out.data_ptr<float>(), (size_t)out.numel(),
                                 current_stream(dev));
        check_rc(rc, "cadl_scale_grad");
        torch::Tensor none;
        return {out, none, none, none, none, none, none, none};
    }
};

// The reference returns zeros(1) -- rank 1, no graph -- when no pixel is valid (depth_loss.h:53-55, 325-327), which it
// learns from masked_select().numel(): a host sync.  The drop-in's default keeps the step free of syncs (0-dim result,
// value 0, zero gradient); a caller that wants the reference's rank too opts into that one 8-byte read.
inline torch::Tensor with_reference_empty_rank(const torch::Tensor& loss, size_t count_offset, const torch::Tensor& like) {
    torch::Tensor results = last_results();          // the block the call that produced `loss` just wrote (this thread)
    TORCH_CHECK(results.defined(), "cadl: loss tensor without its result block");
    const int64_t n = results.narrow(0, (int64_t)count_offset, 8).view(torch::kInt64).item<int64_t>();
    return n == 0 ? torch::zeros(1, like.options().requires_grad(false)) : loss;
}

struct TermCall {
    cadl_params p;
    size_t result_offset;
};

inline torch::Tensor apply_fused(torch::Tensor pred, torch::Tensor gt, torch::Tensor rgb, torch::Tensor K,
                                 torch::Tensor mask, const cadl_params& p, size_t result_offset) {
    const int64_t B = pred.size(0), H = pred.size(2), W = pred.size(3);
    std::string blob(reinterpret_cast<const char*>(&p), sizeof(p));
    const bool want_grad = pred.requires_grad() && torch::GradMode::is_enabled();
    auto none_f = torch::empty({0}, pred.options());
    auto ph = [&](const torch::Tensor& t) { return t.defined() ? t : none_f; };
    return FusedLossFunction::apply(pred, ph(gt), ph(rgb), ph(K), ph(mask), B, H, W, blob, (int64_t)result_offset,
                                    want_grad);
}

}  // namespace cadl_detail
}  // namespace camera_aware_depth

#endif  // CADL_TORCH_H
