// Trainer-shaped step harness for the depth-loss hot path.
//
// This translation unit is written ONLY against the public class interface of
//   loss/depth_loss.h          (reference: src/loss/depth_loss.h:20-479)
//   evaluation/depth_metrics.h (reference: src/evaluation/depth_metrics.h:28-333)
// and is compiled twice from the same source:
//   * with  -I <pkg>/host  -DCADL_DROPIN   -> libcadl_host.so  (the B200 drop-in, product path)
//   * with  -I /root/reference/src         -> oracle/_ref/libcadl_refharness.so (the unmodified
//                                             reference on LibTorch: parity oracle + CPU baseline)
// That one source builds against both header sets is the source-compatibility check of the
// drop-in boundary.
//
// The step it performs is the one the reference trainers perform per batch
// (src/training/production_trainer.h:192-215): stack -> .to(device) -> forwardWithIntrinsics ->
// backward -> loss.item<float>().  The model is replaced by a leaf tensor `pred` with
// requires_grad, so pred.grad() is exactly dLoss/dpred.
//
// Everything crossing this boundary is plain C (host pointers + sizes); no torch types.

#include <sstream>   // the reference's depth_metrics.h uses these without including them
#include <iomanip>   // (src/evaluation/depth_metrics.h:309-312,219)
#include <array>
#include <torch/torch.h>
#include <chrono>
#include <cstring>
#include <cstdint>
#include <string>

#include "loss/depth_loss.h"
#include "evaluation/depth_metrics.h"
#ifndef CADL_DROPIN
#include "ref_restatements.h"      // oracle/: what the reference keeps private or reports only as a float mean
#else
#include "training/validation_metrics.h"
#include "training/grad_clip.h"
#include "training/loss_accumulator.h"
#include "data/batch_prep.h"
#include "preprocessing/ray_directions.h"
#endif

using namespace camera_aware_depth;

extern "C" {

typedef struct {
    int B, H, W;
    int k_batched;   // 1: K is (B,3,3); 0: K is (3,3) broadcast (depth_loss.h:278-280)
    int device;      // -1: CPU, >=0: CUDA ordinal
    float w_si, w_grad, w_smooth, w_reproj;
    int term;        // 0 forwardWithIntrinsics, 1 SI, 2 grad-matching, 3 smoothness, 4 reprojection,
                     // 5 CombinedDepthLoss::forward (no intrinsics)
    float upstream;  // (loss * upstream).backward(); 1.0f reproduces the trainers
} cadh_step_cfg;

}  // extern "C"

namespace {

torch::Device pick_device(int device) {
    return device < 0 ? torch::Device(torch::kCPU) : torch::Device(torch::kCUDA, device);
}

void sync_device(const torch::Device& dev) {
    if (dev.is_cuda()) torch::cuda::synchronize(dev.index());
}

struct Batch {
    torch::Tensor pred, gt, rgb, K;
    torch::optional<torch::Tensor> mask;
};

torch::Tensor host_view(const float* p, at::IntArrayRef shape) {
    return torch::from_blob(const_cast<float*>(p), shape, torch::kFloat32);
}

// production_trainer.h:192-194 : host batch -> device
Batch to_device(const cadh_step_cfg& c, const float* pred, const float* gt, const float* rgb,
                const float* K, const uint8_t* mask, const torch::Device& dev) {
    Batch b;
    b.pred = host_view(pred, {c.B, 1, c.H, c.W}).to(dev).clone();
    b.gt = host_view(gt, {c.B, 1, c.H, c.W}).to(dev).clone();
    if (rgb) b.rgb = host_view(rgb, {c.B, 3, c.H, c.W}).to(dev).clone();
    if (K) {
        b.K = (c.k_batched ? host_view(K, {c.B, 3, 3}) : host_view(K, {3, 3})).to(dev).clone();
    }
    if (mask) {
        b.mask = torch::from_blob(const_cast<uint8_t*>(mask), {c.B, 1, c.H, c.W}, torch::kBool)
                     .to(dev).clone();
    }
    return b;
}

torch::Tensor run_term(const cadh_step_cfg& c, CombinedDepthLoss& combined, Batch& b) {
    switch (c.term) {
        case 0: return combined.forwardWithIntrinsics(b.pred, b.gt, b.rgb, b.K, b.mask);
        case 1: { ScaleInvariantLoss l; return l.forward(b.pred, b.gt, b.mask); }
        case 2: { GradientMatchingLoss l; return l.forward(b.pred, b.gt, b.mask); }
        case 3: { SmoothnessLoss l; return l.forward(b.pred, b.rgb); }
        case 4: { ReprojectionLoss l; return l.forward(b.pred, b.gt, b.K, b.mask); }
        case 5: return combined.forward(b.pred, b.gt, b.rgb, b.mask);
        default: TORCH_CHECK(false, "cadh: unknown term ", c.term);
    }
}

int fail(char* err, int errlen, const std::exception& e) {
    if (err && errlen > 0) {
        std::strncpy(err, e.what(), errlen - 1);
        err[errlen - 1] = 0;
    }
    return 1;
}


const char* kEvalKeys[12] = {"abs_rel", "sq_rel", "rmse", "rmse_log", "mae", "log10",
                             "delta_1.25", "delta_1.25^2", "delta_1.25^3",
                             "num_valid_pixels", "mean_pred_depth", "mean_gt_depth"};

}  // namespace

extern "C" {

const char* cadh_build_info() {
#ifdef CADL_DROPIN
    return "cadl drop-in headers (sm_100a CUDA kernels behind the reference class API)";
#else
    return "unmodified reference headers on LibTorch";
#endif
}

int cadh_is_dropin() {
#ifdef CADL_DROPIN
    return 1;
#else
    return 0;
#endif
}

int cadh_num_threads() { return torch::get_num_threads(); }
void cadh_set_num_threads(int n) { torch::set_num_threads(n); }
int cadh_cuda_available() { return torch::cuda::is_available() ? 1 : 0; }

// One training-shaped step: H2D, forward, backward, D2H of the loss (and of pred.grad if asked).
// out_loss_meta[0] = loss.dim(), [1] = loss.numel()  (ranks are part of the contract: SURVEY 8b).
int cadh_loss_step(const cadh_step_cfg* cfg, const float* pred, const float* gt, const float* rgb,
                   const float* K, const uint8_t* mask, float* out_loss, int64_t* out_loss_meta,
                   float* out_grad, char* err, int errlen) {
    try {
        const cadh_step_cfg& c = *cfg;
        auto dev = pick_device(c.device);
        Batch b = to_device(c, pred, gt, rgb, K, mask, dev);
        b.pred.set_requires_grad(true);
        CombinedDepthLoss combined(c.w_si, c.w_grad, c.w_smooth, c.w_reproj);
        auto loss = run_term(c, combined, b);
        if (out_loss_meta) {
            out_loss_meta[0] = loss.dim();
            out_loss_meta[1] = loss.numel();
        }
        if (out_grad) {
            if (loss.requires_grad()) {
                auto scaled = (c.upstream == 1.0f) ? loss : loss * c.upstream;
                scaled.sum().backward();
            }
            auto g = b.pred.grad();
            if (!g.defined()) g = torch::zeros_like(b.pred);
            auto gh = g.to(torch::kCPU).contiguous();
            std::memcpy(out_grad, gh.data_ptr<float>(), sizeof(float) * gh.numel());
        }
        *out_loss = loss.sum().item<float>();
        return 0;
    } catch (const std::exception& e) {
        return fail(err, errlen, e);
    }
}

// CombinedDepthLoss::getComponentsWithIntrinsics (depth_loss.h:454-467) -> si, grad, smooth, reproj
int cadh_components(const cadh_step_cfg* cfg, const float* pred, const float* gt, const float* rgb,
                    const float* K, const uint8_t* mask, float out4[4], char* err, int errlen) {
    try {
        const cadh_step_cfg& c = *cfg;
        auto dev = pick_device(c.device);
        Batch b = to_device(c, pred, gt, rgb, K, mask, dev);
        torch::NoGradGuard ng;
        CombinedDepthLoss combined(c.w_si, c.w_grad, c.w_smooth, c.w_reproj);
        auto m = combined.getComponentsWithIntrinsics(b.pred, b.gt, b.rgb, b.K, b.mask);
        out4[0] = m.at("si_loss");
        out4[1] = m.at("grad_loss");
        out4[2] = m.at("smooth_loss");
        out4[3] = m.at("reproj_loss");
        return 0;
    } catch (const std::exception& e) {
        return fail(err, errlen, e);
    }
}

// DepthMetrics::compute (depth_metrics.h:40-88): 12 floats in kEvalKeys order + integer counts
// {n_valid, n(delta<1.25), n(<1.25^2), n(<1.25^3)}.
int cadh_metrics_eval(int B, int H, int W, int device, const float* pred, const float* gt,
                      const uint8_t* mask, float min_d, float max_d, float out12[12],
                      int64_t counts[4], char* err, int errlen) {
    try {
        cadh_step_cfg c{};
        c.B = B; c.H = H; c.W = W; c.device = device;
        auto dev = pick_device(device);
        Batch b = to_device(c, pred, gt, nullptr, nullptr, mask, dev);
        torch::NoGradGuard ng;
        auto m = DepthMetrics::compute(b.pred, b.gt, b.mask, min_d, max_d);
        for (int i = 0; i < 12; ++i) out12[i] = m.at(kEvalKeys[i]);
        if (counts) {
#ifdef CADL_DROPIN
            DepthMetrics::computeCounts(b.pred, b.gt, b.mask, min_d, max_d, counts);
#else
            ref_ops_eval_counts(b.pred, b.gt, b.mask, min_d, max_d, counts);
#endif
        }
        return 0;
    } catch (const std::exception& e) {
        return fail(err, errlen, e);
    }
}

// The trainers' computeDepthMetrics (tensorboard_trainer_enhanced.h:400-439) on one (1,H,W) sample
// or any flat set of N = B*H*W values: out7 = abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3.
int cadh_metrics_train(int B, int H, int W, int device, const float* pred, const float* gt,
                       float out7[7], int64_t counts[4], char* err, int errlen) {
    try {
        cadh_step_cfg c{};
        c.B = B; c.H = H; c.W = W; c.device = device;
        auto dev = pick_device(device);
        Batch b = to_device(c, pred, gt, nullptr, nullptr, nullptr, dev);
        torch::NoGradGuard ng;
        auto m = computeDepthMetrics(b.pred, b.gt);
        out7[0] = m.abs_rel; out7[1] = m.sq_rel; out7[2] = m.rmse; out7[3] = m.rmse_log;
        out7[4] = m.a1; out7[5] = m.a2; out7[6] = m.a3;
        if (counts) {
#ifdef CADL_DROPIN
            computeDepthMetricsCounts(b.pred, b.gt, counts);
#else
            ref_ops_train_counts(b.pred, b.gt, counts);
#endif
        }
        return 0;
    } catch (const std::exception& e) {
        return fail(err, errlen, e);
    }
}

// Timing loop over the same step:  per iteration
//   [h2d of rgb/gt/K/pred if include_h2d] zero grad -> forward -> backward [-> metrics] -> loss.item
// out_ms[i] = wall milliseconds of iteration i (device synchronised on both sides).
int cadh_time_steps(const cadh_step_cfg* cfg, const float* pred, const float* gt, const float* rgb,
                    const float* K, const uint8_t* mask, int with_metrics, int include_h2d,
                    int warmup, int iters, double* out_ms, float* out_last_loss, char* err,
                    int errlen) {
    try {
        const cadh_step_cfg& c = *cfg;
        auto dev = pick_device(c.device);
        Batch b = to_device(c, pred, gt, rgb, K, mask, dev);
        // pinned staging copies for the h2d variant
        Batch h;
        if (include_h2d) {
            auto pin = [&](const torch::Tensor& t) {
                auto cpu = t.to(torch::kCPU).contiguous();
                return dev.is_cuda() ? cpu.pin_memory() : cpu;
            };
            h.pred = pin(b.pred); h.gt = pin(b.gt);
            if (b.rgb.defined()) h.rgb = pin(b.rgb);
            if (b.K.defined()) h.K = pin(b.K);
        }
        CombinedDepthLoss combined(c.w_si, c.w_grad, c.w_smooth, c.w_reproj);
        float last = 0.f;
        for (int it = 0; it < warmup + iters; ++it) {
            sync_device(dev);
            auto t0 = std::chrono::steady_clock::now();
            if (include_h2d) {
                b.pred = h.pred.to(dev, /*non_blocking=*/true);
                b.gt = h.gt.to(dev, true);
                if (h.rgb.defined()) b.rgb = h.rgb.to(dev, true);
                if (h.K.defined()) b.K = h.K.to(dev, true);
            }
            b.pred = b.pred.detach();
            b.pred.set_requires_grad(true);       // optimizer_->zero_grad() equivalent: fresh leaf
            auto loss = run_term(c, combined, b);
            if (loss.requires_grad()) loss.sum().backward();
            if (with_metrics) {
                torch::NoGradGuard ng;
                auto pd = b.pred.detach();
                auto m1 = DepthMetrics::compute(pd, b.gt);
                auto m2 = computeDepthMetrics(pd, b.gt);
                last += 0.f * (m1.at("abs_rel") + m2.abs_rel);
            }
            last = loss.sum().item<float>();
            sync_device(dev);
            auto t1 = std::chrono::steady_clock::now();
            if (it >= warmup)
                out_ms[it - warmup] = std::chrono::duration<double, std::milli>(t1 - t0).count();
        }
        if (out_last_loss) *out_last_loss = last;
        return 0;
    } catch (const std::exception& e) {
        return fail(err, errlen, e);
    }
}

// Host-side metric utilities (depth_metrics.h:93-141 computePerSample / average, :259-304 MetricsAccumulator,
// :309-333 formatMetrics) driven the way an evaluation loop drives them: per-sample metrics of a batch, their average,
// a running accumulator fed with `splits` consecutive sub-batches, and the printed block.
//   out_per_sample: B x 12 floats (kEvalKeys order)     out_avg: 12     out_acc: 12 (accumulator average)
//   out_count: accumulator count after the updates and after reset()     out_text: formatMetrics(average)
int cadh_metric_utils(int B, int H, int W, int device, const float* pred, const float* gt, const uint8_t* mask,
                      int splits, float* out_per_sample, float* out_avg, float* out_acc, int* out_count2, char* out_text,
                      int textlen, char* err, int errlen) {
    try {
        cadh_step_cfg c{};
        c.B = B; c.H = H; c.W = W; c.device = device;
        auto dev = pick_device(device);
        Batch b = to_device(c, pred, gt, nullptr, nullptr, mask, dev);
        torch::NoGradGuard ng;
        auto per = DepthMetrics::computePerSample(b.pred, b.gt, b.mask);
        TORCH_CHECK((int)per.size() == B, "computePerSample returned ", per.size(), " maps for ", B, " samples");
        for (int i = 0; i < B; ++i)
            for (int k = 0; k < 12; ++k) out_per_sample[i * 12 + k] = per[i].at(kEvalKeys[k]);
        auto avg = DepthMetrics::average(per);
        for (int k = 0; k < 12; ++k) out_avg[k] = avg.at(kEvalKeys[k]);
        MetricsAccumulator acc;
        for (int sidx = 0; sidx < splits; ++sidx) {
            const int lo = (int)((long long)B * sidx / splits), hi = (int)((long long)B * (sidx + 1) / splits);
            if (hi <= lo) continue;
            torch::optional<torch::Tensor> m;
            if (b.mask.has_value()) m = b.mask.value().slice(0, lo, hi);
            acc.update(DepthMetrics::compute(b.pred.slice(0, lo, hi), b.gt.slice(0, lo, hi), m));
        }
        auto am = acc.average();
        for (int k = 0; k < 12; ++k) out_acc[k] = am.at(kEvalKeys[k]);
        out_count2[0] = acc.count();
        acc.reset();
        out_count2[1] = acc.count();
        const std::string text = formatMetrics(avg);
        if (out_text && textlen > 0) {
            std::strncpy(out_text, text.c_str(), textlen - 1);
            out_text[textlen - 1] = 0;
        }
        return 0;
    } catch (const std::exception& e) {
        return fail(err, errlen, e);
    }
}

#ifdef CADL_DROPIN
// ReprojectionLoss::forwardPhotometricWarp (opt-in extension) + backward: loss and dL/dpred
int cadh_photometric_step(int B, int H, int W, int device, const float* pred, const float* K, const float* T,
                          const float* src, const float* tgt, float upstream, float* out_loss, float* out_grad, char* err,
                          int errlen) {
    try {
        auto dev = pick_device(device);
        auto p = host_view(pred, {B, 1, H, W}).to(dev).clone().set_requires_grad(true);
        auto k = host_view(K, {B, 3, 3}).to(dev);
        auto t = host_view(T, {B, 4, 4}).to(dev);
        auto s_ = host_view(src, {B, 3, H, W}).to(dev);
        auto g_ = host_view(tgt, {B, 3, H, W}).to(dev);
        ReprojectionLoss l;
        auto loss = l.forwardPhotometricWarp(p, k, t, s_, g_);
        (loss * upstream).backward();
        *out_loss = loss.item<float>();
        auto gh = p.grad().to(torch::kCPU).contiguous();
        std::memcpy(out_grad, gh.data_ptr<float>(), sizeof(float) * gh.numel());
        return 0;
    } catch (const std::exception& e) {
        return fail(err, errlen, e);
    }
}

// Ranks of ScaleInvariantLoss / ReprojectionLoss on a batch (out4 = {SI default, SI opted in, reproj default, reproj
// opted in}: dim() of the returned tensor; the reference gives 1 when no pixel is valid, 0 otherwise)
int cadh_empty_rank(int B, int H, int W, int device, const float* pred, const float* gt, const float* K, int out4[4], char* err,
                    int errlen) {
    try {
        auto dev = pick_device(device);
        auto p = host_view(pred, {B, 1, H, W}).to(dev).clone().set_requires_grad(true);
        auto g = host_view(gt, {B, 1, H, W}).to(dev);
        auto k = host_view(K, {B, 3, 3}).to(dev);
        ScaleInvariantLoss si, si2;
        ReprojectionLoss rp, rp2;
        si2.referenceEmptyRank(true);
        rp2.referenceEmptyRank(true);
        out4[0] = (int)si.forward(p, g).dim();
        out4[1] = (int)si2.forward(p, g).dim();
        out4[2] = (int)rp.forward(p, g, k).dim();
        out4[3] = (int)rp2.forward(p, g, k).dim();
        return 0;
    } catch (const std::exception& e) {
        return fail(err, errlen, e);
    }
}

// RayDirectionComputer drop-in: rays of ONE camera on the device -> saveRayDirections -> loadRayDirections.
// out_rays: H*W*3 floats as loaded back; returns the save() result in *saved (0: the reference's "false" paths).
int cadh_rays_roundtrip(int device, int H, int W, const float* K9, const char* path, int corrupt_dims, float* out_rays,
                        int* saved, int* hw_out, char* err, int errlen) {
    try {
        auto dev = pick_device(device);
        auto K = host_view(K9, {3, 3}).to(dev);
        auto rays = RayDirectionComputer::computeRayDirections(K, H, W)[0];
        *saved = RayDirectionComputer::saveRayDirections(rays, corrupt_dims ? H + 1 : H, W, path) ? 1 : 0;
        if (!*saved) return 0;
        int h = 0, w = 0;
        auto back = RayDirectionComputer::loadRayDirections(path, h, w);
        hw_out[0] = h; hw_out[1] = w;
        std::memcpy(out_rays, back.data_ptr<float>(), sizeof(float) * (size_t)back.numel());
        return 0;
    } catch (const std::exception& e) {
        return fail(err, errlen, e);
    }
}

// "next" rows through their C++ wrappers (host/training/grad_clip.h, host/data/batch_prep.h)
int cadh_clip_grad_norm(int device, int count, const float* const* grads_host, const int64_t* sizes, float max_norm,
                        int do_clip, float* out2, float* const* grads_out_host, char* err, int errlen) {
    try {
        auto dev = pick_device(device);
        std::vector<torch::Tensor> params;
        for (int i = 0; i < count; ++i) {
            auto p = torch::zeros({sizes[i]}, torch::TensorOptions().device(dev)).set_requires_grad(true);
            p.mutable_grad() = torch::from_blob(const_cast<float*>(grads_host[i]), {sizes[i]}, torch::kFloat32).to(dev).clone();
            params.push_back(p);
        }
        FusedGradClipper clipper;
        auto out = clipper.run(params, max_norm, do_clip != 0).to(torch::kCPU);
        out2[0] = out[0].item<float>();
        out2[1] = out[1].item<float>();
        for (int i = 0; i < count; ++i) {
            auto g = params[i].grad().to(torch::kCPU).contiguous();
            std::memcpy(grads_out_host[i], g.data_ptr<float>(), sizeof(float) * g.numel());
        }
        return 0;
    } catch (const std::exception& e) {
        return fail(err, errlen, e);
    }
}

// DeviceAccumulator over `batches` loss values with weights: returns the weighted mean (one sync at the end)
int cadh_accumulate(int device, int batches, const float* values, const double* weights, double* mean_out, char* err,
                    int errlen) {
    try {
        auto dev = pick_device(device);
        DeviceAccumulator acc(1);
        for (int i = 0; i < batches; ++i) acc.add(torch::full({1}, values[i], torch::TensorOptions().device(dev)), weights[i]);
        *mean_out = acc.mean()[0];
        return 0;
    } catch (const std::exception& e) {
        return fail(err, errlen, e);
    }
}

int cadh_batch_prep(int device, int B, int h, int w, int H, int W, const float* rgb, const float* depth, const float* K,
                    float* rgb_out, float* depth_out, float* K_out, char* err, int errlen) {
    try {
        auto dev = pick_device(device);
        auto r = host_view(rgb, {B, 3, h, w}).to(dev);
        auto d = host_view(depth, {B, 1, h, w}).to(dev);
        auto k = host_view(K, {B, 3, 3}).to(dev);
        auto res = resizeBatchOnDevice(r, d, k, H, W);
        auto ro = std::get<0>(res).to(torch::kCPU).contiguous(), dd = std::get<1>(res).to(torch::kCPU).contiguous(),
             ko = std::get<2>(res).to(torch::kCPU).contiguous();
        std::memcpy(rgb_out, ro.data_ptr<float>(), sizeof(float) * ro.numel());
        std::memcpy(depth_out, dd.data_ptr<float>(), sizeof(float) * dd.numel());
        std::memcpy(K_out, ko.data_ptr<float>(), sizeof(float) * ko.numel());
        return 0;
    } catch (const std::exception& e) {
        return fail(err, errlen, e);
    }
}
#endif

}  // extern "C"

#include "unet_step.inc"
