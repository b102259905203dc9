// On-device batch preparation: SunRGBDLoader::resizeSample for a whole stacked batch
// (reference src/data/sunrgbd_loader.cpp:445-489: rgb bilinear align_corners=false, depth nearest, K rescaled)
// as ONE kernel after the H2D copy instead of three host interpolate calls per sample.
#ifndef CADL_BATCH_PREP_H
#define CADL_BATCH_PREP_H

#include <torch/torch.h>

#include <tuple>

#include "../cadl_torch.h"

namespace camera_aware_depth {

/// rgb (B,3,h,w), depth (B,1,h,w), K (B,3,3) on a CUDA device -> the same at (H,W) with K rescaled.
inline std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> resizeBatchOnDevice(torch::Tensor rgb, torch::Tensor depth,
                                                                                     torch::Tensor K, int64_t H, int64_t W) {
    using namespace cadl_detail;
    rgb = as_input(rgb, "rgb", 3);
    depth = as_input(depth, "depth", 1);
    TORCH_CHECK(K.is_cuda() && K.scalar_type() == torch::kFloat32 && K.dim() == 3 && K.size(0) == rgb.size(0),
                "cadl: intrinsics must be (B,3,3) float32 on the device");
    K = K.contiguous();
    const auto dev = rgb.device();
    c10::cuda::CUDAGuard guard(dev);
    const int64_t B = rgb.size(0), h = rgb.size(2), w = rgb.size(3);
    auto rgb_o = torch::empty({B, 3, H, W}, rgb.options());
    auto dep_o = torch::empty({B, 1, H, W}, rgb.options());
    auto K_o = torch::empty({B, 3, 3}, rgb.options());
    int rc = cadl_batch_prep(rgb.data_ptr<float>(), depth.data_ptr<float>(), K.data_ptr<float>(), (int)B, (int)h, (int)w,
                             (int)H, (int)W, rgb_o.data_ptr<float>(), dep_o.data_ptr<float>(), K_o.data_ptr<float>(),
                             current_stream(dev));
    check_rc(rc, "cadl_batch_prep");
    return {rgb_o, dep_o, K_o};
}

/// augmentSample (crop / horizontal flip / colour jitter, src/data/sunrgbd_loader.cpp:352-384) + the resize that
/// follows it, for a stacked batch in one launch.  aug: (B, CADL_AUG_STRIDE) float32, drawn by the loader's RNG on the
/// host (include/cadl.h: cadl_batch_augment), any device (copied if needed).
inline std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> augmentBatchOnDevice(torch::Tensor rgb, torch::Tensor depth,
                                                                                      torch::Tensor K, torch::Tensor aug,
                                                                                      int64_t H, int64_t W) {
    using namespace cadl_detail;
    rgb = as_input(rgb, "rgb", 3);
    depth = as_input(depth, "depth", 1);
    TORCH_CHECK(K.is_cuda() && K.scalar_type() == torch::kFloat32 && K.dim() == 3 && K.size(0) == rgb.size(0),
                "cadl: intrinsics must be (B,3,3) float32 on the device");
    TORCH_CHECK(aug.dim() == 2 && aug.size(0) == rgb.size(0) && aug.size(1) == CADL_AUG_STRIDE,
                "cadl: aug must be (B, CADL_AUG_STRIDE)");
    K = K.contiguous();
    const auto dev = rgb.device();
    aug = aug.to(dev, torch::kFloat32).contiguous();
    c10::cuda::CUDAGuard guard(dev);
    const int64_t B = rgb.size(0), h = rgb.size(2), w = rgb.size(3);
    auto rgb_o = torch::empty({B, 3, H, W}, rgb.options());
    auto dep_o = torch::empty({B, 1, H, W}, rgb.options());
    auto K_o = torch::empty({B, 3, 3}, rgb.options());
    int rc = cadl_batch_augment(rgb.data_ptr<float>(), depth.data_ptr<float>(), K.data_ptr<float>(), aug.data_ptr<float>(),
                                (int)B, (int)h, (int)w, (int)H, (int)W, rgb_o.data_ptr<float>(), dep_o.data_ptr<float>(),
                                K_o.data_ptr<float>(), current_stream(dev));
    check_rc(rc, "cadl_batch_augment");
    return {rgb_o, dep_o, K_o};
}

}  // namespace camera_aware_depth

#endif  // CADL_BATCH_PREP_H
