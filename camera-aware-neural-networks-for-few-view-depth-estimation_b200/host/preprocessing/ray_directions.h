// Ray directions for the drop-in (reference: src/preprocessing/ray_direction_computer.{h,cpp}).
//
// The reference class computes unit rays K^-1 [u v 1]^T / |.| on the host with Eigen, one pixel at a time, and writes
// them to <frame>.bin; the loader reads that file back and reshapes it to (3,H,W) (src/data/sunrgbd_loader.cpp:329-350).
// Here the rays come from one kernel (cadl_rays_from_K) as a device tensor in either layout, and the .bin codec is kept
// byte for byte (ray_direction_computer.h:96-99: int32 H, int32 W, H*W*3 float32 row-major) with the reference's
// return / throw behaviour (.cpp:129-201), without the Eigen dependency: matrices are torch tensors.
#ifndef CADL_RAY_DIRECTIONS_H
#define CADL_RAY_DIRECTIONS_H

#include <torch/torch.h>

#include <cstdint>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>

#include "../cadl_torch.h"

namespace camera_aware_depth {

class RayDirectionComputer {
public:
    /// computeRayDirections (.cpp:17-62): K (3,3) or (B,3,3) on a CUDA device -> (B, H*W, 3) unit rays on that device
    static torch::Tensor computeRayDirections(const torch::Tensor& K, int height, int width) {
        return run(K, torch::Tensor(), height, width, 0);
    }
    /// computeRayDirectionsMaps (.cpp:64-101): the loader's layout (B, 3, H, W)
    static torch::Tensor computeRayDirectionsMaps(const torch::Tensor& K, int height, int width) {
        return run(K, torch::Tensor(), height, width, 1);
    }
    /// transformRaysToWorld (.cpp:103-127) fused into the generation: pose (B,4,4), rotation only, re-normalised
    static torch::Tensor computeWorldRayDirections(const torch::Tensor& K, const torch::Tensor& pose, int height, int width,
                                                   int layout = 0) {
        return run(K, pose, height, width, layout);
    }

    /// saveRayDirections (.cpp:129-168): rays (H*W,3) float32 on any device.  false (after printing the reference's
    /// messages to stderr) when the file cannot be opened or the dimensions do not match.
    static bool saveRayDirections(const torch::Tensor& rays, int height, int width, const std::string& filename) {
        std::ofstream file(filename, std::ios::binary);
        if (!file.is_open()) {
            std::cerr << "Error: Could not open file for writing: " << filename << std::endl;
            return false;
        }
        const int32_t h = height, w = width;
        file.write(reinterpret_cast<const char*>(&h), sizeof(int32_t));
        file.write(reinterpret_cast<const char*>(&w), sizeof(int32_t));
        if (rays.dim() != 2 || rays.size(0) != (int64_t)height * width || rays.size(1) != 3) {
            std::cerr << "Error: Ray dimensions mismatch. Expected " << (int64_t)height * width << "x3, got "
                      << (rays.dim() > 0 ? rays.size(0) : 0) << "x" << (rays.dim() > 1 ? rays.size(1) : 0) << std::endl;
            file.close();
            return false;
        }
        auto host = rays.to(torch::kCPU, torch::kFloat32).contiguous();      // one block write instead of 3*H*W calls
        file.write(reinterpret_cast<const char*>(host.data_ptr<float>()), sizeof(float) * (size_t)host.numel());
        file.close();
        return file.good() || true;
    }

    /// loadRayDirections (.cpp:170-201): (H*W,3) float32 CPU tensor; throws std::runtime_error if the file cannot be opened
    static torch::Tensor loadRayDirections(const std::string& filename, int& height, int& width) {
        std::ifstream file(filename, std::ios::binary);
        if (!file.is_open()) throw std::runtime_error("Error: Could not open file for reading: " + filename);
        int32_t h = 0, w = 0;
        file.read(reinterpret_cast<char*>(&h), sizeof(int32_t));
        file.read(reinterpret_cast<char*>(&w), sizeof(int32_t));
        height = h; width = w;
        auto rays = torch::zeros({(int64_t)h * w, 3}, torch::kFloat32);
        file.read(reinterpret_cast<char*>(rays.data_ptr<float>()), sizeof(float) * (size_t)rays.numel());
        file.close();
        return rays;
    }

private:
    static torch::Tensor run(const torch::Tensor& K, const torch::Tensor& pose, int H, int W, int layout) {
        using namespace cadl_detail;
        TORCH_CHECK(K.defined() && K.is_cuda() && K.scalar_type() == torch::kFloat32, "cadl: K must be a float32 CUDA tensor");
        TORCH_CHECK((K.dim() == 2 || K.dim() == 3) && K.size(-1) == 3 && K.size(-2) == 3, "cadl: K must be (3,3) or (B,3,3)");
        const int batched = K.dim() == 3 ? 1 : 0;
        int64_t B = batched ? K.size(0) : 1;
        torch::Tensor P;
        if (pose.defined()) {
            TORCH_CHECK(pose.is_cuda() && pose.scalar_type() == torch::kFloat32 && pose.dim() == 3 && pose.size(1) == 4 &&
                        pose.size(2) == 4, "cadl: pose must be a float32 (B,4,4) CUDA tensor");
            P = pose.contiguous();
            if (!batched) B = P.size(0);
            TORCH_CHECK(P.size(0) == B, "cadl: K and pose differ in batch size");
        }
        const auto dev = K.device();
        c10::cuda::CUDAGuard guard(dev);
        auto Kc = K.contiguous();
        auto out = layout == 0 ? torch::empty({B, (int64_t)H * W, 3}, Kc.options()) : torch::empty({B, 3, H, W}, Kc.options());
        int rc = cadl_rays_from_K(Kc.data_ptr<float>(), batched, P.defined() ? P.data_ptr<float>() : nullptr, (int)B, H, W,
                                  layout, out.data_ptr<float>(), current_stream(dev));
        check_rc(rc, "cadl_rays_from_K");
        return out;
    }
};

}  // namespace camera_aware_depth

#endif  // CADL_RAY_DIRECTIONS_H
