"""GPU tests of the two rows next to the loss path: the ray generator (vs the plain-C oracle) and the
photometric-reprojection extension (vs oracle/oracle_torch.py in fp64; parity unpinned by the reference,
whose forwardPhotometric is a stub)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, check_grad, rel_err

pytestmark = pytest.mark.gpu
RAYS_SO = os.path.join(ROOT, "oracle", "liboracle_rays.so")


@pytest.fixture(scope="module")
def rays_lib():
    if not os.path.exists(RAYS_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "rays"])
    L = C.CDLL(RAYS_SO)
    L.oracle_rays_hw3.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.oracle_rays_3hw.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.oracle_rays_to_world.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    return L


def _ulp_diff(a, b):
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    return np.abs(ai - bi)


@pytest.mark.parametrize("H,W", [(480, 640), (37, 53)])
def test_rays_match_c_oracle(pkg, rays_lib, H, W):
    b = pkg.synth.make_batch(3, H, W, seed=17)
    K = b["K"]
    d = torch.device("cuda:0")
    hw3 = pkg.rays_from_K(K.to(d), H, W, layout=0).cpu().numpy()
    planar = pkg.rays_from_K(K.to(d), H, W, layout=1).cpu().numpy()
    for i in range(3):
        Ki = np.ascontiguousarray(K[i].numpy())
        ref = np.empty((H * W, 3), dtype=np.float32)
        rays_lib.oracle_rays_hw3(Ki.ctypes.data, H, W, ref.ctypes.data)
        assert _ulp_diff(hw3[i], ref).max() == 0            # same operations, no contraction: bit-exact
        assert np.array_equal(planar[i].reshape(3, -1).T, hw3[i])
    # (3,3) broadcast K and world transform
    pose = b["T"][:1].contiguous()
    w = pkg.rays_from_K(K[0].contiguous().to(d), H, W, layout=0, pose=pose.to(d)).cpu().numpy()[0]
    refw = np.empty_like(ref)
    Ki = np.ascontiguousarray(K[0].numpy())
    rays_lib.oracle_rays_hw3(Ki.ctypes.data, H, W, ref.ctypes.data)
    P = np.ascontiguousarray(pose[0].numpy())
    rays_lib.oracle_rays_to_world(ref.ctypes.data, H * W, P.ctypes.data, refw.ctypes.data)
    assert _ulp_diff(w, refw).max() <= 2
    assert np.allclose(np.linalg.norm(w, axis=1), 1.0, atol=1e-6)


def test_rays_bin_roundtrip_from_device(pkg, tmp_path):
    H, W = 48, 64
    b = pkg.synth.make_batch(1, H, W, seed=3)
    r = pkg.rays_from_K(b["K"].cuda(), H, W, layout=0).cpu().numpy()[0]
    f = str(tmp_path / "rays.bin")
    assert pkg.save_ray_directions(r, H, W, f)
    back, h, w = pkg.load_ray_directions(f)
    assert (h, w) == (H, W) and np.array_equal(back, r)
    # loader layout (3,H,W): src/data/sunrgbd_loader.cpp:345-347
    planar = pkg.rays_from_K(b["K"].cuda(), H, W, layout=1).cpu().numpy()[0]
    assert np.array_equal(back.reshape(H, W, 3).transpose(2, 0, 1), planar)


def test_rays_cpp_codec_matches_the_python_codec_and_the_reference_format(pkg, tmp_path):
    """host/preprocessing/ray_directions.h (the drop-in RayDirectionComputer): device rays -> saveRayDirections ->
    loadRayDirections in C++; the file is byte-identical to the Python codec's and to the format of
    ray_direction_computer.h:96-99 (int32 H, int32 W, H*W*3 float32); a dimension mismatch returns false (.cpp:146-152)."""
    H, W = 48, 64
    b = pkg.synth.make_batch(1, H, W, seed=3)
    host = pkg.host_harness()
    f_cpp, f_py = str(tmp_path / "cpp.bin"), str(tmp_path / "py.bin")
    ok, back, hw = host.rays_roundtrip(0, b["K"][0].numpy(), H, W, f_cpp)
    assert ok and hw == (H, W)
    dev_rays = pkg.rays_from_K(b["K"].cuda(), H, W, layout=0).cpu().numpy()[0]
    assert np.array_equal(back, dev_rays)
    assert pkg.save_ray_directions(dev_rays, H, W, f_py)
    assert open(f_cpp, "rb").read() == open(f_py, "rb").read()
    raw = open(f_cpp, "rb").read()
    assert len(raw) == 8 + H * W * 12 and np.frombuffer(raw[:8], "<i4").tolist() == [H, W]
    ok, _, _ = host.rays_roundtrip(0, b["K"][0].numpy(), H, W, str(tmp_path / "bad.bin"), corrupt_dims=True)
    assert not ok


@pytest.mark.parametrize("W", [64, 52, 50])       # 128-bit path, 4-aligned odd count of quads, scalar path
def test_rays_layouts_agree_for_any_width(pkg, W):
    H = 24
    b = pkg.synth.make_batch(2, H, W, seed=9)
    r0 = pkg.rays_from_K(b["K"].cuda(), H, W, layout=0).cpu()
    r1 = pkg.rays_from_K(b["K"].cuda(), H, W, layout=1).cpu()
    assert torch.equal(r0.view(2, H, W, 3).permute(0, 3, 1, 2), r1)
    assert float((r0.norm(dim=-1) - 1).abs().max()) < 1e-6


def _smooth_images(B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    yy = torch.linspace(0, 1, H).view(1, 1, H, 1)
    xx = torch.linspace(0, 1, W).view(1, 1, 1, W)
    ph = torch.rand(B, 3, 1, 1, generator=g) * 6.28
    src = 0.5 + 0.4 * torch.sin(9.0 * xx + 5.0 * yy + ph)
    tgt = 0.5 + 0.4 * torch.sin(9.0 * xx + 5.0 * yy + ph + 0.3)
    return src.contiguous(), tgt.contiguous()


def _photometric_case(pkg, B, H, W, seed):
    b = pkg.synth.make_batch(B, H, W, seed=seed)
    src, tgt = _smooth_images(B, H, W, seed)
    depth = (2.0 + b["pred"] * 0.3).contiguous()
    T = b["T"].clone()
    T[:, 0, 3] = 0.05           # a baseline on top of the <= 10 degree tilt
    return b, src, tgt, depth, T


def _photometric_ties(oracle, depth, K, T, src, tgt, H, W, eps=1e-6):
    """Pixels where the result is discontinuous in the inputs -- the sample point within 2e-4 px of a texel boundary
    (the bilinear cell changes), of the image border (the inside test flips) or Z_s within 1e-5 of eps: fp32 rounding
    decides there, so they are excluded and counted, like the sign ties of the stencil terms."""
    d = depth.double()[:, 0]
    B = d.shape[0]
    v = torch.arange(H, dtype=torch.float64).view(1, H, 1)
    u = torch.arange(W, dtype=torch.float64).view(1, 1, W)
    Kd, Td = K.double(), T.double()
    X = (u - Kd[:, 0, 2].view(B, 1, 1)) / Kd[:, 0, 0].view(B, 1, 1) * d
    Y = (v - Kd[:, 1, 2].view(B, 1, 1)) / Kd[:, 1, 1].view(B, 1, 1) * d
    P = [Td[:, i, 0].view(B, 1, 1) * X + Td[:, i, 1].view(B, 1, 1) * Y + Td[:, i, 2].view(B, 1, 1) * d + Td[:, i, 3].view(B, 1, 1)
         for i in range(3)]
    us = Kd[:, 0, 0].view(B, 1, 1) * P[0] / P[2] + Kd[:, 0, 2].view(B, 1, 1)
    vs = Kd[:, 1, 1].view(B, 1, 1) * P[1] / P[2] + Kd[:, 1, 2].view(B, 1, 1)
    near = lambda x: (x - torch.round(x)).abs() < 1e-4
    tie = near(us) | near(vs) | ((P[2] - eps).abs() < 1e-5)
    # ... or a channel's residual within 2e-6 of zero (the L1 term's sign decides the gradient)
    gx, gy = (2.0 * us + 1.0) / W - 1.0, (2.0 * vs + 1.0) / H - 1.0
    warped = torch.nn.functional.grid_sample(src.double(), torch.stack([gx, gy], -1), mode="bilinear", padding_mode="zeros",
                                             align_corners=False)
    tie = tie | ((warped - tgt.double()).abs() < 2e-6).any(1)
    return tie.unsqueeze(1)


@pytest.mark.parametrize("shape,seed", [((2, 96, 128), 5), ((3, 120, 160), 6), ((1, 240, 320), 7)])
def test_photometric_extension_vs_grid_sample(pkg, oracle, shape, seed):
    """The opt-in photometric warp against the restatement built on F.grid_sample -- in fp32 on CUDA (the tight
    oracle: same sampling arithmetic) at 1e-5, and in fp64 on the CPU -- with boundary-tie pixels excluded and counted.
    Through the C ABI and through the C++ method ReprojectionLoss::forwardPhotometricWarp."""
    B, H, W = shape
    b, src, tgt, depth, T = _photometric_case(pkg, B, H, W, seed)
    d = torch.device("cuda:0")
    ws = pkg.photometric_fwd_bwd(depth.to(d), b["K"].to(d), T.to(d), src.to(d), tgt.to(d))
    torch.cuda.synchronize()
    r = pkg.results_dict(ws.read_results())
    assert r["n_reproj"] > 0.5 * B * H * W
    tie = _photometric_ties(oracle, depth, b["K"], T, src, tgt, H, W)
    for dtype, dev, tol in ((torch.float32, d, 1e-5), (torch.float64, torch.device("cpu"), 1e-5)):
        p = depth.to(device=dev, dtype=dtype).requires_grad_(True)
        loss = oracle.photometric_reprojection(p, b["K"].to(dev, dtype), T.to(dev, dtype), src.to(dev, dtype), tgt.to(dev, dtype))
        loss.sum().backward()
        assert rel_err(r["reproj_loss"], float(loss)) <= tol, (dtype, r["reproj_loss"], float(loss))
        n_excl = check_grad(ws.grad.cpu(), p.grad.detach().cpu().float(), tie, tol, f"photometric {dtype}")
    print(f"photometric {shape}: {int(tie.sum())} boundary-tie pixels excluded of {tie.numel()}")
    # the C++ method (autograd node + backward kernel), upstream 2.5
    host = pkg.host_harness()
    loss_c, grad_c = host.photometric_step(0, depth.numpy(), b["K"].numpy(), T.numpy(), src.numpy(), tgt.numpy(), upstream=2.5)
    assert rel_err(loss_c, r["reproj_loss"]) <= 1e-6
    assert torch.allclose(torch.from_numpy(grad_c), 2.5 * ws.grad.cpu(), rtol=1e-6, atol=0)


def test_photometric_stub_is_kept(pkg, host_stub=None):
    """The drop-in keeps the reference's stub behaviour for forwardPhotometric (zeros(1)); the real warp is the
    opt-in forwardPhotometricWarp / cadl_photometric_fwd_bwd."""
    import re
    src = open(os.path.join(ROOT, pkg.__name__.split(".")[0], "host", "loss", "depth_loss.h")).read()
    body = re.search(r"forwardPhotometric\(.*?\{(.*?)\n    \}", src, flags=re.S).group(1)
    assert "torch::zeros(1" in body
    assert "forwardPhotometricWarp(" in src


# ---------------------------------------------------------------------------------------------------------
# "next" rows: on-device batch prep, fused grad-norm clip
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("src,dst", [((530, 730), (480, 640)), ((427, 561), (240, 320)), ((240, 320), (480, 640))])
def test_batch_prep_matches_aten_interpolate(pkg, oracle, src, dst):
    (h, w), (H, W) = src, dst
    g = torch.Generator().manual_seed(3)
    rgb = torch.rand(3, 3, h, w, generator=g)
    depth = torch.rand(3, 1, h, w, generator=g) * 9 + 0.3
    depth[torch.rand(3, 1, h, w, generator=g) < 0.15] = 0.0
    K = pkg.synth.make_batch(3, h, w, seed=2)["K"]
    d = torch.device("cuda:0")
    r2, d2, K2 = pkg.batch_prep(rgb.to(d), depth.to(d), K.to(d), H, W)
    for dev in ("cpu", "cuda"):
        ro, do, Ko = oracle.resize_sample(rgb.to(dev), depth.to(dev), K.to(dev), H, W)
        assert torch.equal(d2.cpu(), do.cpu())                       # nearest: identical source indices
        assert float((r2.cpu() - ro.cpu()).abs().max()) <= 2e-6      # bilinear weights to rounding
        assert torch.equal(K2.cpu(), Ko.cpu())


def test_clip_grad_norm_matches_torch(pkg, oracle):
    d = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    shapes = [(64, 3, 3, 3), (64,), (128, 64, 3, 3), (1,), (4097,), (512, 512, 3, 3), (7, 5)]
    for scale, max_norm in ((1.0, 1.0), (1e-4, 1.0), (3.0, 0.5)):
        grads = [(torch.randn(*s, generator=g) * scale).to(d) for s in shapes]
        # the reference op itself on parameters carrying these gradients
        params = [torch.nn.Parameter(torch.zeros_like(x)) for x in grads]
        for prm, x in zip(params, grads):
            prm.grad = x.clone()
        total_ref = torch.nn.utils.clip_grad_norm_(params, max_norm)
        tot_o, clipped_o = oracle.clip_grad_norm([x.clone() for x in grads], max_norm)
        clipper = pkg.GradClipper(grads)
        out = clipper(max_norm).cpu()
        assert rel_err(float(out[0]), float(tot_o)) <= 1e-6 and rel_err(float(out[0]), float(total_ref)) <= 1e-5
        coef = min(max_norm / (float(tot_o) + 1e-6), 1.0)
        assert rel_err(float(out[1]), coef) <= 1e-6
        for a, b, prm in zip(grads, clipped_o, params):
            assert float((a - b).abs().max()) <= 1e-6 * max(float(b.abs().max()), 1e-30)
            assert float((a - prm.grad).abs().max()) <= 1e-5 * max(float(prm.grad.abs().max()), 1e-30)
    # norm only (computeGradientNorm): gradients untouched
    grads = [torch.randn(1000, generator=g).to(d)]
    keep = grads[0].clone()
    out = pkg.GradClipper(grads)(0.1, clip=False).cpu()
    assert torch.equal(grads[0], keep) and rel_err(float(out[0]), float(keep.double().norm())) <= 1e-6


def test_batch_augment_matches_loader_restatement(pkg, oracle):
    """augmentSample (crop, flip, colour jitter) + the resize after it, per image, against the op-for-op restatement."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    B, h, w, H, W = 8, 48, 64, 48, 64
    rgb = torch.rand(B, 3, h, w, generator=g)
    depth = torch.rand(B, 1, h, w, generator=g) * 9 + 0.3
    K = torch.tensor([[55.0, 0, 31.5], [0, 56.0, 23.5], [0, 0, 1]]).repeat(B, 1, 1) * (1 + 0.01 * torch.arange(B).view(B, 1, 1))
    K[:, 2, 2] = 1
    aug = torch.tensor([
        [0, 0, 0, 0, 0, 0, 1.0, 1.0],            # identity
        [5, 3, 51, 38, 0, 0, 1.0, 1.0],          # crop (scale 0.8)
        [0, 0, 0, 0, 1, 0, 1.0, 1.0],            # flip
        [0, 0, 0, 0, 0, 1, 1.13, 0.91],          # colour jitter
        [9, 7, 44, 33, 1, 1, 0.87, 1.08],        # all three
        [12, 9, 38, 28, 1, 0, 1.0, 1.0],
        [1, 0, 64, 48, 1, 0, 1.0, 1.0],          # window overshoots by one column: clamped like the reference's Slice (:395-397)
        [60, 44, 30, 30, 0, 0, 1.0, 1.0],        # overshoots in both directions
    ], dtype=torch.float32)
    ro, do, Ko = pkg.batch_augment(rgb.to(dev), depth.to(dev), K.to(dev), aug.to(dev), H, W)
    for b in range(B):
        r, d, k = oracle.augment_resize_sample(rgb[b], depth[b], K[b], aug[b].tolist(), H, W)
        assert float((ro[b].cpu() - r).abs().max()) <= 2e-6, b
        assert torch.equal(do[b].cpu(), d), b
        assert torch.allclose(Ko[b].cpu(), k, rtol=1e-6, atol=0), (b, Ko[b].cpu(), k)
    # no augmentation == cadl_batch_prep
    r0, d0, K0 = pkg.batch_prep(rgb.to(dev), depth.to(dev), K.to(dev), 40, 56)
    z = torch.zeros(B, 8); z[:, 6:] = 1
    r1, d1, K1 = pkg.batch_augment(rgb.to(dev), depth.to(dev), K.to(dev), z.to(dev), 40, 56)
    assert torch.equal(r0, r1) and torch.equal(d0, d1) and torch.equal(K0, K1)


def test_device_accumulator(pkg):
    """cadl_accumulate: batch-size-weighted running sums without a host sync (production_trainer.h:213-216)."""
    dev = torch.device("cuda:0")
    acc = torch.zeros(4, dtype=torch.float64, device=dev)
    vals = [torch.tensor([0.5, 1.25, 3.0]), torch.tensor([0.25, 2.0, 1.0]), torch.tensor([1.5, 0.75, 2.5])]
    ws = [32, 32, 7]
    for v, w in zip(vals, ws):
        pkg.accumulate(v.to(dev), w, acc)
    ref = sum(w * v.double() for v, w in zip(vals, ws))
    out = acc.cpu()
    assert torch.equal(out[:3], ref) and float(out[3]) == float(sum(ws))
