"""Synthetic SUN RGB-D-shaped inputs (SURVEY.md section 8d / BASELINE.md section 4).

Generated on the CPU with a seeded torch generator so the same tensors feed the oracle, the CUDA
kernels and the reference arm on any box; callers move them to the device.
"""
from __future__ import annotations

import math

import torch


def make_batch(B: int, H: int, W: int, seed: int = 1234, hole_frac: float = 0.15, device="cpu", with_rgb: bool = True):
    """Returns dict(pred, gt, rgb, K, T): fp32, NCHW; gt has ~15 % zeros (invalid depth)."""
    g = torch.Generator().manual_seed(seed)
    gt = torch.empty(B, 1, H, W).uniform_(0.3, 9.8, generator=g)
    holes = torch.rand(B, 1, H, W, generator=g) < hole_frac
    noise = torch.randn(B, 1, H, W, generator=g)
    pred = torch.clamp(gt * torch.exp(0.25 * noise), 0.05, 9.99)
    pred_h = torch.empty(B, 1, H, W).uniform_(0.05, 9.95, generator=g)
    pred = torch.where(holes, pred_h, pred)
    gt = torch.where(holes, torch.zeros_like(gt), gt)
    rgb = torch.rand(B, 3, H, W, generator=g) if with_rgb else torch.zeros(B, 3, 1, 1)      # (the big shapes that need no image)
    # Kinect-v1-like intrinsics rescaled to H x W (src/data/sunrgbd_loader.cpp:480-488), +-5 % jitter
    jit = 1.0 + 0.05 * (2.0 * torch.rand(B, 4, generator=g) - 1.0)
    K = torch.zeros(B, 3, 3)
    K[:, 0, 0] = 518.8579 * W / 640.0 * jit[:, 0]
    K[:, 1, 1] = 519.4696 * H / 480.0 * jit[:, 1]
    K[:, 0, 2] = (W - 1) / 2.0 * jit[:, 2]
    K[:, 1, 2] = (H - 1) / 2.0 * jit[:, 3]
    K[:, 2, 2] = 1.0
    # extrinsics: <= 10 degree tilt about x, zero translation (SUN RGB-D stores gravity alignment only)
    ang = (torch.rand(B, generator=g) * 2 - 1) * (10.0 * math.pi / 180.0)
    T = torch.eye(4).repeat(B, 1, 1)
    T[:, 1, 1] = torch.cos(ang); T[:, 1, 2] = -torch.sin(ang)
    T[:, 2, 1] = torch.sin(ang); T[:, 2, 2] = torch.cos(ang)
    out = {"pred": pred.contiguous(), "gt": gt.contiguous(), "rgb": rgb.contiguous(), "K": K.contiguous(),
           "T": T.contiguous()}
    return {k: v.to(device) for k, v in out.items()}


def make_smooth_batch(B: int, H: int, W: int, seed: int = 4321, device="cpu"):
    """Parity-only set: low-frequency depth ramps + 5 % noise with saturated regions where pred == 9.99
    exactly and flat regions where pred == gt exactly -- exercises sign(0) and the clamp edges."""
    g = torch.Generator().manual_seed(seed)
    yy = torch.linspace(0, 1, H).view(1, 1, H, 1)
    xx = torch.linspace(0, 1, W).view(1, 1, 1, W)
    a = torch.rand(B, 1, 1, 1, generator=g) * 4 + 1
    b = torch.rand(B, 1, 1, 1, generator=g) * 4 + 1
    gt = (a * yy + b * xx + 0.5).expand(B, 1, H, W).clone()
    pred = gt * (1.0 + 0.05 * torch.randn(B, 1, H, W, generator=g))
    pred = torch.clamp(pred, 0.05, 9.99)
    pred[:, :, : H // 4, : W // 4] = 9.99                       # saturated block
    pred[:, :, H // 2:, W // 2:] = gt[:, :, H // 2:, W // 2:]    # exact-match block
    gt[:, :, H // 3: H // 3 + max(1, H // 8), :] = 0.0           # a band of holes
    rgb = (0.5 + 0.5 * torch.sin(6.0 * xx + 3.0 * yy)).expand(B, 3, H, W).clone()
    rgb = torch.clamp(rgb + 0.02 * torch.randn(B, 3, H, W, generator=g), 0, 1)
    base = make_batch(B, H, W, seed=seed + 1)
    out = {"pred": pred.contiguous(), "gt": gt.contiguous(), "rgb": rgb.contiguous(), "K": base["K"], "T": base["T"]}
    return {k: v.to(device) for k, v in out.items()}
