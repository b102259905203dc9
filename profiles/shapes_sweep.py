"""Step time of the full stack (+ both metric variants) on the other BASELINE shapes, default dispatch vs the tile
kernel (mode 8) vs pyramid in line (mode 32), and config 5 (reprojection + metrics at 960x1280).
Usage (under gpurun): python profiles/shapes_sweep.py"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("camera-aware-neural-networks-for-few-view-depth-estimation_b200")
dev = torch.device("cuda:0")


def timed(fn, n=100):
    for _ in range(10):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for (B, H, W) in [(4, 240, 320), (32, 240, 320), (32, 480, 640), (16, 960, 1280), (8, 480, 640), (64, 480, 640)]:
    b = pkg.synth.make_batch(B, H, W, seed=1, device=dev)
    ws = pkg.Workspace(B, H, W, dev)
    grad = torch.empty_like(b["pred"])
    p = pkg.default_params(metrics=3)
    out = []
    for mode in (0, 32, 8):
        pkg.force_generic(mode)
        us = timed(lambda: pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"], None, params=p, grad=grad, ws=ws))
        out.append(f"mode {mode}: {us:7.1f} us = {B * H * W / us:8.0f} Mpix/s")
    pkg.force_generic(0)
    p5 = pkg.default_params(terms=pkg.TERM_REPROJ, w_reproj=1.0, metrics=3)
    us5 = timed(lambda: pkg.stack_fwd_bwd(b["pred"], b["gt"], None, b["K"], None, params=p5, grad=grad, ws=ws))
    print(f"B={B:3d} {H}x{W}: " + "  ".join(out) + f"   | reproj+metrics {us5:6.1f} us = {B * H * W / us5:8.0f} Mpix/s")
