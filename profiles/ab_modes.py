"""A/B of dispatch modes (cadl_debug_force_generic bit mask) on the config-3 step, alternating in one process.
Usage (under gpurun):  python profiles/ab_modes.py 0 16 [8 1 ...]"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("camera-aware-neural-networks-for-few-view-depth-estimation_b200")
modes = [int(x) for x in sys.argv[1:]] or [0, 16]
dev = torch.device("cuda:0")
B, H, W = 32, 480, 640
b = pkg.synth.make_batch(B, H, W, seed=1234, device=dev)
ws = pkg.Workspace(B, H, W, dev)
grad = torch.empty_like(b["pred"])
params = pkg.default_params(metrics=3)


def run(mode, n=100):
    pkg.force_generic(mode)
    for _ in range(10):
        pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"], None, params=params, grad=grad, ws=ws)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"], None, params=params, grad=grad, ws=ws)
    e1.record()
    torch.cuda.synchronize()
    pkg.force_generic(0)
    return e0.elapsed_time(e1) / n * 1e3


run(modes[0])
for rep in range(3):
    print("  ".join(f"mode {m}: {run(m):7.1f} us/step" for m in modes))
