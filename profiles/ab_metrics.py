"""Step time with 0/1/2/both metric variants fused, pyramid kernels beside phase A (default) vs in line (mode 32).
Usage (under gpurun): python profiles/ab_metrics.py"""
import importlib, os, sys, torch
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("camera-aware-neural-networks-for-few-view-depth-estimation_b200")
dev = torch.device("cuda:0")
B, H, W = 32, 480, 640
b = pkg.synth.make_batch(B, H, W, seed=1234, device=dev)
ws = pkg.Workspace(B, H, W, dev)
grad = torch.empty_like(b["pred"])
def run(mode, metrics, n=100):
    params = pkg.default_params(metrics=metrics)
    pkg.force_generic(mode)
    f = lambda: pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"], None, params=params, grad=grad, ws=ws)
    for _ in range(10): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); pkg.force_generic(0)
    return e0.elapsed_time(e1) / n * 1e3
run(0, 0)
for m in (0, 1, 2, 3):
    print(f"metrics={m}: overlap {run(0, m):6.1f} us   in-line {run(32, m):6.1f} us")
