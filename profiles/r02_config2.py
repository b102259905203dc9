"""Config 2 (reprojection alone) and config-5 shape timing; [CADL_LIB=...] python profiles/r02_config2.py"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("camera-aware-neural-networks-for-few-view-depth-estimation_b200")
dev = torch.device("cuda:0")
tag = os.path.basename(os.environ.get("CADL_LIB", "libcadl.so"))


def timeit(fn, n=200):
    for _ in range(20):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for (B, H, W, metrics) in ((32, 480, 640, 0), (16, 960, 1280, 3)):
    b = pkg.synth.make_batch(B, H, W, seed=5, device=dev, with_rgb=False)
    ws = pkg.Workspace(B, H, W, dev)
    grad = torch.empty_like(b["pred"])
    p = pkg.default_params(terms=pkg.TERM_REPROJ, w_reproj=1.0, metrics=metrics)
    us = timeit(lambda: pkg.stack_fwd_bwd(b["pred"], b["gt"], None, b["K"], None, params=p, grad=grad, ws=ws))
    px = B * H * W
    print(f"[{tag}] reproj B={B} {H}x{W} metrics={metrics}: {us:7.1f} us/step  {px / us:9.0f} Mpix/s  "
          f"{12 * px / (us * 1e-6) / 1e9 / 6452.5:.3f} of the 12 B/px roofline")
