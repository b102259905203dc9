"""Config-3 step time (with and without the metric variants); [CADL_LIB=...] python profiles/r02_step.py"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("camera-aware-neural-networks-for-few-view-depth-estimation_b200")
dev = torch.device("cuda:0")
tag = os.path.basename(os.environ.get("CADL_LIB", "libcadl.so"))
B, H, W = int(os.environ.get("BATCH", "32")), 480, 640
b = pkg.synth.make_batch(B, H, W, seed=1234, device=dev)
ws = pkg.Workspace(B, H, W, dev)
grad = torch.empty_like(b["pred"])
out = []
for metrics in (3, 0):
    params = pkg.default_params(metrics=metrics)
    fn = lambda: pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"], None, params=params, grad=grad, ws=ws)
    for _ in range(30):
        fn()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(200):
            fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 200 * 1e3)
    out.append(f"metrics={metrics}: {best:6.1f} us/step")
print(f"[{tag}] " + "   ".join(out))
