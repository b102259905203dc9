"""Round-2 A/B aid for tuning builds of the gradient pass: config-3 step time, the gradient kernel alone (as bench.py
times it) and agreement with the round-1 streaming kernel of the in-tree debug library, one line per library.
Usage (under gpurun):  CADL_LIB=<pkg>/csrc/variants/libcadl_<tag>.so python profiles/r02_ab.py"""
import importlib
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("camera-aware-neural-networks-for-few-view-depth-estimation_b200")
dev = torch.device("cuda:0")
B, H, W = 32, 480, 640
b = pkg.synth.make_batch(B, H, W, seed=1234, device=dev)
ws = pkg.Workspace(B, H, W, dev)
grad = torch.empty_like(b["pred"])
tag = os.path.basename(os.environ.get("CADL_LIB", "libcadl.so"))


def step(params):
    pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"], None, params=params, grad=grad, ws=ws)


def timeit(fn, n=300, warm=30):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


params = pkg.default_params(metrics=3)
pkg.force_generic(128)      # round-1 streaming kernel + finish kernel (debug library)
step(params)
torch.cuda.synchronize()
g_old, r_old = grad.clone(), pkg.results_dict(ws.read_results())
pkg.force_generic(0)
grad.zero_()
step(params)
torch.cuda.synchronize()
r_new = pkg.results_dict(ws.read_results())
d = (grad - g_old).abs().max().item() / g_old.abs().max().item()
t_step = timeit(lambda: step(params))
# the same step through the debug library with the loss statistics from phase A (mode 128), alternating with the product
# library: step times drift by a few us between processes, the comparison within one process does not
alt = {"new": [], "phaseA": []}
for rnd in range(4):
    alt["new"].append(timeit(lambda: step(params), n=200, warm=10))
    pkg.force_generic(128)
    alt["phaseA"].append(timeit(lambda: step(params), n=200, warm=10))
    pkg.force_generic(0)
print(f"[{tag}] metrics=3 step, alternating: product " + " ".join(f"{x:.1f}" for x in alt["new"]) + "   phase-A statistics (debug library, mode 128) "
      + " ".join(f"{x:.1f}" for x in alt["phaseA"]), flush=True)
p0 = pkg.default_params(metrics=0)
t_step0 = timeit(lambda: step(p0))
ks = []
for it in range(8):
    pkg.stack_prepare(b["pred"], b["gt"], params, ws)
    pkg.stack_reduce(b["pred"], b["gt"], None, params, ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        pkg.stack_grad(b["pred"], b["gt"], b["rgb"], b["K"], None, params, grad, ws)
    e1.record()
    torch.cuda.synchronize()
    if it >= 2:
        ks.append(e0.elapsed_time(e1) / 20 * 1e3)
print(f"[{tag}] step {t_step:6.1f} us (no metrics {t_step0:6.1f})  grad kernel {statistics.median(ks):6.1f} us   "
      f"vs phase-A statistics flow: max|dg|/max|g| {d:.2e}  loss {r_new['loss_total']:.9f}/{r_old['loss_total']:.9f} "
      + " ".join(f"{k[:-5]}={r_new[k]:.8f}" for k in ("si_loss", "grad_loss", "smooth_loss", "reproj_loss")), flush=True)
