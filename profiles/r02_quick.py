"""Round-2 tuning aid: config-3 step time, per-launch event times and old-vs-new agreement, in one process.
Usage (under gpurun):  [CADL_LIB=...] python profiles/r02_quick.py [gsz ...]"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("camera-aware-neural-networks-for-few-view-depth-estimation_b200")
dev = torch.device("cuda:0")
B, H, W = 32, 480, 640
b = pkg.synth.make_batch(B, H, W, seed=1234, device=dev)
ws = pkg.Workspace(B, H, W, dev)
grad = torch.empty_like(b["pred"])


def step(params):
    pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"], None, params=params, grad=grad, ws=ws)


def timeit(params, n=200):
    for _ in range(20):
        step(params)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        step(params)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


tag = os.path.basename(os.environ.get("CADL_LIB", "libcadl.so"))
for metrics in (3, 0):
    params = pkg.default_params(metrics=metrics)
    # agreement of the two gradient-pass kernels
    pkg.force_generic(128)
    step(params)
    torch.cuda.synchronize()
    g_old, r_old = grad.clone(), pkg.results_dict(ws.read_results())
    pkg.force_generic(0)
    step(params)
    torch.cuda.synchronize()
    r_new = pkg.results_dict(ws.read_results())
    d = (grad - g_old).abs().max().item() / g_old.abs().max().item()
    nd = int((grad != g_old).sum())
    print(f"[{tag}] metrics={metrics}: new vs old gradient: max|diff|/max|g| = {d:.3e}, {nd} differing px; "
          f"loss {r_new['loss_total']:.9f} vs {r_old['loss_total']:.9f}; "
          + " ".join(f"{k}={r_new[k]:.8f}/{r_old[k]:.8f}" for k in ("si_loss", "grad_loss", "smooth_loss", "reproj_loss")))
    for gsz in [8]:
        t_new = timeit(params)
        pkg.force_generic(128)
        t_old = timeit(params)
        pkg.force_generic(0)
        print(f"[{tag}] metrics={metrics} gsz={gsz}: new {t_new:7.1f} us/step   old {t_old:7.1f} us/step")
    pkg.kernel_times(True)
    for _ in range(3):
        step(params)
    kt = pkg.kernel_times(False)
    print(f"[{tag}] in-line launches: " + "  ".join(f"{k} {v * 1e3:.1f}" for k, v in kt))
