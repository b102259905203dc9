"""Per-warp timeline of phase_b_stream_kernel at BASELINE config 3 (cadl_debug_set_trace): which SMs finish late, and why.
Usage (under gpurun):  python profiles/trace_stream.py"""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("camera-aware-neural-networks-for-few-view-depth-estimation_b200")
dev = torch.device("cuda:0")
B, H, W = 32, 480, 640
b = pkg.synth.make_batch(B, H, W, seed=1234, device=dev)
ws = pkg.Workspace(B, H, W, dev)
grad = torch.empty_like(b["pred"])
params = pkg.default_params(metrics=3)
for _ in range(5):
    pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"], None, params=params, grad=grad, ws=ws)
buf = torch.zeros(4096, 4, dtype=torch.int64, device=dev)
pkg.set_trace(buf)
for rep in range(3):
    buf.zero_()
    pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"], None, params=params, grad=grad, ws=ws)
    torch.cuda.synchronize()
    t = buf.cpu().numpy()
    fin = t[-1]
    t = t[:-1]
    t = t[t[:, 2] > 0]
    print(f"   kernel-final reduction: starts {(fin[1] - t[:, 1].min()) / 1e3:.1f} us, ends {(fin[2] - t[:, 1].min()) / 1e3:.1f} us after the first warp started")
    t0 = t[:, 1].min()
    start, end = (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3
    print(f"rep {rep}: warps {len(t)}  start us min/med/max {start.min():.1f}/{np.median(start):.1f}/{start.max():.1f}  "
          f"end us min/p10/med/p90/max {end.min():.1f}/{np.percentile(end,10):.1f}/{np.median(end):.1f}/{np.percentile(end,90):.1f}/{end.max():.1f}  "
          f"items min/med/max {t[:,3].min()}/{int(np.median(t[:,3]))}/{t[:,3].max()}")
    sm_end = {}
    for sm, e in zip(t[:, 0], end):
        sm_end[sm] = max(sm_end.get(sm, 0), e)
    se = np.array(sorted(sm_end.values()))
    print(f"   per-SM finish us: min {se.min():.1f} p25 {np.percentile(se,25):.1f} med {np.median(se):.1f} p75 {np.percentile(se,75):.1f} "
          f"p95 {np.percentile(se,95):.1f} max {se.max():.1f};  slowest SMs: {sorted(sm_end, key=sm_end.get)[-6:]}")
    dur = end - start
    print(f"   per-warp duration us: min {dur.min():.1f} med {np.median(dur):.1f} p90 {np.percentile(dur,90):.1f} max {dur.max():.1f}")
pkg.set_trace(None)
# per-launch CUDA-event times of the same step (each interval includes the launch gap before the kernel)
pkg.kernel_times(True)
acc = {}
for rep in range(20):
    pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"], None, params=params, grad=grad, ws=ws)
    for k, ms in pkg.kernel_times(True):
        acc.setdefault(k, []).append(ms)
pkg.kernel_times(False)
print("per-launch us (median of 20):", {k: round(1e3 * float(np.median(v)), 1) for k, v in acc.items()})
