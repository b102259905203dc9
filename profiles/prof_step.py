"""Minimal driver for ncu: a few steps of the fused path at BASELINE config 3 (or config 2 with --reproj).
Usage (under gpurun):  python profiles/prof_step.py [--steps N] [--reproj] [--B 32 --H 480 --W 640]"""
import argparse
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("camera-aware-neural-networks-for-few-view-depth-estimation_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--reproj", action="store_true")
ap.add_argument("--B", type=int, default=32)
ap.add_argument("--H", type=int, default=480)
ap.add_argument("--W", type=int, default=640)
a = ap.parse_args()
dev = torch.device("cuda:0")
b = pkg.synth.make_batch(a.B, a.H, a.W, seed=1234, device=dev)
ws = pkg.Workspace(a.B, a.H, a.W, dev)
grad = torch.empty_like(b["pred"])
if a.reproj:
    params = pkg.default_params(terms=pkg.TERM_REPROJ, w_reproj=1.0)
else:
    params = pkg.default_params(metrics=pkg.METRICS_EVAL | pkg.METRICS_TRAIN)
for _ in range(a.steps):
    pkg.stack_fwd_bwd(b["pred"], b["gt"], None if a.reproj else b["rgb"], b["K"], None, params=params, grad=grad, ws=ws)
torch.cuda.synchronize()
r = pkg.results_dict(ws.read_results())
print("loss_total", r["loss_total"], "reproj", r["reproj_loss"])
