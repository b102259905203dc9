"""Config-3 step WITHOUT metric variants (the trainers' training step: pool+statistics -> coefficients -> gradient pass);
used for the ncu capture profiles/r02_step_nometrics_kernels.txt.  [CADL_LIB=...] python profiles/r02_step_nometrics.py"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("camera-aware-neural-networks-for-few-view-depth-estimation_b200")
dev = torch.device("cuda:0")
B, H, W = int(os.environ.get("BATCH", "32")), 480, 640
b = pkg.synth.make_batch(B, H, W, seed=1234, device=dev)
ws = pkg.Workspace(B, H, W, dev)
grad = torch.empty_like(b["pred"])
params = pkg.default_params(metrics=0)
fn = lambda: pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"], None, params=params, grad=grad, ws=ws)
for _ in range(20):
    fn()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(100):
    fn()
e1.record(); torch.cuda.synchronize()
print(f"metrics=0: {e0.elapsed_time(e1) / 100 * 1e3:6.1f} us/step")
