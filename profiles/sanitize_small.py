"""Small shapes through every kernel family, for compute-sanitizer (one tool per gpurun call)."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("camera-aware-neural-networks-for-few-view-depth-estimation_b200")
dev = torch.device("cuda:0")
for (B, H, W) in [(2, 48, 128), (2, 56, 72), (1, 37, 53), (2, 8, 8), (3, 96, 160)]:
    b = pkg.synth.make_batch(B, H, W, seed=B * H, device=dev)
    mask = (torch.rand(B, 1, H, W, device=dev) < 0.7)
    for mode in (0, 1, 8, 10, 12):
        pkg.force_generic(mode)
        for terms, K in ((pkg.TERM_ALL, b["K"]), (pkg.TERM_SI | pkg.TERM_GRAD | pkg.TERM_SMOOTH, None),
                         (pkg.TERM_GRAD, None), (pkg.TERM_SMOOTH, None), (pkg.TERM_SI, None), (pkg.TERM_REPROJ, b["K"]),
                         (pkg.TERM_SI | pkg.TERM_REPROJ, b["K"])):
            for m in (None, mask):
                over = {}
                if terms == pkg.TERM_GRAD: over["w_grad"] = 1.0
                p = pkg.default_params(terms=terms, metrics=3, **over)
                ws = pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"] if terms & pkg.TERM_SMOOTH else None, K, m, params=p)
    pkg.force_generic(0)
    pkg.metrics(b["pred"], b["gt"], mask)
    pkg.rays_from_K(b["K"], H, W, layout=0)
    pkg.rays_from_K(b["K"], H, W, layout=1, pose=b["T"])
    src = torch.rand(B, 3, H, W, device=dev)
    pkg.photometric_fwd_bwd(b["pred"], b["K"], b["T"], src, b["rgb"])
    g = torch.randn(B, 1, H, W, device=dev)
    pkg.scale_grad(g, torch.tensor([0.5], device=dev), g)
torch.cuda.synchronize()
print("sanitize_small: done")
