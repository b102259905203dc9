"""Measurement of the SURVEY 8(f) "next" rows on one B200 (not the headline bench; prints one JSON line per row).

  batch_prep : B=32 SUN RGB-D frames 530x730 -> 480x640 (rgb bilinear + depth nearest + K), vs the same three torch
               ops on the device (what the reference's resizeSample would run if it were moved to CUDA unchanged)
  clip       : clip_grad_norm_ over a parameter list shaped like the reference's 25.6M-parameter model
               (many small tensors + a few large), vs torch.nn.utils.clip_grad_norm_ and vs the per-tensor
               norm().item() loop of computeGradientNorm (tensorboard_trainer_enhanced.h:560-571)
"""
import importlib
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("camera-aware-neural-networks-for-few-view-depth-estimation_b200")


def timeit(fn, iters=30, warmup=5, flush=None):
    for _ in range(warmup):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        if flush is not None:
            flush.add_(1.0)
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2]


def main():
    dev = torch.device("cuda:0")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6452.5))
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev).view(torch.float32)

    B, h, w, H, W = 32, 530, 730, 480, 640
    g = torch.Generator(device=dev).manual_seed(1)
    rgb = torch.rand(B, 3, h, w, device=dev, generator=g)
    dep = torch.rand(B, 1, h, w, device=dev, generator=g) * 9 + 0.2
    K = torch.tensor([[529.5, 0, 365.0], [0, 529.5, 265.0], [0, 0, 1]], device=dev).repeat(B, 1, 1)
    ours = timeit(lambda: pkg.batch_prep(rgb, dep, K, H, W), flush=flush)

    def torch_prep():
        r = F.interpolate(rgb, size=(H, W), mode="bilinear", align_corners=False)
        d = F.interpolate(dep, size=(H, W), mode="nearest")
        K2 = K.clone()
        K2[:, 0, 0] *= W / w
        K2[:, 1, 1] *= H / h
        K2[:, 0, 2] *= W / w
        K2[:, 1, 2] *= H / h
        return r, d, K2
    ref = timeit(torch_prep, flush=flush)
    bytes_alg = 4 * B * (4 * h * w + 4 * H * W)          # read every source pixel once, write every output once
    print(json.dumps({"row": "batch_prep", "ms": ours, "torch_cuda_ms": ref, "GBps": bytes_alg / ours / 1e6,
                      "frac_of_hbm": bytes_alg / ours / 1e6 / hbm, "speedup_vs_torch_ops": ref / ours,
                      "workload": f"B={B} {h}x{w}->{H}x{W}"}))

    # a parameter list shaped like a conv encoder/decoder: 160 tensors, 25.6M elements
    sizes = []
    for c in (64, 128, 256, 512):
        for _ in range(8):
            sizes += [c * c * 9, c, c, c]
    sizes += [2048 * 1000, 1000, 3 * 64 * 49]
    scale = 25.6e6 / sum(sizes)
    sizes = [max(1, int(s * scale)) if s > 4096 else s for s in sizes]
    grads = [torch.randn(n, device=dev, generator=g) * 0.01 for n in sizes]
    params = [torch.nn.Parameter(torch.zeros(n, device=dev)) for n in sizes]
    for p, gr in zip(params, grads):
        p.grad = gr
    clipper = pkg.GradClipper(grads)
    ours = timeit(lambda: clipper(1.0), flush=flush)
    ref = timeit(lambda: torch.nn.utils.clip_grad_norm_(params, 1.0), flush=flush)

    def per_tensor_items():
        t = 0.0
        for p in params:
            n = p.grad.norm().item()
            t += n * n
        return t ** 0.5
    loop = timeit(per_tensor_items, iters=5, warmup=1, flush=flush)
    n_el = sum(sizes)
    bytes_alg = 4 * n_el * 3                               # read (norm) + read + write (scale)
    print(json.dumps({"row": "clip_grad_norm", "ms": ours, "torch_foreach_ms": ref, "reference_item_loop_ms": loop,
                      "GBps": bytes_alg / ours / 1e6, "frac_of_hbm": bytes_alg / ours / 1e6 / hbm,
                      "speedup_vs_torch": ref / ours, "speedup_vs_reference_loop": (loop + ref) / ours,
                      "workload": f"{len(sizes)} tensors, {n_el} elements"}))

    # ---- rays: unit ray maps for a batch of cameras (store-bound: 12 B/px written) ----
    Bk = 32
    Kb = pkg.synth.make_batch(Bk, 480, 640, seed=3, with_rgb=False)["K"].to(dev)
    for layout, name in ((1, "planar (B,3,H,W): the loader's layout"), (0, "(B,H*W,3): computeRayDirections' layout")):
        ms = timeit(lambda: pkg.rays_from_K(Kb, 480, 640, layout=layout), flush=flush)
        by = 12 * Bk * 480 * 640
        print(json.dumps({"row": "rays_from_K", "layout": name, "ms": ms, "GBps": by / ms / 1e6, "frac_of_hbm": by / ms / 1e6 / hbm,
                          "workload": f"B={Bk} 480x640, 12 B/px written"}))

    # ---- photometric warp (opt-in extension): 32 B/px (depth 4, target 12, source 12 through L2, gradient 4) ----
    Bp, Hp, Wp = 32, 480, 640
    bp = pkg.synth.make_batch(Bp, Hp, Wp, seed=5, device=dev)
    src = torch.rand(Bp, 3, Hp, Wp, device=dev, generator=g)
    depth = (2.0 + bp["pred"] * 0.3).contiguous()
    T = bp["T"].clone()
    T[:, 0, 3] = 0.05
    wsp = pkg.Workspace(Bp, Hp, Wp, dev)
    ms = timeit(lambda: pkg.photometric_fwd_bwd(depth, bp["K"], T, src, bp["rgb"], ws=wsp), flush=flush)
    by = 32 * Bp * Hp * Wp
    print(json.dumps({"row": "photometric_fwd_bwd", "ms": ms, "GBps": by / ms / 1e6, "frac_of_hbm": by / ms / 1e6 / hbm,
                      "Mpix_per_s": Bp * Hp * Wp / ms / 1e3, "workload": f"B={Bp} {Hp}x{Wp}, <=10 degree tilt + 5 cm baseline, 32 B/px"}))


if __name__ == "__main__":
    main()
