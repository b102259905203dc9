"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libcadl_refharness.so, i.e.
/root/reference/src/loss/depth_loss.h + src/evaluation/depth_metrics.h on LibTorch CPU).

Run in the build container only (it needs oracle/_ref, which needs /root/reference):
    make -C oracle ref && python tests/golden/make_golden.py
The reference's own tests hold no golden vectors for this path (tests/test_models.cpp:365-509 pin rank
and sign only), so these files are the committed pin for the oracle port and, on the GPU box, a
reference-derived check that does not need /root/reference.

Each case stores the inputs' recipe (generator name, shape, seed -- regenerated bit-identically by
<pkg>/synth.py) AND the inputs themselves (small), the five losses with their ranks, dLoss/dpred for
every term, both metric variants and their integer counts.
"""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("camera-aware-neural-networks-for-few-view-depth-estimation_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "libcadl_refharness.so")
OUT = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, generator, B, H, W, seed, options
    ("rand_2x24x40", "make_batch", 2, 24, 40, 1234, {}),
    ("rand_1x64x64", "make_batch", 1, 64, 64, 7, {}),
    ("rand_3x40x72_K33", "make_batch", 3, 40, 72, 99, {"k33": True}),
    ("rand_2x33x50_odd", "make_batch", 2, 33, 50, 5, {}),           # W % 4 != 0, H not a multiple of 8
    ("smooth_2x48x64", "make_smooth_batch", 2, 48, 64, 4321, {}),   # sign(0) regions, saturation, hole band
    ("rand_2x32x48_mask", "make_batch", 2, 32, 48, 11, {"mask": True}),
    ("allinvalid_1x16x16", "make_batch", 1, 16, 16, 3, {"all_invalid": True}),
    ("equal_1x16x32", "make_batch", 1, 16, 32, 21, {"pred_eq_gt": True}),
]


def build_inputs(gen, B, H, W, seed, opt):
    b = getattr(pkg.synth, gen)(B, H, W, seed=seed)
    pred, gt, rgb, K = (b[k].numpy().copy() for k in ("pred", "gt", "rgb", "K"))
    mask = None
    if opt.get("k33"):
        K = K[0].copy()
    if opt.get("mask"):
        g = torch.Generator().manual_seed(seed + 100)
        mask = (torch.rand(B, 1, H, W, generator=g) < 0.7).numpy().astype(np.uint8)
    if opt.get("all_invalid"):
        gt[:] = 0.0
    if opt.get("pred_eq_gt"):
        gt = np.where(gt > 0, gt, 1.0).astype(np.float32)
        pred = gt.copy()
    return pred, gt, rgb, K, mask


def main():
    ref = pkg.StepHarness(REF)
    assert not ref.is_dropin(), "golden vectors must come from the reference build"
    manifest = {}
    for name, gen, B, H, W, seed, opt in CASES:
        pred, gt, rgb, K, mask = build_inputs(gen, B, H, W, seed, opt)
        rec = {"pred": pred, "gt": gt, "rgb": rgb, "K": K}
        if mask is not None:
            rec["mask"] = mask
        for term in range(6):
            cfg = pkg.StepCfg(device=-1, term=term)
            loss, rank, numel, grad = ref.loss_step(cfg, pred, gt, rgb, K, mask)
            rec[f"loss_{term}"] = np.float32(loss)
            rec[f"rank_{term}"] = np.int64(rank)
            rec[f"grad_{term}"] = grad
        # upstream-scaled backward of the trainers' call
        cfg = pkg.StepCfg(device=-1, term=0, upstream=2.5)
        _, _, _, g25 = ref.loss_step(cfg, pred, gt, rgb, K, mask)
        rec["grad_0_up2p5"] = g25
        comps = ref.components(pkg.StepCfg(device=-1), pred, gt, rgb, K, mask)
        rec["components"] = np.array([comps[k] for k in ("si_loss", "grad_loss", "smooth_loss", "reproj_loss")],
                                     dtype=np.float32)
        ev, evc = ref.metrics_eval(-1, pred, gt, mask)
        tr, trc = ref.metrics_train(-1, pred, gt)
        rec["eval"] = np.array(ev, dtype=np.float32)
        rec["eval_counts"] = np.array(evc, dtype=np.int64)
        rec["train"] = np.array(tr, dtype=np.float32)
        rec["train_counts"] = np.array(trc, dtype=np.int64)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        manifest[name] = {"generator": gen, "B": B, "H": H, "W": W, "seed": seed, "options": opt,
                          "loss_total": float(rec["loss_0"])}
        print(name, "total", float(rec["loss_0"]), "ranks", [int(rec[f"rank_{t}"]) for t in range(6)])
    manifest["_provenance"] = {"reference": "RyoK3N/Camera-Aware-Neural-Networks-for-Few-View-Depth-Estimation",
                               "build": ref.info(), "torch": torch.__version__, "device": "cpu",
                               "threads": ref.num_threads()}
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
