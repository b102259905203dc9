"""BASELINE config 4: the reference's BaselineUNet training step with the fused loss, DDP-style, on N GPUs of one node.

    python profiles/config4_unet_step.py [--H 240 --W 320 --B 32 --iters 10]                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P profiles/config4_unet_step.py ...                                      # N GPUs

One process per GPU.  The step itself is C++ (host/harness/unet_step.inc: LibTorch U-Net forward/backward, the drop-in
CombinedDepthLoss, bucketed ncclAllReduce of the 31 M-parameter gradients overlapped with backward, fused grad-clip,
Adam, device-resident loss accumulator); Python only hands the NCCL unique id around (torch.distributed, gloo) and takes
the max over ranks.  Prints one JSON line per variant on rank 0:
   reference-shaped   clip_grad_norm_ + loss.item() per step (what the reference trainers do), drop-in loss
   fused extras       FusedGradClipper + DeviceAccumulator (no host sync in the step)
   reference loss     (one GPU only, when oracle/_ref travelled) the same step with the reference's ATen loss on CUDA
"""
import argparse
import importlib
import json
import os
import statistics
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("camera-aware-neural-networks-for-few-view-depth-estimation_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=32)
ap.add_argument("--H", type=int, default=240)
ap.add_argument("--W", type=int, default=320)
ap.add_argument("--feats", type=int, default=64)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--bucket-mb", type=int, default=25)
a = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
host = pkg.host_harness()
assert host.has_unet(), "libcadl_host.so was built without the reference's model header (build() in the container does it)"
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo")
    assert host.has_nccl()
    uid = [host.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    host.nccl_init(rank, world, uid[0], local)


def max_over_ranks(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


b = pkg.synth.make_batch(a.B, a.H, a.W, seed=1234 + rank)
z = {k: v.numpy() for k, v in b.items()}
P = a.B * a.H * a.W
out = []
for name, kw in (("reference-shaped step, drop-in loss", dict(fused_extras=False)),
                 ("fused grad-clip + device loss accumulator, drop-in loss", dict(fused_extras=True))):
    if world > 1:
        dist.barrier()
    r = host.unet_train(z["rgb"], z["gt"], z["K"], device=local, feats=a.feats, world=world, rank=rank,
                        bucket_mb=a.bucket_mb, warmup=a.warmup, iters=a.iters, **kw)
    ms = max_over_ranks(statistics.median(r["ms"]))
    lms = max_over_ranks(statistics.median(r["loss_ms"]))
    out.append({"variant": name, "n_gpus": world, "B_per_gpu": a.B, "H": a.H, "W": a.W, "step_ms": ms,
                "samples_per_s": world * a.B / (ms * 1e-3), "loss_path_ms": lms, "loss_path_share": lms / ms,
                "params": r["params"], "grad_mb": r["grad_bytes"] / 1e6, "buckets": r["buckets"],
                "last_loss": r["last_loss"]})
ref_so = os.path.join(ROOT, "oracle", "_ref", "libcadl_refharness.so")
if world == 1 and os.path.exists(ref_so):
    ref = pkg.StepHarness(ref_so)
    if ref.has_unet():
        r = ref.unet_train(z["rgb"], z["gt"], z["K"], device=local, feats=a.feats, warmup=a.warmup, iters=a.iters)
        ms, lms = statistics.median(r["ms"]), statistics.median(r["loss_ms"])
        out.append({"variant": "reference-shaped step, REFERENCE loss on CUDA LibTorch (ATen op chain; its backward runs "
                               "inside loss.backward(), so loss_path_ms holds the forward only)", "n_gpus": 1,
                    "B_per_gpu": a.B, "H": a.H, "W": a.W, "step_ms": ms, "samples_per_s": a.B / (ms * 1e-3),
                    "loss_path_ms": lms, "loss_path_share": lms / ms, "params": r["params"], "last_loss": r["last_loss"]})
if rank == 0:
    for o in out:
        print(json.dumps(o), flush=True)
if world > 1:
    host.nccl_finalize()
    dist.destroy_process_group()
