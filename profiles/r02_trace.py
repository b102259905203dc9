"""Per-warp timeline of the gradient pass (tuning build with -DCADL_S3_TRACE, selected through CADL_LIB): when each warp
starts, leaves its row loop, sees its image complete and exits; how many rows found their ring slot not ready and the
time blocked on them.  Usage (under gpurun): CADL_LIB=<pkg>/csrc/variants/libcadl_trace.so python profiles/r02_trace.py"""
import ctypes as C
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
lib = C.CDLL(os.environ["CADL_LIB"])
NW = int(os.environ.get("NW", "1760"))
WPC = int(os.environ.get("WPC", "12"))
for it in range(6):
    pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"], None, params=params, grad=grad, ws=ws)
torch.cuda.synchronize()
for mode in ("in step", "alone"):
    if mode == "alone":
        pkg.stack_prepare(b["pred"], b["gt"], params, ws)
        pkg.stack_reduce(b["pred"], b["gt"], None, params, ws)
        torch.cuda.synchronize()
        pkg.stack_grad(b["pred"], b["gt"], b["rgb"], b["K"], None, params, grad, ws)
        torch.cuda.synchronize()
    out = np.zeros(8 * NW, dtype=np.uint64)
    rc = lib.cadl_debug_s3_trace(out.ctypes.data_as(C.POINTER(C.c_ulonglong)), NW)
    t = out.reshape(NW, 8).astype(np.int64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    q = lambda v: " ".join(f"{x / 1e3:6.1f}" for x in np.percentile(v, [0, 5, 25, 50, 75, 95, 100]))
    print(f"--- {mode}: {len(t)} warps, rc {rc}; percentiles 0 5 25 50 75 95 100 (us)")
    print("start        ", q(t[:, 0] - t0))
    print("row loop done", q(t[:, 1] - t0))
    print("image ready  ", q(t[:, 2] - t0))
    print("exit         ", q(t[:, 3] - t0))
    print("loop length  ", q(t[:, 1] - t[:, 0]))
    print("rows late    ", " ".join(f"{x:6.0f}" for x in np.percentile(t[:, 4], [0, 5, 25, 50, 75, 95, 100])))
    print("blocked us   ", q(t[:, 5]))
    # per SM: spread of loop end
    sm = t[:, 6]
    ends = np.array([t[sm == s, 1].max() - t0 for s in np.unique(sm)])
    print("per-SM last loop end", q(ends), " warps/SM", np.bincount(sm.astype(int)).min(), np.bincount(sm.astype(int)).max())
    # loop length by row-range index (kk) and by image
    lens = (t[:, 1] - t[:, 0]) / 1e3
    print("corr(loop length, blocked)", float(np.corrcoef(lens, t[:, 5])[0, 1]))
    # who finishes late?  rank of the warp's CTA among the CTAs of its SM (0 = lowest block index), and warp in block
    gw = np.arange(NW)[: len(t)]
    cta = gw // WPC
    rank = np.zeros(len(t), dtype=int)
    for s in np.unique(sm):
        idx = np.where(sm == s)[0]
        order = {c: r for r, c in enumerate(sorted(set(cta[idx])))}
        for i in idx:
            rank[i] = order[cta[i]]
    for r in range(rank.max() + 1):
        print(f"CTA rank {r} on its SM: mean loop length {lens[rank == r].mean():6.1f} us ({(rank == r).sum()} warps)  percentiles", q(1e3 * lens[rank == r]))
    for w in range(WPC):
        print(f"warp {w} of its CTA: mean loop length {lens[gw % WPC == w].mean():6.1f} us")
    nsm = len(np.unique(sm))
    print(f"dispatch: rank == block // SMs for {(rank == cta // nsm).mean() * 100:.1f} % of the warps ({nsm} SMs)")
    # the results block's warp: ns after the end of its row loop at which it (0) has its sums out and fenced, (1) has its
    # image ticket, (2) reaches the global ticket, (3) has it, (4) has loaded every record, (5) has reduced them,
    # (6) has written the results; (7) = length of its row loop
    fo = np.zeros(8 * 4096, dtype=np.uint64)
    lib.cadl_debug_s3_trace(fo.ctypes.data_as(C.POINTER(C.c_ulonglong)), 4096)
    print("results warp chain (ns after its last row):", [int(x) for x in fo.reshape(4096, 8)[4095].astype(np.int64)])
