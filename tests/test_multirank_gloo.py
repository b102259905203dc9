"""world_size-2 gloo test (CPU) of the multi-GPU host logic: sharding by image, the all-reduce of the
statistics vector, and the share combination reproduce the single-process oracle on the concatenated batch."""
import importlib
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG_NAME, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, H, W, out_q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module(PKG_NAME)
    oracle = importlib.import_module("oracle.oracle_torch")
    multi = importlib.import_module(PKG_NAME + ".multi")
    full = pkg.synth.make_batch(B, H, W, seed=99)
    lo, hi = multi.shard_range(B, rank, world)
    pred, gt, rgb = full["pred"][lo:hi].double(), full["gt"][lo:hi].double(), full["rgb"][lo:hi].double()
    # per-rank statistics, as phase A produces them (fp64 here: this test is about the host logic)
    eps = 1e-6
    m = gt > eps
    d = torch.log(pred.clamp(eps, 1000)) - torch.log(gt.clamp(eps, 1000))
    stats = torch.zeros(multi.ST_COUNT, dtype=torch.float64)
    stats[multi.ST_SI_N] = m.sum()
    stats[multi.ST_SI_S] = (d * m).sum()
    stats[multi.ST_SI_Q] = (d * d * m).sum()
    stats[multi.ST_RP_N] = m.sum()
    multi.exchange_stats(stats)
    si = multi.si_from_stats(stats)
    # additive shares of the stencil terms: local sums over global denominators == loss * (local_B / B)
    gm_share = float(oracle.gradient_matching_loss(pred, gt)) * (hi - lo) / B
    sm_share = float(oracle.smoothness_loss(pred, rgb)) * (hi - lo) / B
    comb = multi.combine_shares({"d_si": si, "d_grad": gm_share, "d_smooth": sm_share, "d_reproj": 0.0},
                                weights=(1.0, 0.1, 0.001, 0.0))
    if rank == 0:
        out_q.put((si, comb["d_grad"], comb["d_smooth"], comb["d_total"], float(stats[multi.ST_SI_N])))
    dist.destroy_process_group()


def test_two_ranks_reproduce_the_global_batch(pkg, oracle):
    B, H, W = 6, 32, 48
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, H, W, q)) for r in range(2)]
    for p in procs:
        p.start()
    si, gm, sm, total, n = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = pkg.synth.make_batch(B, H, W, seed=99)
    pred, gt, rgb = full["pred"].double(), full["gt"].double(), full["rgb"].double()
    assert n == float((gt > 1e-6).sum())
    assert abs(si - float(oracle.scale_invariant_loss(pred, gt))) < 1e-12
    assert abs(gm - float(oracle.gradient_matching_loss(pred, gt))) < 1e-12
    assert abs(sm - float(oracle.smoothness_loss(pred, rgb))) < 1e-12
    ref_total, _ = oracle.combined_loss(pred, gt, rgb, None, None, 1.0, 0.1, 0.001, 0.0)
    assert abs(total - float(ref_total)) < 1e-12


def test_shard_range_covers_everything(pkg):
    multi = importlib.import_module(PKG_NAME + ".multi")
    for total in (1, 7, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [multi.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
