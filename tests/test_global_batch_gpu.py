"""Exact global-batch mode (SURVEY 8e mode B): a batch split over shards, each with its own workspace, the 32-double
statistics vectors summed between cadl_stack_reduce and cadl_stack_grad with params.global_B set -- the gradient and
the combined losses must equal the oracle on the CONCATENATED batch (reference semantics: SI n and S, reprojection n
over the whole batch, depth_loss.h:52-63, 323-330; stencil means over B_global * H * W edges, :162-165, :230-233).

  * one GPU, two shards, statistics summed with torch            (always runs)
  * one GPU, the peer-memory exchange kernel with world = 1      (always runs)
  * two GPUs, two processes, the peer-memory exchange kernel and NCCL   (skipped with fewer than two devices)
"""
import importlib
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

from conftest import PKG_NAME, ROOT, check_grad, rel_err, tie_mask

pytestmark = pytest.mark.gpu
TOL = 1e-5
WEIGHTS = (1.0, 0.1, 0.001, 0.01)


def _oracle_full(oracle, full, dev):
    t = {k: v.to(dev) for k, v in full.items()}
    p = t["pred"].clone().requires_grad_(True)
    tot, comps = oracle.combined_loss(p, t["gt"], t["rgb"], t["K"], None, *WEIGHTS)
    tot.sum().backward()
    return float(tot), {k: float(v) for k, v in comps.items()}, p.grad.detach()


def _shard_step(pkg, multi, t, lo, hi, B_global, exchange):
    """prepare -> reduce -> exchange -> grad on images [lo, hi); returns (results dict, grad, workspace)."""
    d = t["pred"].device
    sl = {k: v[lo:hi].contiguous() for k, v in t.items()}
    ws = pkg.Workspace(hi - lo, t["pred"].shape[2], t["pred"].shape[3], d)
    params = pkg.default_params(global_B=B_global)
    if pkg.stack_prepare(sl["pred"], sl["gt"], params, ws):
        params.pyramid_prepared = 1
    pkg.stack_reduce(sl["pred"], sl["gt"], None, params, ws)
    exchange(ws)
    grad = torch.empty_like(sl["pred"])
    return sl, params, ws, grad


def test_two_shards_on_one_gpu_equal_the_full_batch(pkg, oracle):
    multi = importlib.import_module(PKG_NAME + ".multi")
    d = torch.device("cuda:0")
    B, H, W = 6, 96, 160
    full = pkg.synth.make_batch(B, H, W, seed=4242)
    t = {k: v.to(d) for k, v in full.items()}
    ref_total, ref_comps, ref_grad = _oracle_full(oracle, full, d)
    off, n = pkg.lib().cadl_stats_offset(), pkg.lib().cadl_stats_count()
    shards = [(0, 4), (4, 6)]                      # unequal on purpose
    state = [_shard_step(pkg, multi, t, lo, hi, B, lambda ws: None) for lo, hi in shards]
    torch.cuda.synchronize()
    # the exchange: sum of the statistics vectors, written back into every shard's workspace
    vecs = [ws.buf[off:off + 8 * n].view(torch.float64) for _, _, ws, _ in state]
    total = torch.stack([v.clone() for v in vecs]).sum(0)
    for v in vecs:
        v.copy_(total)
    grads, res = [], []
    for (sl, params, ws, grad) in state:
        pkg.stack_grad(sl["pred"], sl["gt"], sl["rgb"], sl["K"], None, params, grad, ws)
        torch.cuda.synchronize()
        grads.append(grad)
        res.append(pkg.results_dict(ws.read_results()))
    g = torch.cat(grads, 0)
    excl = tie_mask(full["pred"], full["gt"])
    check_grad(g.cpu(), ref_grad.cpu(), excl, TOL, "global-batch gradient")
    # SI is a function of the exchanged statistics (already global on every shard); the others are additive shares
    for r in res:
        assert rel_err(r["d_si"], ref_comps["si_loss"]) <= TOL
        assert r["n_si"] == int((full["gt"] > 1e-6).sum())
    for key, name in (("d_grad", "grad_loss"), ("d_smooth", "smooth_loss"), ("d_reproj", "reproj_loss")):
        assert rel_err(sum(r[key] for r in res), ref_comps[name]) <= TOL, name
    comb = multi.combine_shares(res[0], WEIGHTS)            # single process: nothing to reduce, shares of shard 0 only
    assert comb["d_si"] == res[0]["d_si"]
    tot = WEIGHTS[0] * res[0]["d_si"] + sum(WEIGHTS[i + 1] * sum(r[k] for r in res)
                                            for i, k in enumerate(("d_grad", "d_smooth", "d_reproj")))
    assert rel_err(tot, ref_total) <= TOL


def test_exchange_kernel_single_rank_and_timeout_flag(pkg):
    """world = 1: the exchange is the identity; the timeout flag stays clear and check() does not raise."""
    multi = importlib.import_module(PKG_NAME + ".multi")
    d = torch.device("cuda:0")
    b = pkg.synth.make_batch(2, 48, 64, seed=3, device=d)
    ws = pkg.Workspace(2, 48, 64, d)
    params = pkg.default_params()
    pkg.stack_reduce(b["pred"], b["gt"], None, params, ws)
    torch.cuda.synchronize()
    off, n = pkg.lib().cadl_stats_offset(), pkg.lib().cadl_stats_count()
    before = ws.buf[off:off + 8 * n].view(torch.float64).clone()
    ex = multi.P2PStatsExchange(pkg, d, timeout_s=5.0)
    ex.exchange(ws)
    torch.cuda.synchronize()
    assert torch.equal(ws.buf[off:off + 8 * n].view(torch.float64), before)
    assert not ex.timed_out()
    ex.check()
    ex.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, B, H, W, how, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    pkg = importlib.import_module(PKG_NAME)
    multi = importlib.import_module(PKG_NAME + ".multi")
    d = torch.device(f"cuda:{rank}")
    full = pkg.synth.make_batch(B, H, W, seed=4242)
    lo, hi = multi.shard_range(B, rank, world)
    t = {k: v[lo:hi].to(d).contiguous() for k, v in full.items()}
    ws = pkg.Workspace(hi - lo, H, W, d)
    params = pkg.default_params(global_B=B)
    ex = multi.P2PStatsExchange(pkg, d, timeout_s=30.0) if how == "p2p" else None
    off, n = pkg.lib().cadl_stats_offset(), pkg.lib().cadl_stats_count()
    out = None
    for _ in range(2):                                    # twice: epochs / parities of the inboxes
        params.pyramid_prepared = 1 if pkg.stack_prepare(t["pred"], t["gt"], params, ws) else 0
        pkg.stack_reduce(t["pred"], t["gt"], None, params, ws)
        if ex is not None:
            ex.exchange(ws)
        else:
            multi.exchange_stats(ws.buf[off:off + 8 * n].view(torch.float64))
        grad = torch.empty_like(t["pred"])
        pkg.stack_grad(t["pred"], t["gt"], t["rgb"], t["K"], None, params, grad, ws)
        torch.cuda.synchronize()
        r = pkg.results_dict(ws.read_results())
        comb = multi.combine_shares(r, WEIGHTS, device=d)
        out = (lo, hi, grad.cpu(), comb)
    if ex is not None:
        ex.check()
        ex.close()
    q.put((rank,) + out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("how", ["p2p", "nccl"])
def test_two_gpus_equal_the_full_batch(pkg, oracle, how):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    B, H, W = 6, 96, 160
    full = pkg.synth.make_batch(B, H, W, seed=4242)
    ref_total, ref_comps, ref_grad = _oracle_full(oracle, full, torch.device("cuda:0"))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, B, H, W, how, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=300) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    g = torch.cat([x[3] for x in got], 0)
    check_grad(g, ref_grad.cpu(), tie_mask(full["pred"], full["gt"]), TOL, f"two GPUs ({how})")
    for x in got:                                          # after combine_shares every rank holds the global losses
        comb = x[4]
        assert rel_err(comb["d_total"], ref_total) <= TOL
        for key, name in (("d_si", "si_loss"), ("d_grad", "grad_loss"), ("d_smooth", "smooth_loss"), ("d_reproj", "reproj_loss")):
            assert rel_err(comb[key], ref_comps[name]) <= TOL, (how, name)
