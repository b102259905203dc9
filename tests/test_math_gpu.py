"""Device-math replicas (csrc/cadl_math.cuh) against the CUDA library forms, bit for bit, and the aligned
fast phase-B kernel against the generic one."""
import struct

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _bits(x):
    return struct.unpack("<I", struct.pack("<f", x))[0]


def test_log_replica_is_bit_identical_to_logf_over_the_whole_clamped_range(pkg):
    """Every float in [1e-6, 1000] -- the range torch::clamp(x, eps, 1000) can produce (depth_loss.h:43-44,115-116)
    -- plus a band below/above for custom eps: 2^28 inputs, scalar and packed fp32x2 forms."""
    assert pkg.selftest(0, _bits(1e-7), _bits(1100.0)) == 0


def test_lg2_approx_error_bound_behind_the_two_tier_signs(pkg):
    """lg2.approx.ftz over every float the clamps can produce.  Measured on B200 (profiles/lg2_probe.py): the absolute
    error incl. the fp32 rounding of the result is about one ulp of the result -- below 2.3e-6 over [1e-6, 1000]
    (|log2 x| < 20), 6.8e-7 over [0.05, 20], 2.9e-7 over [1/4, 4].  The guard band kBand = 2^-15 of cadl_stream3.cuh
    (four such logs and three subtractions, plus the reference's own four logf and three subtractions: 2.2e-5 in log2
    units) and the delta-threshold bands of phase A are derived from these figures."""
    assert pkg.selftest(2, _bits(1e-7), _bits(1100.0), 2.3e-6) == 0
    assert pkg.selftest(2, _bits(0.05), _bits(20.0), 7.0e-7) == 0
    assert pkg.selftest(2, _bits(1e-7), _bits(1100.0), 2.0 ** -24) > 0          # (the check can fail)


@pytest.mark.parametrize("b", [518.8579, 519.4696, 259.43, 1037.7158, 0.3333, 3.0, 7919.0, 1e-3])
def test_markstein_division_is_correctly_rounded(pkg, b):
    """a / (fx + eps) with a = (u - cx) * depth spanning 1e-4 .. 1e5 in both signs, vs __fdiv_rn."""
    assert pkg.selftest(1, _bits(1e-4), _bits(1e5), float(np.float32(b))) == 0
    # dense sweep around typical magnitudes
    assert pkg.selftest(1, _bits(0.5), _bits(2048.0), float(np.float32(b) + np.float32(1e-6))) == 0


@pytest.mark.parametrize("shape,masked", [((3, 40, 200), True),      # 375 blocks of 8x8: the last warp of the pooled-sum pass is partly idle
                                          ((3, 40, 200), False),
                                          ((5, 24, 384), True),      # warps that straddle two images
                                          ((33, 8, 8), False),       # one block per image: every lane of a warp in another image
                                          ((2, 64, 136), False),
                                          ((1, 240, 320), True)])
@pytest.mark.parametrize("terms", ["all", "three"])
def test_loss_statistics_from_the_pooled_sum_pass(pkg, shape, masked, terms):
    """Without metric variants the step has no phase A: SI n / sum d / sum d^2, the reprojection count and the per-image
    sum(pred) come from the pooled-sum kernel (fixed-point integer atomics).  Against the layout that takes them from
    phase A (debug mode 128) and against the generic kernel (mode 1), on shapes that stress its warp-uniform loop."""
    B, H, W = shape
    d = torch.device("cuda:0")
    b = pkg.synth.make_batch(B, H, W, seed=7 * B + H + W, device=d)
    mask = (torch.rand(B, 1, H, W, device=d) < 0.7) if masked else None
    T = pkg.TERM_ALL if terms == "all" else (pkg.TERM_SI | pkg.TERM_GRAD | pkg.TERM_SMOOTH)
    res = {}
    for mode in (0, 128, 1):
        pkg.force_generic(mode)
        try:
            ws = pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"] if terms == "all" else None, mask,
                                   params=pkg.default_params(terms=T, metrics=0))
            torch.cuda.synchronize()
            res[mode] = (pkg.results_dict(ws.read_results()), ws.grad.clone())
        finally:
            pkg.force_generic(0)
    (r0, g0), (ra, ga), (rg, gg) = res[0], res[128], res[1]
    assert r0["n_si"] == ra["n_si"] == rg["n_si"] and r0["n_reproj"] == ra["n_reproj"] == rg["n_reproj"]
    for k in ("loss_total", "si_loss", "grad_loss", "smooth_loss", "reproj_loss"):
        assert rel_err(r0[k], ra[k]) <= 2e-6, (k, r0[k], ra[k])
        assert rel_err(r0[k], rg[k]) <= 2e-6, (k, r0[k], rg[k])
    scale = float(gg.abs().max())
    assert float((g0 - ga).abs().max()) <= 2e-6 * scale
    assert float((g0 - gg).abs().max()) <= 2e-6 * scale
    # twice the same call: bit-identical (integer atomics are order-free)
    ws2 = pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"] if terms == "all" else None, mask,
                            params=pkg.default_params(terms=T, metrics=0))
    torch.cuda.synchronize()
    assert torch.equal(ws2.grad, g0)
    assert rel_err(pkg.results_dict(ws2.read_results())["loss_total"], r0["loss_total"]) == 0.0      # NaN-aware (8x8)


@pytest.mark.parametrize("b_bits", [0x43FFFFFF, 0x447FFFFF, 0x3FFFFFFF, 0x4401B6E8, 0x3EAAAAAB, 0x45F78000])
def test_division_through_double_is_correctly_rounded(pkg, b_bits):
    """The gradient pass's quotient for divisors Markstein's scheme does not cover (all-ones significand: the first three
    values) is the IEEE one as well -- and for ordinary divisors too."""
    b = float(np.array([b_bits], dtype=np.uint32).view(np.float32)[0])
    assert pkg.selftest(3, _bits(1e-4), _bits(1e5), b) == 0
    assert pkg.selftest(3, _bits(1e-30), _bits(1e-25), b) == 0
    assert pkg.selftest(3, _bits(1e30), _bits(3e38), b) == 0


@pytest.mark.parametrize("shape", [(2, 48, 128), (3, 96, 160), (1, 240, 320), (2, 8, 8), (2, 56, 72)])
@pytest.mark.parametrize("terms", ["all", "three", "grad", "smooth"])
def test_fast_kernel_equals_generic_kernel(pkg, shape, terms):
    """Same inputs through both phase-B kernels: losses equal to rounding, gradients equal except where the
    fast path's SFU approximations (1/p, rsqrt, exp2) differ -- far below the 1e-5 parity tolerance."""
    B, H, W = shape
    d = torch.device("cuda:0")
    b = pkg.synth.make_batch(B, H, W, seed=sum(shape), device=d)
    T = pkg
    bits = {"all": T.TERM_ALL, "three": T.TERM_SI | T.TERM_GRAD | T.TERM_SMOOTH, "grad": T.TERM_GRAD,
            "smooth": T.TERM_SMOOTH}[terms]
    over = {"grad": dict(w_grad=1.0), "smooth": dict(w_smooth=1.0)}.get(terms, {})
    res = []
    # 0: default = the product library (pyramid + streaming kernels where they apply); through the debug library:
    # 1: generic kernel; 8: tile fast kernel (TMA staging); 10: tile fast kernel, cp.async staging
    for generic in (8, 1, 10, 0):
        pkg.force_generic(generic)
        try:
            ws = pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"] if bits & T.TERM_REPROJ else None, None,
                                   params=pkg.default_params(terms=bits, **over))
            torch.cuda.synchronize()
            res.append((pkg.results_dict(ws.read_results()), ws.grad.clone()))
        finally:
            pkg.force_generic(False)
    (rf, gf), (rg, gg), (rc, gc), (rs, gs) = res
    # streaming kernel vs tile kernel: same signs; the pointwise terms use approximate logs there (1e-6, a tenth of the parity bar)
    assert float((gs - gf).abs().max()) <= 1e-6 * float(gf.abs().max()), float((gs - gf).abs().max())
    for k in ("loss_total", "si_loss", "grad_loss", "smooth_loss", "reproj_loss"):
        assert rel_err(rs[k], rf[k]) <= 2e-6, (k, rs[k], rf[k])
    for k in ("loss_total", "si_loss", "grad_loss", "smooth_loss", "reproj_loss"):
        assert rel_err(rf[k], rg[k]) <= 2e-6, (k, rf[k], rg[k])
        assert rel_err(rf[k], rc[k]) == 0.0, (k, rf[k], rc[k])   # the two staging paths feed identical values
    scale = float(gg.abs().max())
    assert float((gf - gg).abs().max()) <= 2e-6 * scale
    assert torch.equal(gf, gc)


@pytest.mark.parametrize("shape,masked", [((2400, 8, 8), False),     # more images than resident warps: several shares per warp
                                          ((1, 16, 16), False),      # fewer strip-rows than warps
                                          ((3, 40, 200), True),      # partial last strip (200 = 128 + 72), explicit mask
                                          ((2, 64, 136), False),     # 8 columns in the last strip
                                          ((5, 24, 384), True)])
def test_streaming_path_work_partition(pkg, shape, masked):
    """The streaming fast path against the generic kernel on shapes that stress its work partition: shares that span
    strips, warps with no or several shares, partial strips, masks; with and without the dispatch-mode overrides
    (16 = no programmatic dependent launch; 32 = phase A for the statistics with the pyramid kernels in line, 128 = phase
    A with the pyramid kernels beside it: there the statistics come from another kernel, so the gradients agree to
    rounding instead of bit for bit)."""
    B, H, W = shape
    d = torch.device("cuda:0")
    b = pkg.synth.make_batch(B, H, W, seed=B + H + W, device=d)
    mask = (torch.rand(B, 1, H, W, device=d) < 0.8) if masked else None
    res = {}
    for mode in (1, 0, 16, 32, 128):
        pkg.force_generic(mode)
        try:
            ws = pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"], mask, params=pkg.default_params(metrics=3))
            torch.cuda.synchronize()
            res[mode] = (pkg.results_dict(ws.read_results()), ws.grad.clone())
        finally:
            pkg.force_generic(0)
    (rg, gg), (rs, gs), (rn, gn) = res[1], res[0], res[16]
    assert torch.equal(gs, gn)
    assert torch.equal(res[32][1], res[128][1])
    for m in (32, 128):
        assert float((gs - res[m][1]).abs().max()) <= 2e-6 * float(gg.abs().max())
        assert res[m][0]["eval_counts"] == rs["eval_counts"] and res[m][0]["train_counts"] == rs["train_counts"]
        for k in ("loss_total", "si_loss", "grad_loss", "smooth_loss", "reproj_loss"):
            assert rel_err(rs[k], res[m][0][k]) <= 2e-6, (m, k, rs[k], res[m][0][k])
    for k in ("loss_total", "si_loss", "grad_loss", "smooth_loss", "reproj_loss"):
        assert rel_err(rs[k], rg[k]) <= 2e-6, (k, rs[k], rg[k])
        assert rel_err(rs[k], rn[k]) == 0.0          # NaN-aware (8x8: the coarsest scale has no x-edges, mean of nothing)
    assert float((gs - gg).abs().max()) <= 2e-6 * float(gg.abs().max())
    assert rs["eval_counts"] == rg["eval_counts"] and rs["train_counts"] == rg["train_counts"]


def test_streaming_path_forward_only_and_repeatable(pkg):
    """grad = None (forward only: the streaming kernel's last CTA writes the results itself) gives the same losses,
    and two runs of the full step are bit-identical (fixed reduction order, no floating-point atomics)."""
    d = torch.device("cuda:0")
    b = pkg.synth.make_batch(4, 96, 256, seed=77, device=d)
    p = pkg.default_params(metrics=3)
    w1 = pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"], None, params=p)
    r1, g1 = pkg.results_dict(w1.read_results()), w1.grad.clone()
    w2 = pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"], None, params=p)
    r2, g2 = pkg.results_dict(w2.read_results()), w2.grad.clone()
    assert torch.equal(g1, g2) and r1 == r2
    w3 = pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"], None, params=p, want_grad=False)
    r3 = pkg.results_dict(w3.read_results())
    for k in ("loss_total", "si_loss", "grad_loss", "smooth_loss", "reproj_loss"):
        assert r3[k] == r1[k], (k, r3[k], r1[k])


def test_step_captures_into_a_cuda_graph(pkg):
    """The whole step (auxiliary-stream fork/join, programmatic dependent launches) is capturable: a replayed CUDA
    graph gives the same bits as the eager call."""
    d = torch.device("cuda:0")
    b = pkg.synth.make_batch(4, 96, 256, seed=5, device=d)
    p = pkg.default_params(metrics=3)
    ws = pkg.Workspace(4, 96, 256, d)
    grad = torch.empty_like(b["pred"])
    pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"], None, params=p, grad=grad, ws=ws)     # eager (creates the aux stream)
    torch.cuda.synchronize()
    g_eager, r_eager = grad.clone(), pkg.results_dict(ws.read_results())
    grad.zero_()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(graph, stream=s):
            pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"], None, params=p, grad=grad, ws=ws)
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(grad, g_eager)
    assert pkg.results_dict(ws.read_results()) == r_eager


@pytest.mark.parametrize("shape", [(4, 96, 256), (2, 37, 53)])
def test_split_api_with_prepared_pyramid(pkg, shape):
    """cadl_stack_prepare + cadl_stack_reduce + cadl_stack_grad (the global-batch flow) equals the fused call bit for
    bit; on a shape the streaming path does not take, prepare reports "unsupported" and the flow is unchanged."""
    B, H, W = shape
    d = torch.device("cuda:0")
    b = pkg.synth.make_batch(B, H, W, seed=B * H, device=d)
    p = pkg.default_params(metrics=3)
    w1 = pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"], None, params=p)
    r1, g1 = pkg.results_dict(w1.read_results()), w1.grad.clone()
    ws = pkg.Workspace(B, H, W, d)
    grad = torch.empty_like(b["pred"])
    p2 = pkg.default_params(metrics=3)
    prepared = pkg.stack_prepare(b["pred"], b["gt"], p2, ws)
    assert prepared == (H % 8 == 0 and W % 8 == 0) and p2.pyramid_prepared == int(prepared)
    pkg.stack_reduce(b["pred"], b["gt"], None, p2, ws)
    pkg.stack_grad(b["pred"], b["gt"], b["rgb"], b["K"], None, p2, grad, ws)
    torch.cuda.synchronize()
    r2 = pkg.results_dict(ws.read_results())
    assert torch.equal(grad, g1)
    for k in ("loss_total", "si_loss", "grad_loss", "smooth_loss", "reproj_loss"):
        assert rel_err(r2[k], r1[k]) == 0.0, (k, r2[k], r1[k])
    assert r2["eval_counts"] == r1["eval_counts"] and r2["train_counts"] == r1["train_counts"]


def test_p2p_stats_exchange_single_rank(pkg):
    """cadl_stats_exchange with world = 1 (own inbox only): the statistics vector comes back unchanged, for both
    epoch parities and repeatedly; no timeout is flagged.  (world > 1 is checked against NCCL by bench.py --mode global.)"""
    d = torch.device("cuda:0")
    b = pkg.synth.make_batch(2, 48, 64, seed=9, device=d)
    ws = pkg.Workspace(2, 48, 64, d)
    p = pkg.default_params(metrics=3)
    pkg.stack_reduce(b["pred"], b["gt"], None, p, ws)
    torch.cuda.synchronize()
    before = ws.stats_view().clone()
    x = pkg.multi.P2PStatsExchange(pkg, d)
    try:
        for _ in range(5):
            x.exchange(ws)
        torch.cuda.synchronize()
        assert torch.equal(ws.stats_view(), before)
        assert not x.timed_out()
    finally:
        x.close()


@pytest.mark.parametrize("shape,masked", [((4, 96, 256), False), ((3, 40, 200), True), ((2, 37, 52), False), ((700, 8, 8), False)])
def test_reprojection_alone_cooperative_count(pkg, shape, masked):
    """BASELINE config 2 path: the single cooperative launch (count + gradient) equals the two-launch form bit for
    bit (mode 64), including shapes where the cooperative form does not apply and the call falls back."""
    B, H, W = shape
    d = torch.device("cuda:0")
    b = pkg.synth.make_batch(B, H, W, seed=H + W, device=d)
    mask = (torch.rand(B, 1, H, W, device=d) < 0.7) if masked else None
    out = {}
    for mode in (0, 64):
        pkg.force_generic(mode)
        try:
            ws = pkg.stack_fwd_bwd(b["pred"], b["gt"], None, b["K"], mask,
                                   params=pkg.default_params(terms=pkg.TERM_REPROJ, w_reproj=1.0))
            torch.cuda.synchronize()
            out[mode] = (pkg.results_dict(ws.read_results()), ws.grad.clone())
        finally:
            pkg.force_generic(0)
    assert torch.equal(out[0][1], out[64][1])
    assert out[0][0]["reproj_loss"] == out[64][0]["reproj_loss"] and out[0][0]["n_reproj"] == out[64][0]["n_reproj"]
    # and a second call right after (the kernel leaves its counter clean)
    ws = pkg.stack_fwd_bwd(b["pred"], b["gt"], None, b["K"], mask, params=pkg.default_params(terms=pkg.TERM_REPROJ, w_reproj=1.0))
    torch.cuda.synchronize()
    assert torch.equal(ws.grad, out[0][1])


@pytest.mark.parametrize("seed", list(range(8)))
def test_streaming_equals_generic_on_hostile_inputs(pkg, seed):
    """Random aligned shapes with special values planted in pred/gt/rgb (0, negatives, the clamp bounds, values far
    outside them, +inf, NaN, exact ties pred == gt): the streaming path and the generic kernel must agree pixel for
    pixel -- NaN where the other has NaN, within rounding elsewhere -- and so must the losses."""
    g = torch.Generator().manual_seed(1000 + seed)
    B = int(torch.randint(1, 5, (1,), generator=g))
    H = 8 * int(torch.randint(1, 10, (1,), generator=g))
    W = 8 * int(torch.randint(1, 42, (1,), generator=g))
    d = torch.device("cuda:0")
    b = pkg.synth.make_batch(B, H, W, seed=seed, device=d)
    pred, gt, rgb = b["pred"].clone(), b["gt"].clone(), b["rgb"].clone()
    n = pred.numel()
    specials = [0.0, -1.0, 1e-6, 1e-7, 1000.0, 1e4, 0.25, 0.1, 10.0]
    if seed % 2:
        specials += [float("inf"), float("nan")]
    for t in (pred, gt):
        idx = torch.randint(0, n, (max(4, n // 50),), generator=g).to(d)
        vals = torch.tensor(specials, device=d)[torch.randint(0, len(specials), (idx.numel(),), generator=g).to(d)]
        t.view(-1)[idx] = vals
    tie = torch.randint(0, n, (max(4, n // 20),), generator=g).to(d)
    pred.view(-1)[tie] = gt.view(-1)[tie]                                  # exact ties: sign(0) = 0 paths
    rgb.view(-1)[torch.randint(0, rgb.numel(), (8,), generator=g).to(d)] = 0.0
    res = {}
    for mode in (1, 0):
        pkg.force_generic(mode)
        try:
            ws = pkg.stack_fwd_bwd(pred, gt, rgb, b["K"], None, params=pkg.default_params(metrics=3))
            torch.cuda.synchronize()
            res[mode] = (pkg.results_dict(ws.read_results()), ws.grad.clone())
        finally:
            pkg.force_generic(0)
    (rg, gg), (rs, gs) = res[1], res[0]
    assert torch.equal(torch.isnan(gg), torch.isnan(gs))
    fin = torch.isfinite(gg) & torch.isfinite(gs)
    assert torch.equal(torch.isinf(gg), torch.isinf(gs))
    if fin.any():
        scale = float(gg[fin].abs().max())
        assert float((gg[fin] - gs[fin]).abs().max()) <= 4e-6 * max(scale, 1e-30)
    for k in ("loss_total", "si_loss", "grad_loss", "smooth_loss", "reproj_loss"):
        assert rel_err(rs[k], rg[k]) <= 4e-6, (k, rs[k], rg[k])
    assert rs["eval_counts"] == rg["eval_counts"] and rs["train_counts"] == rg["train_counts"]
    assert rs["n_si"] == rg["n_si"] and rs["n_reproj"] == rg["n_reproj"]
