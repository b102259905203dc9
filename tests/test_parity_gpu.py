"""Parity of the sm_100a kernels (through the C ABI, include/cadl.h) with the oracle.

Tolerances (BASELINE.json north_star / SURVEY 8c):
  losses   |ours - ref| / |ref| <= 1e-5 per term and total
  grads    max|g - g_ref| <= 1e-5 * max|g_ref| and ||g - g_ref||_2 <= 1e-5 * ||g_ref||_2,
           sign-tie pixels (fp64 stencil residual < 1e-6) excluded and counted
  metrics  <= 1e-5 relative on the float metrics; delta counts and n_valid IDENTICAL (integers)
"""
import numpy as np
import pytest
import torch

from conftest import check_grad, golden_cases, load_golden, rel_err, tie_mask

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _dev():
    return torch.device("cuda:0")


def _oracle_all(oracle, pred, gt, rgb, K, mask, device, dtype=torch.float32, weights=(1.0, 0.1, 0.001, 0.01)):
    """Oracle losses + grads for every term on `device`."""
    out = {}
    P = pred.to(device=device, dtype=dtype)
    G = gt.to(device=device, dtype=dtype)
    I = rgb.to(device=device, dtype=dtype)
    Kd = K.to(device=device, dtype=dtype)
    M = mask.to(device) if mask is not None else None

    def run(fn):
        p = P.clone().requires_grad_(True)
        loss = fn(p)
        g = torch.zeros_like(p)
        if loss.requires_grad:
            loss.sum().backward()
            g = p.grad
        return float(loss.sum()), g.detach()

    out["si"] = run(lambda p: oracle.scale_invariant_loss(p, G, M))
    out["grad"] = run(lambda p: oracle.gradient_matching_loss(p, G, M))
    out["smooth"] = run(lambda p: oracle.smoothness_loss(p, I))
    out["reproj"] = run(lambda p: oracle.reprojection_loss(p, G, Kd, M))
    w = weights
    out["total"] = run(lambda p: oracle.combined_loss(p, G, I, Kd, M, *w)[0])
    out["total3"] = run(lambda p: oracle.combined_loss(p, G, I, None, M, w[0], w[1], w[2])[0])
    return out


def _ours(pkg, pred, gt, rgb, K, mask, terms, **over):
    d = _dev()
    p = pkg.default_params(terms=terms, **over)
    args = [t.to(d).contiguous() if t is not None else None for t in (pred, gt, rgb, K)]
    m = mask.to(d).contiguous() if mask is not None else None
    ws = pkg.stack_fwd_bwd(args[0], args[1], args[2], args[3], m, params=p)
    torch.cuda.synchronize()
    return pkg.results_dict(ws.read_results()), ws.grad.cpu()


def _compare_case(pkg, oracle, pred, gt, rgb, K, mask, oracle_device, label):
    ref = _oracle_all(oracle, pred, gt, rgb, K, mask, oracle_device)
    excl_g = tie_mask(pred, gt, smooth=False)
    excl_s = tie_mask(pred, gt, grad=False)
    T = pkg
    singles = [("si", T.TERM_SI, "si_loss", dict(w_si=1.0), None),
               ("grad", T.TERM_GRAD, "grad_loss", dict(w_grad=1.0), excl_g),
               ("smooth", T.TERM_SMOOTH, "smooth_loss", dict(w_smooth=1.0), excl_s),
               ("reproj", T.TERM_REPROJ, "reproj_loss", dict(w_reproj=1.0), None)]
    for name, term, key, over, excl in singles:
        r, g = _ours(pkg, pred, gt, rgb, K, mask, term, **over)
        assert rel_err(r[key], ref[name][0]) <= TOL, f"{label}/{name}: {r[key]} vs {ref[name][0]}"
        check_grad(g, ref[name][1], excl, TOL, f"{label}/{name}")
    r, g = _ours(pkg, pred, gt, rgb, K, mask, T.TERM_ALL)
    assert rel_err(r["loss_total"], ref["total"][0]) <= TOL, f"{label}/total"
    for name, key in (("si", "si_loss"), ("grad", "grad_loss"), ("smooth", "smooth_loss"), ("reproj", "reproj_loss")):
        assert rel_err(r[key], ref[name][0]) <= TOL, f"{label}/total.{name}"
    check_grad(g, ref["total"][1], excl_g | excl_s, TOL, f"{label}/total")
    r, g = _ours(pkg, pred, gt, rgb, None, mask, T.TERM_SI | T.TERM_GRAD | T.TERM_SMOOTH)
    assert rel_err(r["loss_total"], ref["total3"][0]) <= TOL, f"{label}/total3"
    check_grad(g, ref["total3"][1], excl_g | excl_s, TOL, f"{label}/total3")


@pytest.mark.parametrize("name", golden_cases())
def test_golden_vectors(pkg, name):
    """Committed reference-derived vectors (no oracle code involved): losses, ranks aside, grads, metrics."""
    z = load_golden(name)
    pred, gt, rgb, K = (torch.from_numpy(z[k]) for k in ("pred", "gt", "rgb", "K"))
    mask = torch.from_numpy(z["mask"]).bool() if "mask" in z else None
    excl_g = tie_mask(pred, gt, smooth=False)
    excl_s = tie_mask(pred, gt, grad=False)
    T = pkg
    table = {0: (T.TERM_ALL, "loss_total", {}, excl_g | excl_s),
             1: (T.TERM_SI, "si_loss", dict(w_si=1.0), None),
             2: (T.TERM_GRAD, "grad_loss", dict(w_grad=1.0), excl_g),
             3: (T.TERM_SMOOTH, "smooth_loss", dict(w_smooth=1.0), excl_s),
             4: (T.TERM_REPROJ, "reproj_loss", dict(w_reproj=1.0), None),
             5: (T.TERM_SI | T.TERM_GRAD | T.TERM_SMOOTH, "loss_total", {}, excl_g | excl_s)}
    for term, (mask_bits, key, over, excl) in table.items():
        r, g = _ours(pkg, pred, gt, rgb, K if term in (0, 4) else None, mask, mask_bits, **over)
        assert rel_err(r[key], float(z[f"loss_{term}"])) <= TOL, f"{name}/term{term}: {r[key]} vs {z[f'loss_{term}']}"
        check_grad(g, z[f"grad_{term}"], excl, TOL, f"{name}/term{term}")
    r, g = _ours(pkg, pred, gt, rgb, K, mask, T.TERM_ALL, upstream=2.5)
    check_grad(g, z["grad_0_up2p5"], excl_g | excl_s, TOL, f"{name}/upstream")
    # metrics
    d = _dev()
    ws = pkg.metrics(pred.to(d), gt.to(d), mask.to(d) if mask is not None else None)
    torch.cuda.synchronize()
    r = pkg.results_dict(ws.read_results())
    assert r["eval_counts"] == [int(x) for x in z["eval_counts"]]
    for i, k in enumerate(pkg.EVAL_KEYS):
        assert rel_err(r["eval"][k], float(z["eval"][i])) <= TOL, (name, k)
    if mask is None:
        assert r["train_counts"] == [int(x) for x in z["train_counts"]]
        for i, k in enumerate(pkg.TRAIN_KEYS):
            a, b = r["train"][k], float(z["train"][i])
            assert (np.isnan(a) and np.isnan(b)) or rel_err(a, b) <= TOL, (name, k)


@pytest.mark.parametrize("shape,seed", [((2, 48, 64), 1), ((3, 96, 160), 2), ((1, 240, 320), 3), ((2, 37, 53), 4),
                                        ((2, 100, 260), 5), ((4, 16, 16), 6), ((2, 8, 8), 7)])
@pytest.mark.parametrize("oracle_device", ["cuda", "cpu"])
def test_random_vs_oracle(pkg, oracle, shape, seed, oracle_device):
    B, H, W = shape
    b = pkg.synth.make_batch(B, H, W, seed=seed)
    _compare_case(pkg, oracle, b["pred"], b["gt"], b["rgb"], b["K"], None, oracle_device, f"rand{shape}")


@pytest.mark.parametrize("oracle_device", ["cuda", "cpu"])
def test_smooth_set_sign_zero(pkg, oracle, oracle_device):
    b = pkg.synth.make_smooth_batch(2, 96, 128)
    _compare_case(pkg, oracle, b["pred"], b["gt"], b["rgb"], b["K"], None, oracle_device, "smooth")


def _plant(t, values, frac, g):
    n = t.numel()
    idx = torch.randint(0, n, (max(4, int(n * frac)),), generator=g)
    t.view(-1)[idx] = torch.tensor(values)[torch.randint(0, len(values), (idx.numel(),), generator=g)]


@pytest.mark.parametrize("shape", [(2, 48, 136), (3, 37, 53)])      # streaming path / generic kernel
def test_hostile_finite_values_vs_oracle(pkg, oracle, shape):
    """Zeros, negatives, the clamp bounds, values far outside them and exact ties pred == gt planted in pred / gt:
    losses and gradients still match the op-for-op oracle on CUDA (clamp backward on the closed interval, sign(0) = 0,
    masks computed before clamping)."""
    B, H, W = shape
    g = torch.Generator().manual_seed(H * W)
    b = pkg.synth.make_batch(B, H, W, seed=21)
    pred, gt = b["pred"].clone(), b["gt"].clone()
    _plant(pred, [0.0, -1.0, 1e-6, 1e-7, 1000.0, 1e4, 0.25, 0.1, 10.0], 0.03, g)
    _plant(gt, [0.0, -1.0, 1e-6, 1e-7, 1000.0, 1e4, 0.25, 0.1, 10.0], 0.03, g)
    tie = torch.randint(0, pred.numel(), (pred.numel() // 10,), generator=g)
    pred.view(-1)[tie] = gt.view(-1)[tie]
    _compare_case(pkg, oracle, pred, gt, b["rgb"], b["K"], None, "cuda", f"hostile{shape}")


@pytest.mark.parametrize("shape", [(2, 48, 136), (3, 37, 53)])
def test_nan_in_pred_propagates_like_the_reference(pkg, oracle, shape):
    """One NaN in pred: which losses become NaN and which gradient pixels are NaN must be what the reference's op chain
    produces (torch::clamp and the masked means let it through)."""
    B, H, W = shape
    b = pkg.synth.make_batch(B, H, W, seed=3)
    pred = b["pred"].clone()
    pred[1, 0, H // 2, W // 3] = float("nan")
    ref = _oracle_all(oracle, pred, b["gt"], b["rgb"], b["K"], None, "cuda")
    r, g = _ours(pkg, pred, b["gt"], b["rgb"], b["K"], None, pkg.TERM_ALL)
    for name, key in (("si", "si_loss"), ("grad", "grad_loss"), ("smooth", "smooth_loss"), ("reproj", "reproj_loss")):
        assert np.isnan(r[key]) == np.isnan(ref[name][0]), (name, r[key], ref[name][0])
        if not np.isnan(ref[name][0]):
            assert rel_err(r[key], ref[name][0]) <= TOL, name
    gr = ref["total"][1].cpu()
    assert torch.equal(torch.isnan(g), torch.isnan(gr)), (int(torch.isnan(g).sum()), int(torch.isnan(gr).sum()))


def test_user_mask_and_k33(pkg, oracle):
    b = pkg.synth.make_batch(2, 64, 96, seed=9)
    g = torch.Generator().manual_seed(5)
    mask = torch.rand(2, 1, 64, 96, generator=g) < 0.6
    mask &= b["gt"] > 0        # keep log(gt) finite-ish so the comparison is well conditioned
    _compare_case(pkg, oracle, b["pred"], b["gt"], b["rgb"], b["K"][0].clone(), mask, "cuda", "mask+K33")


@pytest.mark.parametrize("num_scales", [1, 2, 3, 4])
def test_num_scales(pkg, oracle, num_scales):
    b = pkg.synth.make_batch(2, 64, 128, seed=31)
    d = _dev()
    p = b["pred"].to(d).requires_grad_(True)
    loss = oracle.gradient_matching_loss(p, b["gt"].to(d), None, num_scales=num_scales)
    loss.sum().backward()
    r, g = _ours(pkg, b["pred"], b["gt"], None, None, None, pkg.TERM_GRAD, w_grad=1.0, num_scales=num_scales)
    assert rel_err(r["grad_loss"], float(loss)) <= TOL
    check_grad(g, p.grad, tie_mask(b["pred"], b["gt"], num_scales=num_scales, smooth=False), TOL, f"S={num_scales}")


def test_known_answers_on_device(pkg):
    b = pkg.synth.make_batch(2, 32, 64, seed=8)
    gt = torch.where(b["gt"] > 0, b["gt"], torch.ones_like(b["gt"]))
    r, g = _ours(pkg, gt, gt, b["rgb"], b["K"], None, pkg.TERM_ALL)
    assert r["si_loss"] == 0.0 and r["grad_loss"] == 0.0
    assert abs(r["reproj_loss"] - 1e-3) < 1e-7          # sqrt(eps)
    c = 1.5
    r, _ = _ours(pkg, c * gt, gt, None, None, None, pkg.TERM_SI, w_si=1.0)
    assert abs(r["si_loss"] - 0.5 * np.log(c) ** 2) < 1e-6
    # all-invalid gt: SI and reprojection are exactly zero with zero gradient (depth_loss.h:53-55, :325-327)
    z = torch.zeros_like(gt)
    r, g = _ours(pkg, b["pred"], z, None, b["K"], None, pkg.TERM_SI | pkg.TERM_REPROJ)
    assert r["si_loss"] == 0.0 and r["reproj_loss"] == 0.0 and float(g.abs().max()) == 0.0
    assert r["n_si"] == 0 and r["n_reproj"] == 0


def test_forward_only_and_workspace_reuse(pkg, oracle):
    """grad_pred = NULL (getComponents*), and one workspace reused across calls stays clean."""
    b = pkg.synth.make_batch(2, 64, 96, seed=12)
    d = _dev()
    t = {k: v.to(d) for k, v in b.items()}
    ws = pkg.Workspace(2, 64, 96, d)
    first = None
    for it in range(3):
        pkg.stack_fwd_bwd(t["pred"], t["gt"], t["rgb"], t["K"], None, params=pkg.default_params(),
                          want_grad=(it != 1), ws=ws)
        torch.cuda.synchronize()
        r = pkg.results_dict(ws.read_results())
        if first is None:
            first = r
        assert r["loss_total"] == first["loss_total"]      # deterministic, bit for bit
    hdr = ws.buf[:32].cpu().view(torch.int32)
    assert int(hdr[0]) == 0 and int(hdr[1]) == 0            # tickets returned to zero


def test_metrics_fused_equals_standalone_and_oracle(pkg, oracle):
    B, H, W = 4, 120, 160
    b = pkg.synth.make_batch(B, H, W, seed=44)
    d = _dev()
    t = {k: v.to(d) for k, v in b.items()}
    p = pkg.default_params(metrics=pkg.METRICS_EVAL | pkg.METRICS_TRAIN)
    ws = pkg.stack_fwd_bwd(t["pred"], t["gt"], t["rgb"], t["K"], None, params=p)
    ws2 = pkg.metrics(t["pred"], t["gt"])
    torch.cuda.synchronize()
    r, r2 = pkg.results_dict(ws.read_results()), pkg.results_dict(ws2.read_results())
    assert r["eval_counts"] == r2["eval_counts"] and r["train_counts"] == r2["train_counts"]
    ev, evc = oracle.metrics_eval(t["pred"], t["gt"])
    tr, trc = oracle.metrics_train(t["pred"], t["gt"])
    assert r["eval_counts"] == evc and r["train_counts"] == trc      # exact
    for k in pkg.EVAL_KEYS:
        assert rel_err(r["eval"][k], ev[k]) <= TOL, k
        assert rel_err(r2["eval"][k], ev[k]) <= TOL, k
    for k in pkg.TRAIN_KEYS:
        assert rel_err(r["train"][k], tr[k]) <= TOL, k


def test_metrics_with_depths_outside_the_eval_range(pkg, oracle):
    """pred and gt planted outside [0.1, 10] (and pred far inside gt's valid range): DepthMetrics::compute clamps pred
    AFTER masking (depth_metrics.h:66,154-161) while the trainers' computeDepthMetrics neither masks by range nor clamps
    (tensorboard_trainer_enhanced.h:410-436) -- the two variants share accumulators in the kernel and must still come
    out separately right (round-1 bug: rmse_log of the trainer variant counted clamped pixels twice)."""
    B, H, W = 2, 64, 96
    b = pkg.synth.make_batch(B, H, W, seed=61)
    g = torch.Generator().manual_seed(7)
    pred, gt = b["pred"].clone(), b["gt"].clone()
    _plant(pred, [0.01, 0.05, 0.0999, 10.0001, 20.0, 55.0, 0.1, 10.0], 0.10, g)
    _plant(gt, [0.05, 0.0999, 10.0001, 15.0, 30.0, 0.1, 10.0, 0.2], 0.05, g)
    d = _dev()
    for fused in (False, True):
        if fused:
            ws = pkg.stack_fwd_bwd(pred.to(d), gt.to(d), b["rgb"].to(d), b["K"].to(d), None, params=pkg.default_params(metrics=3))
        else:
            ws = pkg.metrics(pred.to(d), gt.to(d))
        torch.cuda.synchronize()
        r = pkg.results_dict(ws.read_results())
        ev, evc = oracle.metrics_eval(pred.to(d), gt.to(d))
        tr, trc = oracle.metrics_train(pred.to(d), gt.to(d))
        assert r["eval_counts"] == evc and r["train_counts"] == trc, fused
        assert trc[0] > evc[0]                     # the trainer variant keeps gt outside (0.1, 10)
        for k in pkg.EVAL_KEYS:
            assert rel_err(r["eval"][k], ev[k]) <= TOL, (fused, k, r["eval"][k], ev[k])
        for k in pkg.TRAIN_KEYS:
            assert rel_err(r["train"][k], tr[k]) <= TOL, (fused, k, r["train"][k], tr[k])


@pytest.mark.parametrize("lo,hi,scale", [(0.0, 10.0, 1.0), (100.0, 10000.0, 1000.0), (1e-3, 80.0, 8.0)])
def test_metrics_with_caller_chosen_depth_range(pkg, oracle, lo, hi, scale):
    """DepthMetrics::compute(pred, gt, mask, min, max) forwards caller-chosen bounds (depth_metrics.h:40-46): min = 0,
    millimetre depths with max > 1000 (above the clamp of the loss terms), a far range.  The metric logs are computed
    from clamp(pred, min, max) and gt themselves, not borrowed from the scale-invariant term's clamp to [1e-6, 1000]."""
    b = pkg.synth.make_batch(2, 48, 80, seed=62)
    d = _dev()
    pred, gt = (b["pred"] * scale).to(d), (b["gt"] * scale).to(d)
    ws = pkg.metrics(pred, gt, which=pkg.METRICS_EVAL, min_depth=lo, max_depth=hi)
    torch.cuda.synchronize()
    r = pkg.results_dict(ws.read_results())
    ev, evc = oracle.metrics_eval(pred, gt, None, lo, hi)
    assert r["eval_counts"] == evc and evc[0] > 0
    for k in pkg.EVAL_KEYS:
        assert rel_err(r["eval"][k], ev[k]) <= TOL, (k, r["eval"][k], ev[k])


def test_full_size_config3_properties(pkg, oracle):
    """BASELINE config 3 (32x480x640): vs the oracle on the same device, plus size-independent properties:
    linearity in upstream, determinism, count conservation, batch-permutation invariance of the loss."""
    B, H, W = 32, 480, 640
    b = pkg.synth.make_batch(B, H, W, seed=1234)
    d = _dev()
    t = {k: v.to(d) for k, v in b.items()}
    p = pkg.default_params(metrics=3)
    ws = pkg.stack_fwd_bwd(t["pred"], t["gt"], t["rgb"], t["K"], None, params=p)
    torch.cuda.synchronize()
    r = pkg.results_dict(ws.read_results())
    g1 = ws.grad.clone()
    # oracle on the GPU (seconds)
    pr = t["pred"].clone().requires_grad_(True)
    tot, comps = oracle.combined_loss(pr, t["gt"], t["rgb"], t["K"])
    tot.sum().backward()
    assert rel_err(r["loss_total"], float(tot)) <= TOL
    for k in ("si_loss", "grad_loss", "smooth_loss", "reproj_loss"):
        assert rel_err(r[k], float(comps[k])) <= TOL, k
    excl = tie_mask(t["pred"], t["gt"])
    n_excl = check_grad(g1, pr.grad, excl, TOL, "config3")
    print(f"config3: {n_excl} sign-tie pixels excluded of {g1.numel()}")
    ev, evc = oracle.metrics_eval(t["pred"], t["gt"])
    tr, trc = oracle.metrics_train(t["pred"], t["gt"])
    assert r["eval_counts"] == evc and r["train_counts"] == trc
    assert r["n_si"] == int((t["gt"] > 1e-6).sum()) == r["n_reproj"]
    # determinism + linearity in upstream (x2 is exact in binary floating point)
    # (same call otherwise: the reduction partition -- hence the last bits of the statistics -- follows the launch
    #  plan, which depends on whether metric variants are fused into the reduce pass)
    ws2 = pkg.stack_fwd_bwd(t["pred"], t["gt"], t["rgb"], t["K"], None, params=pkg.default_params(upstream=2.0, metrics=3))
    torch.cuda.synchronize()
    assert torch.equal(ws2.grad, 2.0 * g1)
    # permuting the batch permutes the gradient and leaves every loss term within rounding
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1)).to(d)
    ws3 = pkg.stack_fwd_bwd(t["pred"][perm].contiguous(), t["gt"][perm].contiguous(), t["rgb"][perm].contiguous(),
                            t["K"][perm].contiguous(), None, params=pkg.default_params())
    torch.cuda.synchronize()
    r3 = pkg.results_dict(ws3.read_results())
    assert rel_err(r3["loss_total"], r["loss_total"]) <= 1e-6
    assert float((ws3.grad - g1[perm]).abs().max()) <= 1e-6 * float(g1.abs().max())


def test_reproj_full_size_config2(pkg, oracle):
    B, H, W = 32, 480, 640
    b = pkg.synth.make_batch(B, H, W, seed=1235)
    d = _dev()
    t = {k: v.to(d) for k, v in b.items()}
    r, g = _ours(pkg, b["pred"], b["gt"], None, b["K"], None, pkg.TERM_REPROJ, w_reproj=1.0)
    pr = t["pred"].clone().requires_grad_(True)
    loss = oracle.reprojection_loss(pr, t["gt"], t["K"])
    loss.backward()
    assert rel_err(r["reproj_loss"], float(loss)) <= TOL
    check_grad(g, pr.grad, None, TOL, "config2")


def test_config5_shape_reproj_and_metrics(pkg, oracle):
    """BASELINE config 5 per-GPU shape (B=16, 960x1280: 19.66 M pixels > 2^24): reprojection fwd+bwd with both
    metric variants fused into the same pass; integer delta counts must be exact where a float mean cannot be."""
    B, H, W = 16, 960, 1280
    b = pkg.synth.make_batch(B, H, W, seed=1236)
    d = _dev()
    t = {k: v.to(d) for k, v in b.items()}
    p = pkg.default_params(terms=pkg.TERM_REPROJ, w_reproj=1.0, metrics=pkg.METRICS_EVAL | pkg.METRICS_TRAIN)
    ws = pkg.stack_fwd_bwd(t["pred"], t["gt"], None, t["K"], None, params=p)
    torch.cuda.synchronize()
    r = pkg.results_dict(ws.read_results())
    pr = t["pred"].clone().requires_grad_(True)
    loss = oracle.reprojection_loss(pr, t["gt"], t["K"])
    loss.backward()
    assert rel_err(r["reproj_loss"], float(loss)) <= TOL
    check_grad(ws.grad, pr.grad, None, TOL, "config5")
    ev, evc = oracle.metrics_eval(t["pred"], t["gt"])
    tr, trc = oracle.metrics_train(t["pred"], t["gt"])
    assert r["eval_counts"] == evc and r["train_counts"] == trc and evc[0] > 2 ** 23
    for k in pkg.EVAL_KEYS:
        if k.startswith("delta") or k == "num_valid_pixels":
            continue          # the reference's float means of 0/1 over > 2^24 elements are themselves inexact
        assert rel_err(r["eval"][k], ev[k]) <= TOL, k
    for k in ("abs_rel", "sq_rel", "rmse", "rmse_log"):
        assert rel_err(r["train"][k], tr[k]) <= TOL, k
    # fraction from the exact counts
    assert abs(r["eval"]["delta_1.25"] - evc[1] / evc[0]) < 1e-6


def test_scale_grad(pkg):
    d = _dev()
    g = torch.randn(3, 1, 40, 52, device=d)
    keep = g.clone()
    one = torch.ones(1, device=d)
    pkg.scale_grad(g, one, g)
    torch.cuda.synchronize()
    assert torch.equal(g, keep)
    up = torch.tensor([0.37], device=d)
    out = torch.empty_like(g)
    pkg.scale_grad(g, up, out)
    torch.cuda.synchronize()
    assert torch.equal(out, keep * up)


def test_errors_are_status_codes(pkg):
    d = _dev()
    b = pkg.synth.make_batch(1, 16, 16, seed=1)
    t = {k: v.to(d) for k, v in b.items()}
    with pytest.raises(pkg.CadlError):      # smoothness without an image
        pkg.stack_fwd_bwd(t["pred"], t["gt"], None, t["K"], None, params=pkg.default_params())
    with pytest.raises(pkg.CadlError):      # CPU tensor: no fallback
        pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"], None, params=pkg.default_params())
    with pytest.raises(pkg.CadlError):      # 4 scales need H, W >= 8 (torch's avg_pool2d raises too)
        small = pkg.synth.make_batch(1, 4, 4, seed=1)
        s = {k: v.to(d) for k, v in small.items()}
        pkg.stack_fwd_bwd(s["pred"], s["gt"], s["rgb"], s["K"], None, params=pkg.default_params())
