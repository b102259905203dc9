"""Pins the oracle port (oracle/oracle_torch.py) before anything trusts it:
  (1) against the committed golden vectors generated from the unmodified reference build
      (tests/golden/make_golden.py), and
  (2) live against oracle/_ref/libcadl_refharness.so when that build is present.
CPU only."""
import numpy as np
import pytest
import torch

from conftest import golden_cases, load_golden, rel_err

TERMS = {1: "si", 2: "grad", 3: "smooth", 4: "reproj"}


def _oracle_term(oracle, term, pred, gt, rgb, K, mask):
    if term == 0:
        return oracle.combined_loss(pred, gt, rgb, K, mask)[0]
    if term == 1:
        return oracle.scale_invariant_loss(pred, gt, mask)
    if term == 2:
        return oracle.gradient_matching_loss(pred, gt, mask)
    if term == 3:
        return oracle.smoothness_loss(pred, rgb)
    if term == 4:
        return oracle.reprojection_loss(pred, gt, K, mask)
    if term == 5:
        return oracle.combined_loss(pred, gt, rgb, None, mask)[0]
    raise ValueError(term)


def _run_oracle(oracle, term, z, upstream=1.0):
    pred = torch.from_numpy(z["pred"]).clone().requires_grad_(True)
    gt = torch.from_numpy(z["gt"])
    rgb = torch.from_numpy(z["rgb"])
    K = torch.from_numpy(z["K"])
    mask = torch.from_numpy(z["mask"]).bool() if "mask" in z else None
    loss = _oracle_term(oracle, term, pred, gt, rgb, K, mask)
    grad = torch.zeros_like(pred)
    if loss.requires_grad:
        (loss * upstream if upstream != 1.0 else loss).sum().backward()
        grad = pred.grad
    return loss, grad


@pytest.mark.parametrize("name", golden_cases())
@pytest.mark.parametrize("term", [0, 1, 2, 3, 4, 5])
def test_port_matches_golden_losses_and_grads(oracle, name, term):
    z = load_golden(name)
    loss, grad = _run_oracle(oracle, term, z)
    # same ATen kernels, same thread-independent sizes: expect bit-identical; allow 1e-6 for safety
    assert rel_err(float(loss.sum()), float(z[f"loss_{term}"])) <= 1e-6
    assert loss.dim() == int(z[f"rank_{term}"])          # ranks are part of the contract
    ref = torch.from_numpy(z[f"grad_{term}"])
    scale = float(ref.abs().max())
    assert float((grad - ref).abs().max()) <= 1e-6 * max(scale, 1e-30)


@pytest.mark.parametrize("name", golden_cases())
def test_port_matches_golden_upstream(oracle, name):
    z = load_golden(name)
    _, grad = _run_oracle(oracle, 0, z, upstream=2.5)
    ref = torch.from_numpy(z["grad_0_up2p5"])
    assert float((grad - ref).abs().max()) <= 1e-6 * max(float(ref.abs().max()), 1e-30)


@pytest.mark.parametrize("name", golden_cases())
def test_port_matches_golden_metrics(oracle, name):
    z = load_golden(name)
    pred, gt = torch.from_numpy(z["pred"]), torch.from_numpy(z["gt"])
    mask = torch.from_numpy(z["mask"]).bool() if "mask" in z else None
    ev, evc = oracle.metrics_eval(pred, gt, mask)
    tr, trc = oracle.metrics_train(pred, gt)
    for i, k in enumerate(oracle.EVAL_KEYS):
        assert rel_err(ev[k], z["eval"][i]) <= 1e-6, k
    for i, k in enumerate(oracle.TRAIN_KEYS):
        a, b = tr[k], float(z["train"][i])
        assert (np.isnan(a) and np.isnan(b)) or rel_err(a, b) <= 1e-6, k
    assert evc == [int(x) for x in z["eval_counts"]]
    assert trc == [int(x) for x in z["train_counts"]]


def test_known_answers(oracle):
    """Known-answer micro-cases: the reference has none, these follow from its formulas."""
    g = torch.Generator().manual_seed(0)
    gt = torch.empty(2, 1, 16, 24).uniform_(0.5, 8.0, generator=g)
    rgb = torch.rand(2, 3, 16, 24, generator=g)
    K = torch.tensor([[500.0, 0, 11.5], [0, 500.0, 7.5], [0, 0, 1]])
    # pred == gt: SI 0, gradient matching 0, reprojection sqrt(eps)
    assert float(oracle.scale_invariant_loss(gt, gt)) == 0.0
    assert float(oracle.gradient_matching_loss(gt, gt)) == 0.0
    assert abs(float(oracle.reprojection_loss(gt, gt, K)) - 1e-3) < 1e-7
    ev, cnt = oracle.metrics_eval(gt, gt)
    assert ev["abs_rel"] == 0.0 and ev["delta_1.25"] == 1.0 and cnt[1] == cnt[0] == gt.numel()
    # pred = c * gt: SI = (1 - lambda) * log(c)^2
    c = 1.5
    si = float(oracle.scale_invariant_loss(c * gt, gt))
    assert abs(si - 0.5 * np.log(c) ** 2) < 1e-6
    # all-invalid gt: zeros(1) for SI and reprojection (depth_loss.h:53-55, :325-327)
    z = torch.zeros_like(gt)
    assert oracle.scale_invariant_loss(gt, z).shape == (1,) and float(oracle.scale_invariant_loss(gt, z)) == 0.0
    assert oracle.reprojection_loss(gt, z, K).shape == (1,)
    # constant image: smoothness weights are exp(0) = 1
    const = torch.full_like(rgb, 0.3)
    dn = gt / (gt.mean(dim=(2, 3), keepdim=True) + 1e-6)
    expect = (dn[..., :, 1:] - dn[..., :, :-1]).abs().mean() + (dn[..., 1:, :] - dn[..., :-1, :]).abs().mean()
    assert abs(float(oracle.smoothness_loss(gt, const)) - float(expect)) < 1e-7
    # photometric stub (depth_loss.h:343-351)
    assert float(oracle.photometric_stub(gt)) == 0.0


def test_port_matches_reference_build_live(oracle, pkg, ref_harness):
    """Fresh inputs, reference .so vs port, CPU (skipped where oracle/_ref was not built/shipped)."""
    if ref_harness is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    assert not ref_harness.is_dropin()
    b = pkg.synth.make_batch(2, 40, 56, seed=2024)
    z = {k: v.numpy() for k, v in b.items()}
    for term in range(6):
        loss, rank, numel, grad = ref_harness.loss_step(pkg.StepCfg(device=-1, term=term), z["pred"], z["gt"],
                                                        z["rgb"], z["K"])
        ol, og = _run_oracle(oracle, term, z)
        assert rel_err(float(ol.sum()), loss) <= 1e-6
        assert ol.dim() == rank
        assert float((og - torch.from_numpy(grad)).abs().max()) <= 1e-6 * float(np.abs(grad).max())
    ev, evc = ref_harness.metrics_eval(-1, z["pred"], z["gt"])
    oev, oevc = oracle.metrics_eval(b["pred"], b["gt"])
    assert evc == oevc
    for i, k in enumerate(oracle.EVAL_KEYS):
        assert rel_err(oev[k], ev[i]) <= 1e-6
    tr, trc = ref_harness.metrics_train(-1, z["pred"], z["gt"])
    otr, otrc = oracle.metrics_train(b["pred"], b["gt"])
    assert trc == otrc
    for i, k in enumerate(oracle.TRAIN_KEYS):
        assert rel_err(otr[k], tr[i]) <= 1e-6


def test_closed_form_backward_matches_autograd_fp64(oracle, pkg):
    """The explicit backward formulas the CUDA kernels implement against autograd in fp64."""
    b = pkg.synth.make_batch(2, 24, 32, seed=77)
    pred = b["pred"].double().requires_grad_(True)
    gt, rgb, K = b["gt"].double(), b["rgb"].double(), b["K"].double()
    eps = 1e-6
    # SI
    oracle.scale_invariant_loss(pred, gt).backward()
    m = gt > eps
    pc, gc = pred.detach().clamp(eps, 1000), gt.clamp(eps, 1000)
    d = torch.log(pc) - torch.log(gc)
    n = m.sum()
    S = (d * m).sum()
    cm = (pred.detach() >= eps) & (pred.detach() <= 1000)
    g_si = m * cm * (2 * d / n - 2 * 0.5 * S / n ** 2) / pred.detach()
    assert torch.allclose(pred.grad, g_si, rtol=1e-10, atol=1e-16)
    pred.grad = None
    # reprojection, factored form
    oracle.reprojection_loss(pred, gt, K).backward()
    B, _, H, W = pred.shape
    u = torch.arange(W).double().view(1, 1, 1, W)
    v = torch.arange(H).double().view(1, 1, H, 1)
    fx, fy = K[:, 0, 0].view(B, 1, 1, 1), K[:, 1, 1].view(B, 1, 1, 1)
    cx, cy = K[:, 0, 2].view(B, 1, 1, 1), K[:, 1, 2].view(B, 1, 1, 1)
    r2 = ((u - cx) / (fx + eps)) ** 2 + ((v - cy) / (fy + eps)) ** 2 + 1
    dz = pred.detach() - gt
    e = torch.sqrt(dz * dz * r2 + eps)
    g_rp = m * dz * r2 / (e * n)
    assert torch.allclose(pred.grad, g_rp, rtol=1e-9, atol=1e-16)
    pred.grad = None
    # smoothness: a_b * G_j - a_b * L_b / (H*W)   (Euler homogeneity)
    oracle.smoothness_loss(pred, rgb).backward()
    p = pred.detach()
    ab = 1.0 / (p.mean(dim=(2, 3), keepdim=True) + eps)
    wx = torch.exp(-(rgb[..., :, 1:] - rgb[..., :, :-1]).abs().mean(1, keepdim=True))
    wy = torch.exp(-(rgb[..., 1:, :] - rgb[..., :-1, :]).abs().mean(1, keepdim=True))
    dx = p[..., :, 1:] - p[..., :, :-1]
    dy = p[..., 1:, :] - p[..., :-1, :]
    Nx, Ny = B * H * (W - 1), B * (H - 1) * W
    G = torch.zeros_like(p)
    G[..., :, 1:] += wx * torch.sign(dx) / Nx
    G[..., :, :-1] -= wx * torch.sign(dx) / Nx
    G[..., 1:, :] += wy * torch.sign(dy) / Ny
    G[..., :-1, :] -= wy * torch.sign(dy) / Ny
    Lb = ab * ((wx * dx.abs()).sum(dim=(1, 2, 3), keepdim=True) / Nx + (wy * dy.abs()).sum(dim=(1, 2, 3), keepdim=True) / Ny)
    g_sm = ab * G - ab * Lb / (H * W)
    assert torch.allclose(pred.grad, g_sm, rtol=1e-8, atol=1e-15)
