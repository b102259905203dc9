import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "camera-aware-neural-networks-for-few-view-depth-estimation_b200"
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libcadl_refharness.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # gpu-marked tests are skipped, not failed, where no device exists
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def oracle():
    """oracle/oracle_torch.py -- test infrastructure; only tests may import it."""
    return importlib.import_module("oracle.oracle_torch")


@pytest.fixture(scope="session")
def ref_harness(pkg):
    """The unmodified reference on LibTorch (built in the container, travels as a .so); None if absent."""
    if not os.path.exists(REF_SO):
        return None
    return pkg.StepHarness(REF_SO)


def golden_cases():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def rel_err(a, b):
    a, b = float(a), float(b)
    if a == b or (np.isnan(a) and np.isnan(b)):
        return 0.0
    return abs(a - b) / max(abs(b), 1e-30)


# --------------------------------------------------------------------------------------------
# sign-tie bookkeeping (SURVEY.md section 7 "sign discontinuities", 8c "parity definition")
# --------------------------------------------------------------------------------------------
def tie_mask(pred, gt, num_scales=4, eps=1e-6, thr=1e-6, smooth=True, grad=True):
    """Boolean (B,1,H,W) mask of pixels whose gradient depends on sign(r) for a stencil residual r with
    |r| < thr in fp64 (but r != 0 exactly in fp32 inputs is still ambiguous only through rounding): there
    a 1-ulp difference in logf / mean flips a sign term, so they are excluded from the elementwise
    gradient check and counted."""
    import torch.nn.functional as F
    p = pred.double()
    g = gt.double()
    B, _, H, W = p.shape
    out = torch.zeros(B, 1, H, W, dtype=torch.bool, device=pred.device)

    def mark(cell_mask, f):
        # cell_mask (B,1,Hs,Ws) -> pixel mask
        up = cell_mask.repeat_interleave(f, dim=2).repeat_interleave(f, dim=3)
        out[:, :, : up.shape[2], : up.shape[3]] |= up

    if grad:
        for s in range(num_scales):
            f = 2 ** s
            ps = F.avg_pool2d(p, f, stride=f) if s else p
            gs = F.avg_pool2d(g, f, stride=f) if s else g
            lp = torch.log(torch.clamp(ps, eps, 1000.0))
            lg = torch.log(torch.clamp(gs, eps, 1000.0))
            ex = (lp[..., :, 1:] - lp[..., :, :-1]) - (lg[..., :, 1:] - lg[..., :, :-1])
            ey = (lp[..., 1:, :] - lp[..., :-1, :]) - (lg[..., 1:, :] - lg[..., :-1, :])
            tx = (ex.abs() < thr) & (ex != 0)
            ty = (ey.abs() < thr) & (ey != 0)
            cm = torch.zeros_like(lp, dtype=torch.bool)
            cm[..., :, :-1] |= tx
            cm[..., :, 1:] |= tx
            cm[..., :-1, :] |= ty
            cm[..., 1:, :] |= ty
            mark(cm, f)
    if smooth:
        mean = p.mean(dim=(2, 3), keepdim=True)
        dn = p / (mean + eps)
        dx = dn[..., :, 1:] - dn[..., :, :-1]
        dy = dn[..., 1:, :] - dn[..., :-1, :]
        tx = (dx.abs() < thr) & (dx != 0)
        ty = (dy.abs() < thr) & (dy != 0)
        cm = torch.zeros_like(p, dtype=torch.bool)
        cm[..., :, :-1] |= tx
        cm[..., :, 1:] |= tx
        cm[..., :-1, :] |= ty
        cm[..., 1:, :] |= ty
        out |= cm
    return out


def check_grad(g, g_ref, excl=None, tol=1e-5, what=""):
    """max|g - g_ref| <= tol * max|g_ref|  and  ||g - g_ref||_2 <= tol * ||g_ref||_2, ties excluded."""
    g = torch.as_tensor(g).double().cpu()
    g_ref = torch.as_tensor(g_ref).double().cpu()
    assert g.shape == g_ref.shape, (g.shape, g_ref.shape)
    assert torch.isfinite(g).all() == torch.isfinite(g_ref).all()
    keep = torch.ones_like(g, dtype=torch.bool)
    n_excl = 0
    if excl is not None:
        keep = ~torch.as_tensor(excl).cpu()
        n_excl = int((~keep).sum())
    scale = float(g_ref.abs().max())
    diff = ((g - g_ref) * keep).abs()
    linf = float(diff.max())
    l2 = float(diff.norm())
    l2_ref = float((g_ref * keep).norm())
    frac_excl = n_excl / g.numel()
    assert frac_excl < 1e-3, f"{what}: too many sign-tie pixels excluded ({n_excl})"
    if scale == 0.0:
        assert linf == 0.0, f"{what}: reference gradient is zero, ours is not ({linf})"
        return n_excl
    assert linf <= tol * scale, f"{what}: max err {linf:.3e} > {tol} * max|g_ref| {scale:.3e} (excluded {n_excl})"
    assert l2 <= tol * l2_ref, f"{what}: l2 err {l2:.3e} > {tol} * ||g_ref|| {l2_ref:.3e}"
    return n_excl
