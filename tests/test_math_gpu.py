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


@pytest.mark.parametrize("b", [518.8579, 519.4696, 259.43, 1037.7158, 0.3333, 3.0, 7919.0, 1e-3])
def test_markstein_division_is_correctly_rounded(pkg, b):
    """a / (fx + eps) with a = (u - cx) * depth spanning 1e-4 .. 1e5 in both signs, vs __fdiv_rn."""
    assert pkg.selftest(1, _bits(1e-4), _bits(1e5), float(np.float32(b))) == 0
    # dense sweep around typical magnitudes
    assert pkg.selftest(1, _bits(0.5), _bits(2048.0), float(np.float32(b) + np.float32(1e-6))) == 0


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
    # 0: default (streaming split where it applies); 1: generic kernel; 8: tile fast kernel (TMA staging);
    # 10: tile fast kernel, cp.async staging; 12: warp-specialised persistent tile kernel
    for generic in (8, 1, 10, 12, 0):
        pkg.force_generic(generic)
        try:
            ws = pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"] if bits & T.TERM_REPROJ else None, None,
                                   params=pkg.default_params(terms=bits, **over))
            torch.cuda.synchronize()
            res.append((pkg.results_dict(ws.read_results()), ws.grad.clone()))
        finally:
            pkg.force_generic(False)
    (rf, gf), (rg, gg), (rc, gc), (rt, gt_), (rs, gs) = res
    assert torch.equal(gf, gt_)                            # plain fast kernel == warp-specialised, bit for bit
    # streaming split vs tile kernel: the same operations per pixel, different reduction trees for the loss sums
    assert float((gs - gf).abs().max()) <= 1e-7 * float(gf.abs().max()), float((gs - gf).abs().max())
    for k in ("loss_total", "si_loss", "grad_loss", "smooth_loss", "reproj_loss"):
        assert rel_err(rs[k], rf[k]) <= 2e-6, (k, rs[k], rf[k])
    for k in ("loss_total", "si_loss", "grad_loss", "smooth_loss", "reproj_loss"):
        assert rel_err(rf[k], rg[k]) <= 2e-6, (k, rf[k], rg[k])
        assert rel_err(rf[k], rc[k]) == 0.0, (k, rf[k], rc[k])   # the two staging paths feed identical values
    scale = float(gg.abs().max())
    assert float((gf - gg).abs().max()) <= 2e-6 * scale
    assert torch.equal(gf, gc)
