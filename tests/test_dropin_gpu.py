"""The drop-in C++ classes (host/loss/depth_loss.h, host/evaluation/depth_metrics.h,
host/training/validation_metrics.h) driven through the trainer-shaped harness -- the same harness
source the reference build is driven through -- and compared with (a) the unmodified reference on
LibTorch CUDA and CPU when oracle/_ref travelled to this box, (b) the committed golden vectors."""
import numpy as np
import pytest
import torch

from conftest import check_grad, golden_cases, load_golden, rel_err, tie_mask

pytestmark = pytest.mark.gpu
TOL = 1e-5
EXPECTED_RANK = {0: 1, 1: 0, 2: 1, 3: 0, 4: 0, 5: 1}     # SURVEY 8b "Outputs / ranks"


@pytest.fixture(scope="module")
def host(pkg):
    h = pkg.host_harness()
    assert h.is_dropin()
    return h


@pytest.mark.parametrize("name", golden_cases())
def test_dropin_vs_golden(pkg, host, name):
    z = load_golden(name)
    mask = z.get("mask")
    pt, gtt = torch.from_numpy(z["pred"]), torch.from_numpy(z["gt"])
    excl = {0: tie_mask(pt, gtt), 2: tie_mask(pt, gtt, smooth=False), 3: tie_mask(pt, gtt, grad=False)}
    excl[5] = excl[0]
    for term in range(6):
        loss, rank, numel, grad = host.loss_step(pkg.StepCfg(device=0, term=term), z["pred"], z["gt"], z["rgb"],
                                                 z["K"], mask)
        assert rel_err(loss, float(z[f"loss_{term}"])) <= TOL, (name, term)
        assert numel == 1
        if "allinvalid" not in name:
            assert rank == int(z[f"rank_{term}"]) == EXPECTED_RANK[term], (name, term)
        check_grad(grad, z[f"grad_{term}"], excl.get(term), TOL, f"{name}/term{term}")
    # (loss * 2.5).backward(): the autograd backward kernel (cadl_scale_grad) path
    _, _, _, g = host.loss_step(pkg.StepCfg(device=0, term=0, upstream=2.5), z["pred"], z["gt"], z["rgb"], z["K"], mask)
    check_grad(g, z["grad_0_up2p5"], excl[0], TOL, f"{name}/upstream")
    comps = host.components(pkg.StepCfg(device=0), z["pred"], z["gt"], z["rgb"], z["K"], mask)
    for i, k in enumerate(("si_loss", "grad_loss", "smooth_loss", "reproj_loss")):
        assert rel_err(comps[k], float(z["components"][i])) <= TOL, (name, k)
    ev, evc = host.metrics_eval(0, z["pred"], z["gt"], mask)
    assert evc == [int(x) for x in z["eval_counts"]]
    for i in range(12):
        assert rel_err(ev[i], float(z["eval"][i])) <= TOL, (name, i)
    tr, trc = host.metrics_train(0, z["pred"], z["gt"])
    assert trc == [int(x) for x in z["train_counts"]]
    for i in range(7):
        assert rel_err(tr[i], float(z["train"][i])) <= TOL, (name, i)


@pytest.mark.parametrize("ref_device", [0, -1])
def test_dropin_vs_reference_build(pkg, host, ref_harness, ref_device):
    """Same harness source, reference headers vs drop-in headers, same inputs (the tight oracle is the
    reference on CUDA LibTorch: same device logf)."""
    if ref_harness is None:
        pytest.skip("oracle/_ref did not travel to this box")
    b = pkg.synth.make_batch(4, 120, 160, seed=321)
    z = {k: v.numpy() for k, v in b.items()}
    excl_all = tie_mask(b["pred"], b["gt"])
    for term in range(6):
        cfg_o = pkg.StepCfg(device=0, term=term)
        cfg_r = pkg.StepCfg(device=ref_device, term=term)
        lo, ranko, _, go = host.loss_step(cfg_o, z["pred"], z["gt"], z["rgb"], z["K"])
        lr, rankr, _, gr = ref_harness.loss_step(cfg_r, z["pred"], z["gt"], z["rgb"], z["K"])
        assert rel_err(lo, lr) <= TOL, term
        assert ranko == rankr, term
        check_grad(go, gr, excl_all if term in (0, 2, 3, 5) else None, TOL, f"term{term}/ref{ref_device}")
    ev, evc = host.metrics_eval(0, z["pred"], z["gt"])
    rv, rvc = ref_harness.metrics_eval(ref_device, z["pred"], z["gt"])
    assert evc == rvc
    assert all(rel_err(a, c) <= TOL for a, c in zip(ev, rv))
    tr, trc = host.metrics_train(0, z["pred"], z["gt"])
    rt, rtc = ref_harness.metrics_train(ref_device, z["pred"], z["gt"])
    assert trc == rtc
    assert all(rel_err(a, c) <= TOL for a, c in zip(tr, rt))


def test_strict_parity_vs_cuda_reference_no_exclusions(pkg, host, ref_harness):
    """With the pooling sums in the reference's order and the same device logf, the gradient should match
    the CUDA-LibTorch reference WITHOUT excluding any pixel; report how many pixels differ if not."""
    if ref_harness is None:
        pytest.skip("oracle/_ref did not travel to this box")
    b = pkg.synth.make_batch(8, 240, 320, seed=55)
    z = {k: v.numpy() for k, v in b.items()}
    _, _, _, go = host.loss_step(pkg.StepCfg(device=0, term=2), z["pred"], z["gt"], z["rgb"], z["K"])
    _, _, _, gr = ref_harness.loss_step(pkg.StepCfg(device=0, term=2), z["pred"], z["gt"], z["rgb"], z["K"])
    bad = np.abs(go - gr) > 1e-5 * np.abs(gr).max()
    print(f"gradient-matching: {int(bad.sum())} of {bad.size} pixels differ from the CUDA reference with no exclusion")
    assert bad.mean() < 1e-5


def test_cpu_tensor_is_an_error(pkg, host):
    b = pkg.synth.make_batch(1, 16, 16, seed=2)
    z = {k: v.numpy() for k, v in b.items()}
    with pytest.raises(RuntimeError, match="CUDA"):
        host.loss_step(pkg.StepCfg(device=-1, term=0), z["pred"], z["gt"], z["rgb"], z["K"])


def test_timing_loop_runs(pkg, host):
    b = pkg.synth.make_batch(4, 240, 320, seed=3)
    z = {k: v.numpy() for k, v in b.items()}
    ms, last = host.time_steps(pkg.StepCfg(device=0, term=0), z["pred"], z["gt"], z["rgb"], z["K"], with_metrics=True,
                               include_h2d=True, warmup=2, iters=3)
    assert len(ms) == 3 and all(m > 0 for m in ms) and np.isfinite(last)
