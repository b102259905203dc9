"""CPU-side checks of the C ABI: the library loads, exports every symbol include/cadl.h declares, the
ctypes mirrors agree with the C structs, argument errors are status codes, and -- with no GPU -- the
product path fails loudly instead of falling back.  No compute calls here."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT


def _declared_symbols(debug=False):
    txt = open(os.path.join(ROOT, "include", "cadl.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    dbg = "".join(re.findall(r"#ifdef CADL_DEBUG(.*?)#endif", txt, flags=re.S))
    txt = re.sub(r"#ifdef CADL_DEBUG.*?#endif", "", txt, flags=re.S)
    find = lambda t: sorted(set(re.findall(r"\b(cadl_[a-zA-Z0-9_]+)\s*\(", t)))
    return (find(dbg), find(txt)) if debug else find(txt)


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.lib()
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"libcadl.so does not export {n}"
    assert set(names) == set(pkg.ABI_SYMBOLS)


def test_debug_hooks_live_in_the_debug_library_only(pkg):
    """Dispatch switches and per-launch timing are compiled out of the product library (-DCADL_DEBUG builds
    libcadl_dbg.so, which the kernel-vs-kernel tests select through pkg.force_generic)."""
    dbg_names, names = _declared_symbols(debug=True)
    assert dbg_names == ["cadl_debug_force_generic", "cadl_debug_kernel_times"]
    L, D = pkg.lib(), pkg.debug_lib()
    for n in dbg_names:
        assert not hasattr(L, n), f"product library exports {n}"
        assert hasattr(D, n)
    for n in names:
        assert hasattr(D, n)


def test_struct_mirrors(pkg):
    L = pkg.lib()
    assert L.cadl_sizeof_params() == C.sizeof(pkg.CadlParams)
    assert L.cadl_sizeof_results() == C.sizeof(pkg.CadlResults)
    p = pkg.default_params()
    # reference constructor defaults: depth_loss.h:22,84,180,257,368-371; depth_metrics.h:44-45
    assert (p.w_si, p.num_scales, p.terms) == (1.0, 4, 15)
    assert abs(p.w_grad - 0.1) < 1e-7 and abs(p.w_smooth - 0.001) < 1e-9 and abs(p.w_reproj - 0.01) < 1e-8
    assert abs(p.si_lambda - 0.5) == 0 and abs(p.eps_si - 1e-6) < 1e-12 and abs(p.min_depth - 0.1) < 1e-7
    assert p.max_depth == 10.0 and p.upstream == 1.0
    assert L.cadl_version() == 100


def test_workspace_and_errors(pkg):
    L = pkg.lib()
    assert L.cadl_workspace_bytes(0, 4, 4) == 0
    a, b = L.cadl_workspace_bytes(1, 64, 64), L.cadl_workspace_bytes(32, 480, 640)
    assert 0 < a < b < 64 << 20
    assert L.cadl_stats_count() == 32 and L.cadl_stats_offset() % 256 == 0
    p = pkg.default_params()
    # NULL workspace / params are argument errors, detected before any CUDA call
    rc = L.cadl_stack_fwd_bwd(None, None, None, None, None, 1, 8, 8, C.byref(p), None, None, None, 0, None)
    assert rc == 1 and b"NULL" in L.cadl_error_string(rc)
    rc = L.cadl_stack_fwd_bwd(None, None, None, None, None, 0, 8, 8, C.byref(p), None, None, None, 0, None)
    assert rc == 2
    assert b"CUDA error" in L.cadl_error_string(1000 + 100)


def test_no_cpu_fallback(pkg):
    b = pkg.synth.make_batch(1, 16, 16)
    with pytest.raises(pkg.CadlError, match="CUDA"):
        pkg.stack_fwd_bwd(b["pred"], b["gt"], b["rgb"], b["K"])
    with pytest.raises(pkg.CadlError):
        pkg.metrics(b["pred"], b["gt"])


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a device")
def test_without_a_device_calls_return_cuda_errors(pkg):
    L = pkg.lib()
    buf = (C.c_char * 4096)()
    rc = L.cadl_workspace_init(C.addressof(buf), 4096, None)
    assert rc >= 1000      # 1000 + cudaError (no device / insufficient driver); never a silent success


def test_host_library_links_the_dropin_headers(pkg):
    """libcadl_host.so is the step harness compiled against host/loss/depth_loss.h etc.; the very same source
    compiles against the reference headers (oracle/_ref).  Loading needs no GPU."""
    h = pkg.StepHarness(pkg.LIBHOST_PATH)
    assert h.is_dropin() and "drop-in" in h.info()
    for sym in ("cadh_loss_step", "cadh_components", "cadh_metrics_eval", "cadh_metrics_train", "cadh_time_steps"):
        assert hasattr(h.L, sym)
    if not torch.cuda.is_available():
        b = pkg.synth.make_batch(1, 16, 16)
        z = {k: v.numpy() for k, v in b.items()}
        with pytest.raises(RuntimeError):     # CPU tensors are refused by the drop-in: no fallback
            h.loss_step(pkg.StepCfg(device=-1, term=0), z["pred"], z["gt"], z["rgb"], z["K"])
