"""B200-native camera-aware depth-loss kernels: Python access to the C ABI.

The product is ``csrc/libcadl.so`` (hand-written sm_100a kernels behind ``include/cadl.h``) and the
C++ drop-in headers under ``host/``.  This package is only the thin ctypes layer that tests, the
benchmark and ``__graft_entry__`` use to reach them; PyTorch supplies device memory and streams.

There is no CPU implementation: every entry point raises if the library is missing or a tensor is
not on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
LIBCADL_PATH = os.environ.get("CADL_LIB", os.path.join(_HERE, "csrc", "libcadl.so"))   # CADL_LIB: tuning builds (profiles/)
LIBCADL_DBG_PATH = os.path.join(_HERE, "csrc", "libcadl_dbg.so")   # same sources with -DCADL_DEBUG: dispatch switches, per-launch timing
LIBHOST_PATH = os.path.join(_HERE, "host", "libcadl_host.so")

TERM_SI, TERM_GRAD, TERM_SMOOTH, TERM_REPROJ, TERM_ALL = 1, 2, 4, 8, 15
METRICS_EVAL, METRICS_TRAIN = 1, 2

EVAL_KEYS = ("abs_rel", "sq_rel", "rmse", "rmse_log", "mae", "log10",
             "delta_1.25", "delta_1.25^2", "delta_1.25^3",
             "num_valid_pixels", "mean_pred_depth", "mean_gt_depth")
TRAIN_KEYS = ("abs_rel", "sq_rel", "rmse", "rmse_log", "a1", "a2", "a3")


class CadlParams(C.Structure):
    """Mirror of ``cadl_params`` (include/cadl.h)."""
    _fields_ = [
        ("terms", C.c_uint32), ("metrics", C.c_uint32),
        ("w_si", C.c_float), ("w_grad", C.c_float), ("w_smooth", C.c_float), ("w_reproj", C.c_float),
        ("si_lambda", C.c_float),
        ("eps_si", C.c_float), ("eps_grad", C.c_float), ("eps_smooth", C.c_float), ("eps_reproj", C.c_float),
        ("num_scales", C.c_int32), ("k_batched", C.c_int32),
        ("min_depth", C.c_float), ("max_depth", C.c_float),
        ("upstream", C.c_float),
        ("global_B", C.c_int32),
        ("pyramid_prepared", C.c_int32),
    ]


class CadlResults(C.Structure):
    """Mirror of ``cadl_results`` (include/cadl.h)."""
    _fields_ = [
        ("loss_total", C.c_float), ("loss_si", C.c_float), ("loss_grad", C.c_float),
        ("loss_smooth", C.c_float), ("loss_reproj", C.c_float), ("_pad0", C.c_float * 3),
        ("d_total", C.c_double), ("d_si", C.c_double), ("d_grad", C.c_double),
        ("d_smooth", C.c_double), ("d_reproj", C.c_double),
        ("n_si", C.c_int64), ("n_reproj", C.c_int64),
        ("eval", C.c_float * 12), ("eval_counts", C.c_int64 * 4),
        ("train", C.c_float * 8), ("train_counts", C.c_int64 * 4),
    ]


# every symbol include/cadl.h declares (tests/test_abi.py checks the .so exports each one)
ABI_SYMBOLS = (
    "cadl_default_params", "cadl_version", "cadl_sizeof_params", "cadl_sizeof_results", "cadl_error_string",
    "cadl_selftest", "cadl_workspace_bytes", "cadl_workspace_init",
    "cadl_stack_fwd_bwd", "cadl_stack_reduce", "cadl_stack_grad", "cadl_stats_offset", "cadl_stats_count",
    "cadl_si_fwd_bwd", "cadl_gradmatch_fwd_bwd", "cadl_smooth_fwd_bwd", "cadl_reproj_fwd_bwd",
    "cadl_scale_grad", "cadl_metrics", "cadl_rays_from_K", "cadl_photometric_fwd_bwd",
    "cadl_batch_prep", "cadl_clip_workspace_bytes", "cadl_clip_grad_norm",
    "cadl_batch_augment", "cadl_accumulate", "cadl_stack_prepare",
    "cadl_p2p_inbox_bytes", "cadl_p2p_alloc", "cadl_p2p_open", "cadl_p2p_close", "cadl_stats_exchange", "cadl_p2p_error",
)

_lib = None        # the product library
_dbg = None        # the debug build (tests / profiling only)
_active = None     # what the wrappers below call: the product library unless force_generic() selected a debug dispatch


def lib() -> C.CDLL:
    """The library the wrappers call: libcadl.so (built by ``__graft_entry__.build()``), or -- while a debug dispatch
    mode is selected with ``force_generic`` / ``kernel_times`` -- libcadl_dbg.so.  Fails loudly if it is not there."""
    global _lib
    if _active is not None:
        return _active
    if _lib is None:
        _lib = _load(LIBCADL_PATH)
    return _lib


def debug_lib() -> C.CDLL:
    global _dbg
    if _dbg is None:
        _dbg = _load(LIBCADL_DBG_PATH, debug=True)
    return _dbg


def _load(path: str, debug: bool = False) -> C.CDLL:
    if not os.path.exists(path):
        raise RuntimeError(
            f"cadl: {path} is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'`"
            " (nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(path)
    vp, f32p, u8p = C.c_void_p, C.c_void_p, C.c_void_p
    L.cadl_default_params.argtypes = [C.POINTER(CadlParams)]
    L.cadl_default_params.restype = None
    L.cadl_version.restype = C.c_int
    L.cadl_sizeof_params.restype = C.c_size_t
    L.cadl_sizeof_results.restype = C.c_size_t
    L.cadl_error_string.argtypes = [C.c_int]
    L.cadl_error_string.restype = C.c_char_p
    L.cadl_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    L.cadl_workspace_bytes.restype = C.c_size_t
    L.cadl_workspace_init.argtypes = [vp, C.c_size_t, vp]
    L.cadl_stats_offset.restype = C.c_size_t
    L.cadl_stats_count.restype = C.c_int
    L.cadl_stack_fwd_bwd.argtypes = [f32p, f32p, f32p, f32p, u8p, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(CadlParams), f32p, vp, vp, C.c_size_t, vp]
    L.cadl_stack_reduce.argtypes = [f32p, f32p, u8p, C.c_int, C.c_int, C.c_int, C.POINTER(CadlParams),
                                    vp, C.c_size_t, vp]
    L.cadl_stack_grad.argtypes = L.cadl_stack_fwd_bwd.argtypes
    L.cadl_si_fwd_bwd.argtypes = [f32p, f32p, u8p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                  f32p, vp, vp, C.c_size_t, vp]
    L.cadl_gradmatch_fwd_bwd.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                         f32p, vp, vp, C.c_size_t, vp]
    L.cadl_smooth_fwd_bwd.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                      f32p, vp, vp, C.c_size_t, vp]
    L.cadl_reproj_fwd_bwd.argtypes = [f32p, f32p, f32p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_float,
                                      C.c_float, f32p, vp, vp, C.c_size_t, vp]
    L.cadl_scale_grad.argtypes = [f32p, f32p, f32p, C.c_size_t, vp]
    L.cadl_metrics.argtypes = [f32p, f32p, u8p, C.c_size_t, C.c_uint32, C.c_float, C.c_float, vp, vp,
                               C.c_size_t, vp]
    L.cadl_rays_from_K.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_int, C.c_int, C.c_int, f32p, vp]
    L.cadl_photometric_fwd_bwd.argtypes = [f32p, f32p, C.c_int, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int,
                                           C.c_float, C.c_float, f32p, vp, vp, C.c_size_t, vp]
    L.cadl_batch_prep.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p, f32p, f32p, vp]
    L.cadl_clip_workspace_bytes.restype = C.c_size_t
    L.cadl_clip_grad_norm.argtypes = [vp, vp, vp, C.c_int, C.c_longlong, C.c_float, f32p, vp, C.c_size_t, C.c_int, vp]
    if debug:
        L.cadl_debug_force_generic.argtypes = [C.c_int]
        L.cadl_debug_force_generic.restype = None
        L.cadl_debug_kernel_times.argtypes = [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_char_p), C.c_int]
        L.cadl_debug_kernel_times.restype = C.c_int
    L.cadl_selftest.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_float, vp, vp]
    for name in ABI_SYMBOLS:
        getattr(L, name)   # AttributeError here = the .so does not export what include/cadl.h declares
    if L.cadl_sizeof_params() != C.sizeof(CadlParams) or L.cadl_sizeof_results() != C.sizeof(CadlResults):
        raise RuntimeError("cadl: ctypes mirrors of cadl_params/cadl_results disagree with libcadl.so")
    return L


class CadlError(RuntimeError):
    pass


def _check(rc: int, what: str):
    if rc != 0:
        raise CadlError(f"{what}: {lib().cadl_error_string(rc).decode()} (code {rc})")


def default_params(**over) -> CadlParams:
    p = CadlParams()
    lib().cadl_default_params(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise CadlError("cadl: tensors must live on a CUDA device (no CPU fallback)")


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


class Workspace:
    """Caller-owned scratch for one (B,H,W) problem on one stream; zeroed once, left clean by every call."""

    def __init__(self, B: int, H: int, W: int, device):
        self.shape = (B, H, W)
        self.bytes = int(lib().cadl_workspace_bytes(B, H, W))
        if self.bytes == 0:
            raise CadlError("cadl: bad problem size")
        self.buf = torch.zeros(self.bytes, dtype=torch.uint8, device=device)
        self.results = torch.zeros(C.sizeof(CadlResults), dtype=torch.uint8, device=device)

    def stats_view(self) -> torch.Tensor:
        """The exchangeable fp64 statistics vector (all-reduce it between reduce and grad in mode B)."""
        off = int(lib().cadl_stats_offset())
        n = int(lib().cadl_stats_count())
        return self.buf[off:off + 8 * n].view(torch.float64)

    def read_results(self) -> CadlResults:
        host = self.results.cpu().numpy().tobytes()
        return CadlResults.from_buffer_copy(host)


def results_dict(r: CadlResults) -> Dict[str, object]:
    return {
        "loss_total": r.loss_total, "si_loss": r.loss_si, "grad_loss": r.loss_grad,
        "smooth_loss": r.loss_smooth, "reproj_loss": r.loss_reproj,
        "d_total": r.d_total, "d_si": r.d_si, "d_grad": r.d_grad, "d_smooth": r.d_smooth, "d_reproj": r.d_reproj,
        "n_si": r.n_si, "n_reproj": r.n_reproj,
        "eval": {k: r.eval[i] for i, k in enumerate(EVAL_KEYS)},
        "eval_counts": [int(x) for x in r.eval_counts],
        "train": {k: r.train[i] for i, k in enumerate(TRAIN_KEYS)},
        "train_counts": [int(x) for x in r.train_counts],
    }


def stack_fwd_bwd(pred, gt, rgb, K, mask=None, params: Optional[CadlParams] = None,
                  grad: Optional[torch.Tensor] = None, want_grad: bool = True,
                  ws: Optional[Workspace] = None) -> Workspace:
    """Launch the fused forward+backward (asynchronous).  Results stay on the device in ``ws.results``;
    the gradient is written to ``grad`` (allocated as ``ws.grad`` if not given)."""
    _require_cuda(pred, gt, rgb, K, mask, grad)
    B, _, H, W = pred.shape
    p = params if params is not None else default_params()
    if K is not None:
        p.k_batched = 1 if K.dim() == 3 else 0
    if ws is None:
        ws = Workspace(B, H, W, pred.device)
    if want_grad and grad is None:
        grad = torch.empty_like(pred)
    ws.grad = grad if want_grad else None
    with torch.cuda.device(pred.device):
        rc = lib().cadl_stack_fwd_bwd(_ptr(pred), _ptr(gt), _ptr(rgb), _ptr(K), _ptr(mask), B, H, W, C.byref(p),
                                      _ptr(ws.grad), _ptr(ws.results), _ptr(ws.buf), ws.bytes, _stream(pred))
    _check(rc, "cadl_stack_fwd_bwd")
    return ws


def stack_prepare(pred, gt, params: CadlParams, ws: Workspace) -> bool:
    """Split API: start the pooled-pyramid kernels beside stack_reduce / the all-reduce.  Sets
    ``params.pyramid_prepared`` and returns True when the streaming path applies, else leaves it 0."""
    _require_cuda(pred, gt)
    B, _, H, W = pred.shape
    params.pyramid_prepared = 0
    with torch.cuda.device(pred.device):
        rc = lib().cadl_stack_prepare(_ptr(pred), _ptr(gt), B, H, W, C.byref(params), _ptr(ws.buf), ws.bytes, _stream(pred))
    if rc == 4:      # CADL_ERR_UNSUPPORTED: the generic / tile path will run, nothing to prepare
        return False
    _check(rc, "cadl_stack_prepare")
    params.pyramid_prepared = 1
    return True


def stack_reduce(pred, gt, mask, params: CadlParams, ws: Workspace):
    _require_cuda(pred, gt, mask)
    B, _, H, W = pred.shape
    with torch.cuda.device(pred.device):
        rc = lib().cadl_stack_reduce(_ptr(pred), _ptr(gt), _ptr(mask), B, H, W, C.byref(params), _ptr(ws.buf),
                                     ws.bytes, _stream(pred))
    _check(rc, "cadl_stack_reduce")


def stack_grad(pred, gt, rgb, K, mask, params: CadlParams, grad, ws: Workspace):
    _require_cuda(pred, gt, rgb, K, mask, grad)
    B, _, H, W = pred.shape
    ws.grad = grad
    with torch.cuda.device(pred.device):
        rc = lib().cadl_stack_grad(_ptr(pred), _ptr(gt), _ptr(rgb), _ptr(K), _ptr(mask), B, H, W, C.byref(params),
                                   _ptr(grad), _ptr(ws.results), _ptr(ws.buf), ws.bytes, _stream(pred))
    _check(rc, "cadl_stack_grad")


def metrics(pred, gt, mask=None, which: int = METRICS_EVAL | METRICS_TRAIN, min_depth: float = 0.1,
            max_depth: float = 10.0, ws: Optional[Workspace] = None) -> Workspace:
    _require_cuda(pred, gt, mask)
    n = pred.numel()
    if ws is None:
        ws = Workspace(1, 1, n, pred.device)
    with torch.cuda.device(pred.device):
        rc = lib().cadl_metrics(_ptr(pred), _ptr(gt), _ptr(mask), n, which, min_depth, max_depth, _ptr(ws.results),
                                _ptr(ws.buf), ws.bytes, _stream(pred))
    _check(rc, "cadl_metrics")
    return ws


def force_generic(on):
    """Test hook (bit mask, see include/cadl.h under CADL_DEBUG): a non-zero mode routes every following call through
    libcadl_dbg.so with that dispatch (1 = generic phase-B kernel, 8 = one tile kernel instead of the pyramid +
    streaming kernels, 8|2 = that kernel staged with cp.async, 16 = no programmatic dependent launch, 32 = pyramid
    kernels in line, 64 = reprojection alone with the separate count kernel, 128 = loss statistics from phase A instead
    of the pooled-sum pass); 0 returns to the product library."""
    global _active
    on = int(on)
    if on:
        _active = debug_lib()
        _active.cadl_debug_force_generic(on)
    else:
        if _dbg is not None:
            _dbg.cadl_debug_force_generic(0)
        _active = None


def kernel_times(enable: bool = True):
    """Per-launch event timing of stack_fwd_bwd (debug library: selected while enabled); returns [(kernel, ms)] of the
    last timed call."""
    global _active
    L = debug_lib()
    ms = (C.c_float * 12)()
    names = (C.c_char_p * 12)()
    n = L.cadl_debug_kernel_times(int(enable), ms, names, 12)
    _active = L if enable else None
    return [(names[i].decode(), float(ms[i])) for i in range(n)]


def selftest(which: int, lo_bits: int, hi_bits: int, param: float = 0.0, device="cuda:0") -> int:
    """Number of inputs in [lo_bits, hi_bits] where a device-math replica differs from the CUDA library form."""
    out = torch.zeros(1, dtype=torch.int64, device=device)
    with torch.cuda.device(out.device):
        rc = lib().cadl_selftest(which, lo_bits, hi_bits, param, _ptr(out),
                                 C.c_void_p(torch.cuda.current_stream(out.device).cuda_stream))
    _check(rc, "cadl_selftest")
    return int(out.item())


def scale_grad(grad_in, upstream_dev, grad_out):
    _require_cuda(grad_in, upstream_dev, grad_out)
    with torch.cuda.device(grad_in.device):
        rc = lib().cadl_scale_grad(_ptr(grad_in), _ptr(upstream_dev), _ptr(grad_out), grad_in.numel(),
                                   _stream(grad_in))
    _check(rc, "cadl_scale_grad")


def rays_from_K(K, H: int, W: int, layout: int = 1, pose=None) -> torch.Tensor:
    """Unit ray directions from intrinsics: layout 0 -> (B,H*W,3), layout 1 -> (B,3,H,W)."""
    _require_cuda(K, pose)
    kb = 1 if K.dim() == 3 else 0
    B = K.shape[0] if kb else (pose.shape[0] if pose is not None else 1)
    out = torch.empty((B, H * W, 3) if layout == 0 else (B, 3, H, W), dtype=torch.float32, device=K.device)
    with torch.cuda.device(K.device):
        rc = lib().cadl_rays_from_K(_ptr(K), kb, _ptr(pose), B, H, W, layout, _ptr(out), _stream(K))
    _check(rc, "cadl_rays_from_K")
    return out


def photometric_fwd_bwd(pred, K, T, source, target, eps: float = 1e-6, upstream: float = 1.0,
                        want_grad: bool = True, ws: Optional[Workspace] = None) -> Workspace:
    _require_cuda(pred, K, T, source, target)
    B, _, H, W = pred.shape
    if ws is None:
        ws = Workspace(B, H, W, pred.device)
    ws.grad = torch.empty_like(pred) if want_grad else None
    kb = 1 if K.dim() == 3 else 0
    with torch.cuda.device(pred.device):
        rc = lib().cadl_photometric_fwd_bwd(_ptr(pred), _ptr(K), kb, _ptr(T), _ptr(source), _ptr(target), B, H, W,
                                            eps, upstream, _ptr(ws.grad), _ptr(ws.results), _ptr(ws.buf), ws.bytes,
                                            _stream(pred))
    _check(rc, "cadl_photometric_fwd_bwd")
    return ws


def batch_prep(rgb, depth, K, H: int, W: int):
    """SunRGBDLoader::resizeSample on the device: (rgb bilinear, depth nearest, K rescaled) -> (B,.,H,W)."""
    _require_cuda(rgb, depth, K)
    B, _, h, w = rgb.shape
    rgb_o = torch.empty(B, 3, H, W, dtype=torch.float32, device=rgb.device)
    dep_o = torch.empty(B, 1, H, W, dtype=torch.float32, device=rgb.device)
    K_o = torch.empty(B, 3, 3, dtype=torch.float32, device=rgb.device)
    with torch.cuda.device(rgb.device):
        rc = lib().cadl_batch_prep(_ptr(rgb), _ptr(depth), _ptr(K), B, h, w, H, W, _ptr(rgb_o), _ptr(dep_o), _ptr(K_o),
                                   _stream(rgb))
    _check(rc, "cadl_batch_prep")
    return rgb_o, dep_o, K_o


def batch_augment(rgb, depth, K, aug, H: int, W: int):
    """augmentSample + resize on the device; aug: (B, 8) float32 CUDA tensor (see include/cadl.h: cadl_batch_augment)."""
    _require_cuda(rgb, depth, K, aug)
    B, _, h, w = rgb.shape
    assert aug.shape == (B, 8) and aug.dtype == torch.float32 and aug.is_contiguous()
    rgb_o = torch.empty(B, 3, H, W, dtype=torch.float32, device=rgb.device)
    dep_o = torch.empty(B, 1, H, W, dtype=torch.float32, device=rgb.device)
    K_o = torch.empty(B, 3, 3, dtype=torch.float32, device=rgb.device)
    with torch.cuda.device(rgb.device):
        rc = lib().cadl_batch_augment(_ptr(rgb), _ptr(depth), _ptr(K), _ptr(aug), B, h, w, H, W, _ptr(rgb_o), _ptr(dep_o),
                                      _ptr(K_o), _stream(rgb))
    _check(rc, "cadl_batch_augment")
    return rgb_o, dep_o, K_o


def accumulate(values, weight: float, acc):
    """acc[:n] += weight * values, acc[n] += weight, on the device (no host sync)."""
    _require_cuda(values, acc)
    n = values.numel()
    assert values.dtype == torch.float32 and acc.dtype == torch.float64 and acc.numel() >= n + 1
    L = lib()
    L.cadl_accumulate.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
    with torch.cuda.device(values.device):
        rc = L.cadl_accumulate(_ptr(values), n, float(weight), _ptr(acc), _stream(values))
    _check(rc, "cadl_accumulate")


class GradClipper:
    """clip_grad_norm_ over a fixed list of gradient tensors, no host sync: the device arrays of pointers and
    sizes are built once (per model), each call is two launches; total norm and coefficient stay on the device."""

    CHUNK = 4096

    def __init__(self, grads):
        grads = [g for g in grads if g is not None]
        assert grads and all(g.is_cuda and g.dtype == torch.float32 and g.is_contiguous() for g in grads)
        self.grads = grads
        dev = grads[0].device
        sizes = [g.numel() for g in grads]
        prefix = [0]
        for n in sizes:
            prefix.append(prefix[-1] + (n + self.CHUNK - 1) // self.CHUNK)
        self.total_chunks = prefix[-1]
        self.ptrs = torch.tensor([g.data_ptr() for g in grads], dtype=torch.int64, device=dev)
        self.sizes = torch.tensor(sizes, dtype=torch.int64, device=dev)
        self.prefix = torch.tensor(prefix, dtype=torch.int64, device=dev)
        self.out = torch.zeros(2, dtype=torch.float32, device=dev)          # total norm, clip coefficient
        self.ws_bytes = int(lib().cadl_clip_workspace_bytes())
        self.ws = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=dev)

    def __call__(self, max_norm: float, clip: bool = True) -> torch.Tensor:
        dev = self.out.device
        with torch.cuda.device(dev):
            rc = lib().cadl_clip_grad_norm(_ptr(self.ptrs), _ptr(self.sizes), _ptr(self.prefix), len(self.grads),
                                           self.total_chunks, float(max_norm), _ptr(self.out), _ptr(self.ws), self.ws_bytes,
                                           1 if clip else 0, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _check(rc, "cadl_clip_grad_norm")
        return self.out


from .harness import StepHarness, StepCfg  # noqa: E402
from . import synth  # noqa: E402
from . import multi  # noqa: E402
from .rays_io import save_ray_directions, load_ray_directions  # noqa: E402


def host_harness() -> StepHarness:
    """The drop-in C++ classes (host/loss/depth_loss.h ...) behind the trainer-shaped C harness."""
    if not os.path.exists(LIBHOST_PATH):
        raise RuntimeError(f"cadl: {LIBHOST_PATH} is missing -- run __graft_entry__.build()")
    lib()  # libcadl.so first, so the host library resolves against the in-tree build
    return StepHarness(LIBHOST_PATH)
