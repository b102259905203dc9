"""ctypes view of the trainer-shaped step harness (host/harness/step_harness.cpp).

The same C surface is exported by two builds of that one source file:
  * host/libcadl_host.so                -- the drop-in headers (product path, CUDA only);
  * oracle/_ref/libcadl_refharness.so   -- the unmodified reference headers on LibTorch (oracle).
All buffers crossing it are HOST numpy arrays; the harness moves them to its device the way the
reference trainers do (src/training/production_trainer.h:192-194).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

TERM_NAMES = {0: "combined+K", 1: "si", 2: "grad", 3: "smooth", 4: "reproj", 5: "combined"}


class _Cfg(C.Structure):
    _fields_ = [("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("k_batched", C.c_int), ("device", C.c_int),
                ("w_si", C.c_float), ("w_grad", C.c_float), ("w_smooth", C.c_float), ("w_reproj", C.c_float),
                ("term", C.c_int), ("upstream", C.c_float)]


@dataclass
class StepCfg:
    device: int = -1          # -1 CPU, >= 0 CUDA ordinal
    term: int = 0             # see TERM_NAMES
    w_si: float = 1.0
    w_grad: float = 0.1
    w_smooth: float = 0.001
    w_reproj: float = 0.01
    upstream: float = 1.0


class _UNetCfg(C.Structure):
    _fields_ = [("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("device", C.c_int), ("feats", C.c_int),
                ("max_depth", C.c_float), ("lr", C.c_float), ("clip", C.c_float), ("fused_extras", C.c_int),
                ("world", C.c_int), ("rank", C.c_int), ("bucket_mb", C.c_int)]


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class StepHarness:
    def __init__(self, path: str):
        self.path = path
        L = C.CDLL(path)
        L.cadh_build_info.restype = C.c_char_p
        vp = C.c_void_p
        L.cadh_loss_step.argtypes = [C.POINTER(_Cfg), vp, vp, vp, vp, vp, vp, vp, vp, C.c_char_p, C.c_int]
        L.cadh_components.argtypes = [C.POINTER(_Cfg), vp, vp, vp, vp, vp, vp, C.c_char_p, C.c_int]
        L.cadh_metrics_eval.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_float, C.c_float, vp, vp,
                                        C.c_char_p, C.c_int]
        L.cadh_metrics_train.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.c_char_p, C.c_int]
        L.cadh_time_steps.argtypes = [C.POINTER(_Cfg), vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp,
                                      C.c_char_p, C.c_int]
        L.cadh_set_num_threads.argtypes = [C.c_int]
        if hasattr(L, "cadh_clip_grad_norm"):        # drop-in build only ("next" rows)
            L.cadh_clip_grad_norm.argtypes = [C.c_int, C.c_int, vp, vp, C.c_float, C.c_int, vp, vp, C.c_char_p, C.c_int]
            L.cadh_batch_prep.argtypes = [C.c_int] * 6 + [vp] * 6 + [C.c_char_p, C.c_int]
            L.cadh_accumulate.argtypes = [C.c_int, C.c_int, vp, vp, vp, C.c_char_p, C.c_int]
            L.cadh_empty_rank.argtypes = [C.c_int] * 4 + [vp] * 4 + [C.c_char_p, C.c_int]
            L.cadh_rays_roundtrip.argtypes = [C.c_int, C.c_int, C.c_int, vp, C.c_char_p, C.c_int, vp, vp, vp, C.c_char_p, C.c_int]
            L.cadh_photometric_step.argtypes = [C.c_int] * 4 + [vp] * 5 + [C.c_float, vp, vp, C.c_char_p, C.c_int]
        L.cadh_metric_utils.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, vp, C.c_char_p,
                                        C.c_int, C.c_char_p, C.c_int]
        if hasattr(L, "cadh_unet_train"):            # builds that saw the reference's model header
            L.cadh_unet_train.argtypes = [C.POINTER(_UNetCfg), vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_char_p, C.c_int]
        if hasattr(L, "cadh_nccl_init"):
            L.cadh_nccl_unique_id.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
            L.cadh_nccl_init.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        self.L = L

    # ------------------------------------------------------------------
    def has_unet(self) -> bool:
        return hasattr(self.L, "cadh_unet_train") and bool(self.L.cadh_has_unet())

    def has_nccl(self) -> bool:
        return hasattr(self.L, "cadh_nccl_init") and bool(self.L.cadh_has_nccl())

    def nccl_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        err = C.create_string_buffer(2048)
        if self.L.cadh_nccl_unique_id(buf, err, len(err)):
            self._raise(err)
        return buf.raw

    def nccl_init(self, rank: int, world: int, unique_id: bytes, device: int):
        err = C.create_string_buffer(2048)
        if self.L.cadh_nccl_init(rank, world, unique_id, device, err, len(err)):
            self._raise(err)

    def nccl_finalize(self):
        self.L.cadh_nccl_finalize()

    def unet_train(self, rgb, gt, K, device: int = -1, feats: int = 64, max_depth: float = 10.0, lr: float = 1e-4,
                   clip: float = 1.0, fused_extras: bool = False, world: int = 1, rank: int = 0, bucket_mb: int = 25,
                   warmup: int = 1, iters: int = 3):
        """BaselineUNet training-shaped steps (harness/unet_step.inc).  Returns dict(ms, loss_ms, last_loss, params,
        grad_bytes, buckets)."""
        rgb, gt, K = _f32(rgb), _f32(gt), _f32(K)
        B, _, H, W = rgb.shape
        cfg = _UNetCfg(B, H, W, device, feats, max_depth, lr, clip, int(fused_extras), world, rank, bucket_mb)
        ms, lms = (C.c_double * iters)(), (C.c_double * iters)()
        last = C.c_float(0)
        info = (C.c_longlong * 3)()
        err = C.create_string_buffer(2048)
        rc = self.L.cadh_unet_train(C.byref(cfg), _p(rgb), _p(gt), _p(K), warmup, iters, ms, lms, C.byref(last), info, err, len(err))
        if rc:
            self._raise(err)
        return {"ms": [float(x) for x in ms], "loss_ms": [float(x) for x in lms], "last_loss": float(last.value),
                "params": int(info[0]), "grad_bytes": int(info[1]), "buckets": int(info[2])}

    def metric_utils(self, device: int, pred, gt, mask=None, splits: int = 2):
        """computePerSample / average / MetricsAccumulator / formatMetrics (depth_metrics.h:93-141, 259-333)."""
        pred, gt = _f32(pred), _f32(gt)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        B, _, H, W = pred.shape
        per = np.empty((B, 12), np.float32)
        avg, acc = np.empty(12, np.float32), np.empty(12, np.float32)
        cnt = (C.c_int * 2)()
        text = C.create_string_buffer(4096)
        err = C.create_string_buffer(2048)
        rc = self.L.cadh_metric_utils(B, H, W, device, _p(pred), _p(gt), _p(m), splits, _p(per), _p(avg), _p(acc), cnt, text,
                                      len(text), err, len(err))
        if rc:
            self._raise(err)
        return {"per_sample": per, "average": avg, "accumulated": acc, "count": (int(cnt[0]), int(cnt[1])),
                "text": text.value.decode()}

    def info(self) -> str:
        return self.L.cadh_build_info().decode()

    def is_dropin(self) -> bool:
        return bool(self.L.cadh_is_dropin())

    def num_threads(self) -> int:
        return int(self.L.cadh_num_threads())

    def set_num_threads(self, n: int):
        self.L.cadh_set_num_threads(int(n))

    def cuda_available(self) -> bool:
        return bool(self.L.cadh_cuda_available())

    def _cfg(self, cfg: StepCfg, pred, K) -> _Cfg:
        B, _, H, W = pred.shape
        kb = 1 if (K is not None and K.ndim == 3) else 0
        return _Cfg(B, H, W, kb, cfg.device, cfg.w_si, cfg.w_grad, cfg.w_smooth, cfg.w_reproj, cfg.term, cfg.upstream)

    @staticmethod
    def _raise(err):
        raise RuntimeError(err.value.decode(errors="replace"))

    # ------------------------------------------------------------------
    def loss_step(self, cfg: StepCfg, pred, gt, rgb=None, K=None, mask=None, want_grad: bool = True):
        """One training-shaped step.  Returns (loss float, loss rank, loss numel, grad ndarray|None)."""
        pred, gt, rgb, K = _f32(pred), _f32(gt), _f32(rgb), _f32(K)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        c = self._cfg(cfg, pred, K)
        loss = C.c_float(0)
        meta = (C.c_int64 * 2)()
        grad = np.empty_like(pred) if want_grad else None
        err = C.create_string_buffer(2048)
        rc = self.L.cadh_loss_step(C.byref(c), _p(pred), _p(gt), _p(rgb), _p(K), _p(m), C.byref(loss), meta, _p(grad),
                                   err, len(err))
        if rc:
            self._raise(err)
        return float(loss.value), int(meta[0]), int(meta[1]), grad

    def components(self, cfg: StepCfg, pred, gt, rgb, K, mask=None):
        pred, gt, rgb, K = _f32(pred), _f32(gt), _f32(rgb), _f32(K)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        c = self._cfg(cfg, pred, K)
        out = (C.c_float * 4)()
        err = C.create_string_buffer(2048)
        rc = self.L.cadh_components(C.byref(c), _p(pred), _p(gt), _p(rgb), _p(K), _p(m), out, err, len(err))
        if rc:
            self._raise(err)
        return {"si_loss": out[0], "grad_loss": out[1], "smooth_loss": out[2], "reproj_loss": out[3]}

    def metrics_eval(self, device: int, pred, gt, mask=None, min_depth=0.1, max_depth=10.0):
        pred, gt = _f32(pred), _f32(gt)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        B, _, H, W = pred.shape
        out = (C.c_float * 12)()
        cnt = (C.c_int64 * 4)()
        err = C.create_string_buffer(2048)
        rc = self.L.cadh_metrics_eval(B, H, W, device, _p(pred), _p(gt), _p(m), min_depth, max_depth, out, cnt, err,
                                      len(err))
        if rc:
            self._raise(err)
        return [float(x) for x in out], [int(x) for x in cnt]

    def metrics_train(self, device: int, pred, gt):
        pred, gt = _f32(pred), _f32(gt)
        B, _, H, W = pred.shape
        out = (C.c_float * 7)()
        cnt = (C.c_int64 * 4)()
        err = C.create_string_buffer(2048)
        rc = self.L.cadh_metrics_train(B, H, W, device, _p(pred), _p(gt), out, cnt, err, len(err))
        if rc:
            self._raise(err)
        return [float(x) for x in out], [int(x) for x in cnt]

    def clip_grad_norm(self, device: int, grads, max_norm: float, clip: bool = True):
        """FusedGradClipper (host/training/grad_clip.h) over a list of gradient arrays -> (norm, coef, clipped)."""
        gs = [_f32(g).reshape(-1) for g in grads]
        outs = [np.empty_like(g) for g in gs]
        n = len(gs)
        inp = (C.c_void_p * n)(*[g.ctypes.data for g in gs])
        outp = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        sizes = (C.c_int64 * n)(*[g.size for g in gs])
        out2 = (C.c_float * 2)()
        err = C.create_string_buffer(2048)
        rc = self.L.cadh_clip_grad_norm(device, n, inp, sizes, max_norm, int(clip), out2, outp, err, len(err))
        if rc:
            self._raise(err)
        return float(out2[0]), float(out2[1]), outs

    def accumulate(self, device: int, values, weights) -> float:
        """DeviceAccumulator (host/training/loss_accumulator.h): weighted mean of per-batch losses."""
        v = np.ascontiguousarray(values, dtype=np.float32)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        out = C.c_double(0.0)
        err = C.create_string_buffer(2048)
        rc = self.L.cadh_accumulate(device, len(v), _p(v), w.ctypes.data_as(C.c_void_p), C.byref(out), err, len(err))
        if rc:
            self._raise(err)
        return float(out.value)

    def empty_rank(self, device: int, pred, gt, K):
        """dim() of ScaleInvariantLoss / ReprojectionLoss results: (SI, SI opted in, reproj, reproj opted in)."""
        pred, gt, K = _f32(pred), _f32(gt), _f32(K)
        B, _, H, W = pred.shape
        out = (C.c_int * 4)()
        err = C.create_string_buffer(2048)
        if self.L.cadh_empty_rank(B, H, W, device, _p(pred), _p(gt), _p(K), out, err, len(err)):
            self._raise(err)
        return tuple(int(x) for x in out)

    def rays_roundtrip(self, device: int, K33, H: int, W: int, path: str, corrupt_dims: bool = False):
        """RayDirectionComputer::computeRayDirections -> saveRayDirections -> loadRayDirections (host/preprocessing).
        Returns (saved?, rays (H*W,3) or None, (h, w))."""
        K33 = _f32(K33)
        out = np.empty((H * W, 3), np.float32)
        saved = C.c_int(0)
        hw = (C.c_int * 2)()
        err = C.create_string_buffer(2048)
        rc = self.L.cadh_rays_roundtrip(device, H, W, _p(K33), path.encode(), int(corrupt_dims), _p(out), C.byref(saved), hw,
                                        err, len(err))
        if rc:
            self._raise(err)
        return bool(saved.value), (out if saved.value else None), (int(hw[0]), int(hw[1]))

    def photometric_step(self, device: int, pred, K, T, src, tgt, upstream: float = 1.0):
        """ReprojectionLoss::forwardPhotometricWarp + backward (host/loss/depth_loss.h) -> (loss, dL/dpred)."""
        pred, K, T, src, tgt = _f32(pred), _f32(K), _f32(T), _f32(src), _f32(tgt)
        B, _, H, W = pred.shape
        loss = C.c_float(0)
        grad = np.empty_like(pred)
        err = C.create_string_buffer(2048)
        rc = self.L.cadh_photometric_step(B, H, W, device, _p(pred), _p(K), _p(T), _p(src), _p(tgt), upstream,
                                          C.byref(loss), _p(grad), err, len(err))
        if rc:
            self._raise(err)
        return float(loss.value), grad

    def batch_prep(self, device: int, rgb, depth, K, H: int, W: int):
        """resizeBatchOnDevice (host/data/batch_prep.h)."""
        rgb, depth, K = _f32(rgb), _f32(depth), _f32(K)
        B, _, h, w = rgb.shape
        ro = np.empty((B, 3, H, W), np.float32)
        do = np.empty((B, 1, H, W), np.float32)
        ko = np.empty((B, 3, 3), np.float32)
        err = C.create_string_buffer(2048)
        rc = self.L.cadh_batch_prep(device, B, h, w, H, W, _p(rgb), _p(depth), _p(K), _p(ro), _p(do), _p(ko), err, len(err))
        if rc:
            self._raise(err)
        return ro, do, ko

    def time_steps(self, cfg: StepCfg, pred, gt, rgb, K, mask=None, with_metrics=False, include_h2d=False,
                   warmup=1, iters=5):
        """Per-iteration wall milliseconds of [h2d] forward+backward [+metrics] + loss.item()."""
        pred, gt, rgb, K = _f32(pred), _f32(gt), _f32(rgb), _f32(K)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        c = self._cfg(cfg, pred, K)
        ms = (C.c_double * iters)()
        last = C.c_float(0)
        err = C.create_string_buffer(2048)
        rc = self.L.cadh_time_steps(C.byref(c), _p(pred), _p(gt), _p(rgb), _p(K), _p(m), int(with_metrics),
                                    int(include_h2d), warmup, iters, ms, C.byref(last), err, len(err))
        if rc:
            self._raise(err)
        return [float(x) for x in ms], float(last.value)
