"""ORACLE (test infrastructure only; never imported by the product package).

Op-for-op torch restatement ("port") of the reference's depth-loss hot path.  Every function cites
the reference lines it follows.  It is a *port*, so it is pinned before being trusted:

  * against the unmodified reference headers compiled here (oracle/_ref/libcadl_refharness.so,
    see oracle/Makefile) -- tests/test_oracle_pin.py, bit-identical on CPU;
  * against the golden vectors committed under tests/golden/ that were generated from that
    reference build by tests/golden/make_golden.py.

The reference's own tests hold no golden vector for this path (tests/test_models.cpp:365-509 pin
rank and sign only), so those two are what "pinned" means here.

Works on CPU or CUDA tensors, fp32 (parity) or fp64 (truth bound).  autograd supplies backward,
exactly as in the reference.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# a1  ScaleInvariantLoss::forward            reference src/loss/depth_loss.h:33-64
# --------------------------------------------------------------------------------------------
def scale_invariant_loss(pred, gt, valid_mask: Optional[torch.Tensor] = None,
                         lam: float = 0.5, eps: float = 1e-6):
    mask = valid_mask if valid_mask is not None else (gt > eps)          # :38-40
    pred = torch.clamp(pred, eps, 1000.0)                                 # :43
    gt = torch.clamp(gt, eps, 1000.0)                                     # :44
    log_diff = torch.log(pred) - torch.log(gt)                            # :47
    masked_diff = log_diff.masked_select(mask)                            # :50
    n = masked_diff.numel()                                               # :52
    if n == 0:
        return torch.zeros(1, dtype=pred.dtype, device=pred.device)       # :53-55
    term1 = torch.pow(masked_diff, 2).sum() / n                           # :58
    term2 = lam * torch.pow(masked_diff.sum(), 2) / (n * n)               # :61
    return term1 - term2                                                  # :63


# --------------------------------------------------------------------------------------------
# a2  GradientMatchingLoss::forward          reference src/loss/depth_loss.h:95-124,135-166
# --------------------------------------------------------------------------------------------
def _gradient_loss_one_scale(pred, gt):
    pgx = pred[:, :, :, 1:] - pred[:, :, :, :-1]                          # :140-143
    ggx = gt[:, :, :, 1:] - gt[:, :, :, :-1]                              # :145-148
    pgy = pred[:, :, 1:, :] - pred[:, :, :-1, :]                          # :151-154
    ggy = gt[:, :, 1:, :] - gt[:, :, :-1, :]                              # :156-159
    loss_x = torch.abs(pgx - ggx).mean()                                  # :162
    loss_y = torch.abs(pgy - ggy).mean()                                  # :163
    return loss_x + loss_y                                                # :165


def gradient_matching_loss(pred, gt, valid_mask=None, num_scales: int = 4, eps: float = 1e-6):
    # valid_mask is accepted and ignored, as in the reference (:137 -- never used in the body)
    total = torch.zeros(1, dtype=pred.dtype, device=pred.device)          # :99
    for scale in range(num_scales):                                       # :101
        ps, gs = pred, gt
        if scale > 0:
            f = int(math.pow(2, scale))                                   # :107
            ps = F.avg_pool2d(pred, f, stride=f)                          # :108-109
            gs = F.avg_pool2d(gt, f, stride=f)                            # :110-111
        ps = torch.log(torch.clamp(ps, eps, 1000.0))                      # :115
        gs = torch.log(torch.clamp(gs, eps, 1000.0))                      # :116
        total = total + _gradient_loss_one_scale(ps, gs)                  # :119-120
    return total / num_scales                                             # :123


# --------------------------------------------------------------------------------------------
# a3  SmoothnessLoss::forward                reference src/loss/depth_loss.h:189-234
# --------------------------------------------------------------------------------------------
def smoothness_loss(pred, image, eps: float = 1e-6):
    depth_mean = pred.mean(dim=(2, 3), keepdim=True)                      # :192
    depth_norm = pred / (depth_mean + eps)                                # :193
    dgx = torch.abs(depth_norm[:, :, :, 1:] - depth_norm[:, :, :, :-1])   # :196-201
    dgy = torch.abs(depth_norm[:, :, 1:, :] - depth_norm[:, :, :-1, :])   # :203-208
    igx = torch.abs(image[:, :, :, 1:] - image[:, :, :, :-1]).mean(1, keepdim=True)   # :211-216
    igy = torch.abs(image[:, :, 1:, :] - image[:, :, :-1, :]).mean(1, keepdim=True)   # :218-223
    wx = torch.exp(-igx)                                                  # :226
    wy = torch.exp(-igy)                                                  # :227
    return (dgx * wx).mean() + (dgy * wy).mean()                          # :230-233


# --------------------------------------------------------------------------------------------
# a4  ReprojectionLoss::forward              reference src/loss/depth_loss.h:268-331
# --------------------------------------------------------------------------------------------
def reprojection_loss(pred, gt, intrinsics, valid_mask=None, eps: float = 1e-6):
    B, _, H, W = pred.shape                                               # :273-275
    if intrinsics.dim() == 2:
        intrinsics = intrinsics.unsqueeze(0).expand(B, 3, 3)              # :278-280
    gy = torch.arange(0, H, dtype=pred.dtype, device=pred.device).view(1, H, 1).expand(1, H, W)  # :283
    gx = torch.arange(0, W, dtype=pred.dtype, device=pred.device).view(1, 1, W).expand(1, H, W)  # :284
    fx = intrinsics[:, 0, 0].reshape(B, 1, 1, 1)                          # :290
    fy = intrinsics[:, 1, 1].reshape(B, 1, 1, 1)                          # :291
    cx = intrinsics[:, 0, 2].reshape(B, 1, 1, 1)                          # :292
    cy = intrinsics[:, 1, 2].reshape(B, 1, 1, 1)                          # :293
    pX = (gx - cx) * pred / (fx + eps)                                    # :299
    pY = (gy - cy) * pred / (fy + eps)                                    # :300
    pZ = pred                                                             # :301
    gX = (gx - cx) * gt / (fx + eps)                                      # :304
    gY = (gy - cy) * gt / (fy + eps)                                      # :305
    gZ = gt                                                               # :306
    dX, dY, dZ = pX - gX, pY - gY, pZ - gZ                                # :309-311
    err = torch.sqrt(dX * dX + dY * dY + dZ * dZ + eps)                   # :313-315
    mask = valid_mask if valid_mask is not None else (gt > eps)           # :318-320
    masked = err.masked_select(mask)                                      # :323
    if masked.numel() == 0:
        return torch.zeros(1, dtype=pred.dtype, device=pred.device)       # :325-327
    return masked.mean()                                                  # :330


def photometric_stub(pred, *_):
    """ReprojectionLoss::forwardPhotometric -- a stub in the reference (depth_loss.h:343-351)."""
    return torch.zeros(1, dtype=pred.dtype, device=pred.device)


# --------------------------------------------------------------------------------------------
# a5  CombinedDepthLoss                      reference src/loss/depth_loss.h:366-479
# --------------------------------------------------------------------------------------------
def combined_loss(pred, gt, image, intrinsics=None, valid_mask=None,
                  w_si=1.0, w_grad=0.1, w_smooth=0.001, w_reproj=0.01):
    si = scale_invariant_loss(pred, gt, valid_mask)                       # :395 / :422
    grad = gradient_matching_loss(pred, gt, valid_mask)                   # :396 / :423
    smooth = smoothness_loss(pred, image)                                 # :397 / :424
    total = w_si * si + w_grad * grad + w_smooth * smooth                 # :399-401 / :427-429
    comps = {"si_loss": si, "grad_loss": grad, "smooth_loss": smooth}
    if intrinsics is not None:
        reproj = reprojection_loss(pred, gt, intrinsics, valid_mask)      # :425
        total = total + w_reproj * reproj                                 # :430
        comps["reproj_loss"] = reproj
    return total, comps


# --------------------------------------------------------------------------------------------
# a8  DepthMetrics::compute                  reference src/evaluation/depth_metrics.h:40-88,147-253
# --------------------------------------------------------------------------------------------
EVAL_KEYS = ("abs_rel", "sq_rel", "rmse", "rmse_log", "mae", "log10",
             "delta_1.25", "delta_1.25^2", "delta_1.25^3",
             "num_valid_pixels", "mean_pred_depth", "mean_gt_depth")


def _f32(x: float) -> float:
    return float(torch.tensor(x, dtype=torch.float32))


def metrics_eval(pred, gt, valid_mask=None, min_depth: float = 0.1, max_depth: float = 10.0):
    """Returns (dict of 12 floats, int counts [n_valid, n_d1, n_d2, n_d3])."""
    if pred.dim() == 3:
        pred = pred.unsqueeze(1)                                          # :50
    if gt.dim() == 3:
        gt = gt.unsqueeze(1)                                              # :51
    mask = (gt > min_depth) & (gt < max_depth)                            # :154
    if valid_mask is not None:
        um = valid_mask
        if um.dim() == 3:
            um = um.unsqueeze(1)                                          # :159
        mask = mask & um.to(torch.bool)                                   # :160
    p = pred.masked_select(mask)                                          # :57
    g = gt.masked_select(mask)                                            # :58
    n = p.numel()                                                         # :60
    if n == 0:
        return {k: 0.0 for k in EVAL_KEYS}, [0, 0, 0, 0]                  # :61-63,238-253
    p = torch.clamp(p, min_depth, max_depth)                              # :66
    out: Dict[str, float] = {}
    out["abs_rel"] = float((torch.abs(p - g) / g).mean())                 # :170
    out["sq_rel"] = float((torch.pow(p - g, 2) / g).mean())               # :177
    out["rmse"] = float(torch.sqrt(torch.pow(p - g, 2).mean()))           # :184
    ld = torch.log(p) - torch.log(g)                                      # :191
    out["rmse_log"] = float(torch.sqrt(torch.pow(ld, 2).mean()))          # :192
    out["mae"] = float(torch.abs(p - g).mean())                           # :199
    out["log10"] = float(torch.abs(torch.log10(p) - torch.log10(g)).mean())   # :206
    ratio = torch.max(p / g, g / p)                                       # :221
    # thresholds are float products 1.25f, 1.25f*1.25f, 1.25f*1.25f*1.25f (:224) -- all exact
    thr = (1.25, 1.5625, 1.953125)
    counts = [n]
    for key, t in zip(("delta_1.25", "delta_1.25^2", "delta_1.25^3"), thr):
        below = ratio < t                                                 # :228
        out[key] = float(below.to(torch.float32).mean())                  # :228-229
        counts.append(int(below.sum()))
    out["num_valid_pixels"] = _f32(float(n))                              # :83 static_cast<float>
    out["mean_pred_depth"] = float(p.mean())                              # :84
    out["mean_gt_depth"] = float(g.mean())                                # :85
    return out, counts


# --------------------------------------------------------------------------------------------
# a9  trainers' computeDepthMetrics          reference src/training/tensorboard_trainer_enhanced.h:400-439
# --------------------------------------------------------------------------------------------
TRAIN_KEYS = ("abs_rel", "sq_rel", "rmse", "rmse_log", "a1", "a2", "a3")


def metrics_train(pred, gt):
    pf = pred.reshape(-1)                                                 # :406
    gf = gt.reshape(-1)                                                   # :407
    vm = gf > 0.0                                                         # :410
    p = pf.masked_select(vm)                                              # :411
    g = gf.masked_select(vm)                                              # :412
    if p.numel() == 0:
        return {k: 0.0 for k in TRAIN_KEYS}, [0, 0, 0, 0]                 # :414-416
    ad = torch.abs(p - g)                                                 # :419
    out = {}
    out["abs_rel"] = float((ad / g).mean())                               # :420
    out["sq_rel"] = float(((ad * ad) / g).mean())                         # :423
    out["rmse"] = float(torch.sqrt((ad * ad).mean()))                     # :426
    ld = torch.abs(torch.log(p + 1e-8) - torch.log(g + 1e-8))             # :429
    out["rmse_log"] = float(torch.sqrt((ld * ld).mean()))                 # :430
    ratio = torch.max(p / g, g / p)                                       # :433
    counts = [p.numel()]
    for key, t in zip(("a1", "a2", "a3"), (1.25, 1.5625, 1.953125)):      # :434-436
        below = ratio < t
        out[key] = float(below.to(torch.float32).mean())
        counts.append(int(below.sum()))
    return out, counts


# --------------------------------------------------------------------------------------------
# a10 MetricsAccumulator / average           reference src/evaluation/depth_metrics.h:122-141,259-304
# --------------------------------------------------------------------------------------------
class MetricsAccumulator:
    def __init__(self):
        self.running_sum: Dict[str, float] = {}
        self.count_ = 0                                                   # :261

    def update(self, metrics: Dict[str, float]):                          # :266-271
        for k, v in metrics.items():
            self.running_sum[k] = _f32(self.running_sum.get(k, 0.0) + _f32(v))
        self.count_ += 1

    def average(self) -> Dict[str, float]:                                # :276-286
        if self.count_ == 0:
            return {}
        return {k: _f32(s / self.count_) for k, s in self.running_sum.items()}

    def reset(self):                                                      # :291-294
        self.running_sum.clear()
        self.count_ = 0

    def count(self) -> int:                                               # :299
        return self.count_


# --------------------------------------------------------------------------------------------
# a7 extension oracle: real photometric reprojection.  NOT in the reference (its forwardPhotometric
# is a stub returning zeros, depth_loss.h:343-351): PARITY UNPINNED.  Formulas from
# documents/algorithms_and_theory.md:22-25,62-77; sampling convention from
# src/layers/pcl_layer.h:104-108 (grid_sample bilinear, zeros padding, align_corners=false).
# --------------------------------------------------------------------------------------------
def photometric_reprojection(pred, intrinsics, T, source_image, target_image, eps: float = 1e-6):
    """Warp `source_image` into the target view with pred depth, K and T=[R|t] (target->source),
    return mean over valid pixels of the channel-mean L1 photometric residual.

    X_t = d * K^-1 [u,v,1]^T ; X_s = R X_t + t ; [u',v'] = K X_s / Z_s ; I_s(u',v') bilinear.
    Valid = Z_s > eps and the sample point inside the source image.
    """
    B, _, H, W = pred.shape
    if intrinsics.dim() == 2:
        intrinsics = intrinsics.unsqueeze(0).expand(B, 3, 3)
    dt, dev = pred.dtype, pred.device
    v = torch.arange(0, H, dtype=dt, device=dev).view(1, H, 1).expand(1, H, W)
    u = torch.arange(0, W, dtype=dt, device=dev).view(1, 1, W).expand(1, H, W)
    fx = intrinsics[:, 0, 0].reshape(B, 1, 1)
    fy = intrinsics[:, 1, 1].reshape(B, 1, 1)
    cx = intrinsics[:, 0, 2].reshape(B, 1, 1)
    cy = intrinsics[:, 1, 2].reshape(B, 1, 1)
    d = pred[:, 0]
    X = (u - cx) / fx * d
    Y = (v - cy) / fy * d
    Z = d
    R = T[:, :3, :3]
    t = T[:, :3, 3]
    Xs = R[:, 0, 0].view(B, 1, 1) * X + R[:, 0, 1].view(B, 1, 1) * Y + R[:, 0, 2].view(B, 1, 1) * Z + t[:, 0].view(B, 1, 1)
    Ys = R[:, 1, 0].view(B, 1, 1) * X + R[:, 1, 1].view(B, 1, 1) * Y + R[:, 1, 2].view(B, 1, 1) * Z + t[:, 1].view(B, 1, 1)
    Zs = R[:, 2, 0].view(B, 1, 1) * X + R[:, 2, 1].view(B, 1, 1) * Y + R[:, 2, 2].view(B, 1, 1) * Z + t[:, 2].view(B, 1, 1)
    zok = Zs > eps
    Zc = torch.where(zok, Zs, torch.ones_like(Zs))
    us = fx * Xs / Zc + cx
    vs = fy * Ys / Zc + cy
    # pixel centres -> normalised grid, align_corners=false: g = (2*u + 1)/W - 1
    gx = (2.0 * us + 1.0) / W - 1.0
    gy = (2.0 * vs + 1.0) / H - 1.0
    grid = torch.stack([gx, gy], dim=-1)
    warped = F.grid_sample(source_image, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    inside = (us >= 0) & (us <= W - 1) & (vs >= 0) & (vs <= H - 1) & zok
    resid = torch.abs(warped - target_image).mean(1)
    m = inside
    sel = resid.masked_select(m)
    if sel.numel() == 0:
        return torch.zeros(1, dtype=dt, device=dev)
    return sel.mean()


# --------------------------------------------------------------------------------------------
# "next" rows.  The reference ops themselves: ATen interpolate / clip_grad_norm_.
# --------------------------------------------------------------------------------------------
def resize_sample(rgb, depth, K, H: int, W: int):
    """SunRGBDLoader::resizeSample, batched         reference src/data/sunrgbd_loader.cpp:445-489"""
    h, w = rgb.shape[-2:]
    rgb2 = F.interpolate(rgb, size=(H, W), mode="bilinear", align_corners=False)     # :453-459
    depth2 = F.interpolate(depth, size=(H, W), mode="nearest")                        # :461-467
    sx = torch.tensor(float(W), dtype=torch.float32) / w                              # :480  (float division)
    sy = torch.tensor(float(H), dtype=torch.float32) / h                              # :481
    K2 = K.clone()                                                                    # :483
    K2[:, 0, 0] = K[:, 0, 0] * sx                                                     # :484
    K2[:, 1, 1] = K[:, 1, 1] * sy                                                     # :485
    K2[:, 0, 2] = K[:, 0, 2] * sx                                                     # :486
    K2[:, 1, 2] = K[:, 1, 2] * sy                                                     # :487
    return rgb2, depth2, K2


def augment_resize_sample(rgb, depth, K, aug, H: int, W: int):
    """SunRGBDLoader::augmentSample followed by resizeSample, one sample      reference src/data/sunrgbd_loader.cpp:352-384,161-166
    rgb (3,h,w), depth (1,h,w), K (3,3); aug = [crop_x, crop_y, crop_w, crop_h, flip, jitter, contrast, brightness]."""
    cx, cy, cw, ch = (int(v) for v in aug[:4])
    K = K.clone()
    if cw > 0:                                                                         # applyCrop :388-415
        rgb = rgb[:, cy:cy + ch, cx:cx + cw]
        depth = depth[:, cy:cy + ch, cx:cx + cw]
        K[0, 2] = K[0, 2] - cx                                                         # :412
        K[1, 2] = K[1, 2] - cy                                                         # :413
    if aug[4] != 0:                                                                    # applyHorizontalFlip :417-432
        rgb = torch.flip(rgb, [2])
        depth = torch.flip(depth, [2])
        Wc = rgb.shape[2]
        K[0, 2] = torch.tensor(float(Wc), dtype=torch.float32) - K[0, 2] - 1           # :430  W - cx - 1
    if aug[5] != 0:                                                                    # applyColorJitter :434-443
        c = torch.tensor(float(aug[6]), dtype=torch.float32)
        bconst = torch.tensor(float(aug[7]), dtype=torch.float32)
        rgb = torch.clamp(rgb * c + bconst - 1.0, 0.0, 1.0)                            # :442
    r, d, k = resize_sample(rgb[None].contiguous(), depth[None].contiguous(), K[None], H, W)
    return r[0], d[0], k[0]


def clip_grad_norm(grads, max_norm: float):
    """torch::nn::utils::clip_grad_norm_ (tensorboard_trainer_enhanced.h:300-302) and computeGradientNorm (:560-571).
    Returns (total_norm, clipped grads)."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return total, [g * coef for g in grads]
