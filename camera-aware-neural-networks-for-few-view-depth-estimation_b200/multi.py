"""Multi-GPU host logic (one process per GPU, torch.distributed).

The path shards by image along B: every kernel is per-pixel or per-image-stencil, so the
only coupling between ranks is the statistics vector phase A produces (counts and sums, 32 doubles).

  mode "local"  (default)  each rank evaluates the reference loss on its own B/N images -- what DDP would
                           do with the reference -- and no data-path collective exists;
  mode "global"            exact global-batch loss: all-reduce(sum) of the statistics vector between
                           phase A and phase B (cadl_stack_reduce / cadl_stack_grad), then the additive
                           shares of the two stencil terms are summed for logging.  cadl_stack_prepare first
                           puts the pooled-pyramid kernels on an auxiliary stream, where they run beside
                           phase A and the all-reduce instead of in front of the gradient pass.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

# indices into the statistics vector (csrc/cadl_common.cuh enum Stat)
ST_SI_N, ST_SI_S, ST_SI_Q, ST_RP_N = 0, 1, 2, 3
ST_COUNT = 32


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of `total` images for `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def exchange_stats(stats: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Sum the per-rank statistics vectors in place (NCCL on device tensors, gloo on CPU tensors)."""
    assert stats.dtype == torch.float64 and stats.numel() == ST_COUNT
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def si_from_stats(stats: torch.Tensor, lam: float = 0.5) -> float:
    """ScaleInvariantLoss from (n, sum d, sum d^2): depth_loss.h:58-63 (zero when n == 0, :53-55)."""
    n, S, Q = float(stats[ST_SI_N]), float(stats[ST_SI_S]), float(stats[ST_SI_Q])
    return 0.0 if n == 0 else Q / n - lam * S * S / (n * n)


def combine_shares(local: Dict[str, float], weights=(1.0, 0.1, 0.001, 0.01),
                   group: Optional[dist.ProcessGroup] = None, device="cpu") -> Dict[str, float]:
    """In mode "global" the SI and reprojection losses a rank reports are already global, while the
    gradient-matching and smoothness losses are this rank's additive share (their denominators use the
    global batch size).  Sum the shares and rebuild the weighted total (depth_loss.h:427-430)."""
    t = torch.tensor([local["d_grad"], local["d_smooth"], local.get("reproj_sum_e", 0.0)], dtype=torch.float64,
                     device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    out = dict(local)
    out["d_grad"], out["d_smooth"] = float(t[0]), float(t[1])
    w = weights
    out["d_total"] = w[0] * out["d_si"] + w[1] * out["d_grad"] + w[2] * out["d_smooth"] + w[3] * out["d_reproj"]
    return out
