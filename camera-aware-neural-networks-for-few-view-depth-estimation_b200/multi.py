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


class P2PStatsExchange:
    """Mode "global" without a collective call: one small kernel per rank pushes the statistics vector into every
    rank's inbox over NVLink peer memory and sums the inboxes in rank order (include/cadl.h: cadl_stats_exchange).
    Setup (once): every rank allocates its inbox and the 64-byte CUDA IPC handles are all-gathered through
    torch.distributed; per step: ``exchange(ws)`` between stack_reduce and stack_grad, on the current stream."""

    def __init__(self, pkg, device: torch.device, group: Optional[dist.ProcessGroup] = None, timeout_s: float = 600.0):
        import ctypes as C
        self.timeout_s = float(timeout_s)
        self.pkg, self.C, self.L = pkg, C, pkg.lib()
        self.device = device
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.epoch = 0
        L = self.L
        L.cadl_p2p_alloc.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_char_p]
        L.cadl_p2p_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.cadl_p2p_close.argtypes = [C.c_void_p, C.c_int]
        L.cadl_stats_exchange.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_ulonglong, C.c_double, C.c_void_p]
        L.cadl_p2p_error.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int]
        with torch.cuda.device(device):
            own = C.c_void_p()
            handle = C.create_string_buffer(64)
            pkg._check(L.cadl_p2p_alloc(self.world, C.byref(own), handle), "cadl_p2p_alloc")
            self.own = own
            mine = torch.tensor(list(handle.raw), dtype=torch.uint8)
            if self.world > 1:
                gathered = [torch.zeros(64, dtype=torch.uint8, device=device) for _ in range(self.world)]
                dist.all_gather(gathered, mine.to(device), group=group)
                handles = [bytes(g.cpu().tolist()) for g in gathered]
            else:
                handles = [bytes(mine.tolist())]
            self.ptrs = (C.c_void_p * self.world)()
            for r in range(self.world):
                if r == self.rank:
                    self.ptrs[r] = own.value
                else:
                    q = C.c_void_p()
                    pkg._check(L.cadl_p2p_open(handles[r], C.byref(q)), "cadl_p2p_open")
                    self.ptrs[r] = q.value
            torch.cuda.synchronize(device)
        if self.world > 1:
            dist.barrier(group=group)          # every inbox is mapped everywhere before the first push

    def exchange(self, ws) -> None:
        """Sum the statistics vectors of all ranks in place (stream-ordered, no host sync)."""
        self.epoch += 1
        with torch.cuda.device(self.device):
            rc = self.L.cadl_stats_exchange(self.pkg._ptr(ws.buf), self.ptrs, self.rank, self.world, self.epoch,
                                            self.timeout_s,
                                            self.C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        self.pkg._check(rc, "cadl_stats_exchange")

    def timed_out(self, clear: bool = False) -> bool:
        e = self.C.c_int(0)
        with torch.cuda.device(self.device):
            self.pkg._check(self.L.cadl_p2p_error(self.own, self.world, self.C.byref(e), int(clear)), "cadl_p2p_error")
        return e.value != 0

    def check(self) -> None:
        """Call where the step's results are read on the host (a sync point anyway): a rank that waited longer than
        timeout_s for a peer has NaN statistics for that step; say so instead of training on.  Clears the flag."""
        if self.timed_out(clear=True):
            raise RuntimeError(f"cadl: rank {self.rank} waited more than {self.timeout_s:g} s for a peer's loss statistics "
                               "(cadl_stats_exchange); this step's loss and gradient are NaN")

    def close(self) -> None:
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for r in range(self.world):
                if self.ptrs[r]:
                    self.L.cadl_p2p_close(self.ptrs[r], int(r == self.rank))
                    self.ptrs[r] = None


def si_from_stats(stats: torch.Tensor, lam: float = 0.5) -> float:
    """ScaleInvariantLoss from (n, sum d, sum d^2): depth_loss.h:58-63 (zero when n == 0, :53-55)."""
    n, S, Q = float(stats[ST_SI_N]), float(stats[ST_SI_S]), float(stats[ST_SI_Q])
    return 0.0 if n == 0 else Q / n - lam * S * S / (n * n)


def combine_shares(local: Dict[str, float], weights=(1.0, 0.1, 0.001, 0.01),
                   group: Optional[dist.ProcessGroup] = None, device="cpu") -> Dict[str, float]:
    """In mode "global" the SI loss a rank reports is already global (it is a function of the exchanged statistics),
    while the gradient-matching, smoothness and reprojection losses are this rank's additive share (local sums over
    global denominators: B_global * H * W edges, n_global valid pixels).  Sum the shares and rebuild the weighted total
    (depth_loss.h:427-430)."""
    t = torch.tensor([local["d_grad"], local["d_smooth"], local["d_reproj"]], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    out = dict(local)
    out["d_grad"], out["d_smooth"], out["d_reproj"] = float(t[0]), float(t[1]), float(t[2])
    w = weights
    out["d_total"] = w[0] * out["d_si"] + w[1] * out["d_grad"] + w[2] * out["d_smooth"] + w[3] * out["d_reproj"]
    return out
