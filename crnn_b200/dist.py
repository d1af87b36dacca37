"""Trajectory sharding across ranks and the one exchange step of the path.

Trajectories are independent, so each rank owns a contiguous shard and the only collective is
the all-reduce of [sum loss, n, grad_sum] (np + 2 doubles) per optimiser step — NCCL over
NVLink on GPUs, gloo in the CPU tests.  One process per GPU (torchrun), `torch.distributed`
does the plumbing.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(N: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced split of N trajectories: the first N % world ranks get one extra."""
    base, rem = divmod(N, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_loss_grad(loss_sum: float, n: int, grad_sum, device=None):
    """Sum [loss_sum, n, grad_sum...] over all ranks -> (mean loss, mean gradient).

    `grad_sum` may be a numpy array or a torch tensor (CUDA for NCCL).  With no initialised
    process group this is the identity (single-GPU path)."""
    import torch
    import torch.distributed as dist
    if isinstance(grad_sum, np.ndarray):
        g = torch.from_numpy(np.asarray(grad_sum, dtype=np.float64))
    else:
        g = grad_sum.detach().to(torch.float64)
    buf = torch.empty(g.numel() + 2, dtype=torch.float64, device=device if device is not None else g.device)
    buf[0] = float(loss_sum); buf[1] = float(n); buf[2:] = g.to(buf.device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    out = buf.cpu().numpy()
    total = max(out[1], 1.0)
    return out[0] / total, out[2:] / total
