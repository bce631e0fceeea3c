"""Data-parallel host logic of the ELBO step (device agnostic, so it is testable with gloo on CPU).

Design: every per-image coefficient inside the kernels is divided by the GLOBAL batch, so the
only exchange is one all-reduce(sum) of the flat gradient buffer; the gate sample is drawn from a seed shared
by all ranks, per-image noise from a per-rank seed; the L1 term on mu is added once (each rank adds 1/world)."""
from __future__ import annotations

import torch


def dist_or_none():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def world_and_rank(dist):
    return (dist.get_world_size(), dist.get_rank()) if dist is not None else (1, 0)


def batch_global(batch_local: int, world: int) -> int:
    """weak scaling: every rank holds `batch_local` images of the global batch."""
    return batch_local * world


def l1_scale(world: int) -> float:
    return 1.0 / world


def gate_seed(seed: int) -> int:
    """the gate sample c must be identical on every rank (gated_ccvae.py:244 draws ONE c per step)."""
    return int(seed)


def data_seed(seed: int, rank: int) -> int:
    """per-image noise (eps, eps_k, U_y) must differ between ranks."""
    return int(seed) + 7919 * (rank + 1)


def allreduce_sum_(dist, flat: torch.Tensor, n: int):
    """in-place sum over ranks of the first n elements of a flat buffer (the trainable prefix)."""
    if dist is not None and dist.get_world_size() > 1:
        dist.all_reduce(flat[:n], op=dist.ReduceOp.SUM)
    return flat


def global_scalar(dist, t: torch.Tensor) -> torch.Tensor:
    """each rank holds its share of a globally-normalised scalar (the loss); returns the sum."""
    if dist is not None and dist.get_world_size() > 1:
        t = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t
