"""Data-parallel host logic of the ELBO step (device agnostic, so it is testable with gloo on CPU).

Design: every per-image coefficient inside the kernels is divided by the GLOBAL batch, so the
only exchange is one all-reduce(sum) of the flat gradient buffer; the gate sample is drawn from a seed shared
by all ranks, per-image noise from a per-rank seed; the L1 term on mu is added once (each rank adds 1/world).

Two exchange paths: `PeerExchange` - the gradient buffers of all ranks of the node live in symmetric memory and ONE
kernel per call (csrc/dp.cu) does rank barrier + two-shot all-reduce over NVLink + Adam, inside the step's CUDA
graph, the bulk of it under the last dgrad; or, where peer memory cannot be set up, `allreduce_sum_` - NCCL's
all-reduce between the replayed graph and the optimiser."""
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


def shard_bounds(n4: int, world: int, rank: int):
    """the slice [lo, hi) of a range of n4 16-byte units that rank `rank` sums and redistributes in the two-shot
    all-reduce (the partition csrc/dp.cu uses: ceil(n4 / world) units per rank, the last shards may be short or empty)."""
    per = (n4 + world - 1) // world
    lo = min(n4, rank * per)
    return lo, min(n4, lo + per)


def two_shot_allreduce_reference(bufs):
    """what the kernel computes, restated on host arrays: every rank sums its shard over all buffers in rank order and
    writes the sum into every buffer.  bufs: list (one per rank) of equal-length 1-D float32 tensors whose length is a
    multiple of 4; they are modified in place."""
    world, n4 = len(bufs), bufs[0].numel() // 4
    for rank in range(world):
        lo, hi = shard_bounds(n4, world, rank)
        if hi <= lo:
            continue
        acc = bufs[0][4 * lo:4 * hi].clone()
        for q in range(1, world):
            acc += bufs[q][4 * lo:4 * hi]
        for q in range(world):
            bufs[q][4 * lo:4 * hi] = acc
    return bufs


class PeerExchange:
    """Gradient exchange over NVLink peer memory (one node): allocates this rank's flat gradient buffer and the ranks'
    barrier words in symmetric memory (torch.distributed._symmetric_memory), so that the kernel of csrc/dp.cu can read
    and write every rank's buffer directly.  Construction is collective; it raises if peer memory cannot be set up
    (the caller then stays on the NCCL path)."""

    def __init__(self, dist, n_floats: int, device, push_floats: int = 0):
        import torch.distributed._symmetric_memory as symm
        from ._lib import DP_MAX_RANKS, DP_SYNC_WORDS
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if self.world > DP_MAX_RANKS:
            raise RuntimeError("PeerExchange supports up to {} ranks".format(DP_MAX_RANKS))
        group = dist.group.WORLD
        try:
            symm.enable_symm_mem_for_group(group.group_name)
        except Exception:       # newer torch: rendezvous sets the group up itself
            pass
        self.grad = symm.empty(n_floats, dtype=torch.float32, device=device)
        self.sync = symm.empty(DP_SYNC_WORDS, dtype=torch.int32, device=device)
        self.grad.zero_()
        self.sync.zero_()
        self._h_grad = symm.rendezvous(self.grad, group)
        self._h_sync = symm.rendezvous(self.sync, group)
        # receive slots of the one-barrier exchange of the step's last gradients: one slot per rank
        self.recv_stride = (int(push_floats) + 4 + 3) // 4 * 4 if push_floats else 0
        self.recv, self.recv_ptrs = None, []
        if self.recv_stride:
            self.recv = symm.empty(self.world * self.recv_stride, dtype=torch.float32, device=device)
            self.recv.zero_()
            self._h_recv = symm.rendezvous(self.recv, group)
            self.recv_ptrs = [int(p) for p in self._h_recv.buffer_ptrs]
        self.grad_ptrs = [int(p) for p in self._h_grad.buffer_ptrs]
        self.sync_ptrs = [int(p) for p in self._h_sync.buffer_ptrs]
        if len(self.grad_ptrs) != self.world or self.grad_ptrs[self.rank] != self.grad.data_ptr():
            raise RuntimeError("symmetric memory rendezvous returned inconsistent pointers")
        torch.cuda.synchronize(device)
        dist.barrier()          # every rank's barrier words are zero before the first kernel signals into them

    def fill_args(self, a):
        """rank / world / peer pointer arrays of a `_lib.DpArgs`."""
        a.world, a.rank = self.world, self.rank
        for q in range(self.world):
            a.grad[q] = self.grad_ptrs[q]
            a.sync[q] = self.sync_ptrs[q]
            if self.recv_ptrs:
                a.recv[q] = self.recv_ptrs[q]
        a.recv_stride = self.recv_stride
        return a
