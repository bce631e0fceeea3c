"""World-size-2 and world-size-8 gloo tests (CPU) of the data-parallel design: shard the batch, give every rank the SAME gate
noise, normalise by the global batch, add the L1 term once, all-reduce(sum) the flat gradient buffer -> the result
equals the single-process step on the whole batch.  The arithmetic here is the
oracle's; the host logic under test is the product's dp.py (the same functions Learner calls on the GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gccvae_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B_LOCAL = {2: 3, 8: 1}       # images per rank


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cfg():
    mu = np.load(os.path.join(ROOT, "tests", "golden", "data", "gating_matrix_0.5.npy"))
    return dict(gate_type="learnable", gate_subtype=None, mu_init=mu, gating_reg=0.2)


def _flat(p_names, grads):
    return torch.cat([grads[k].reshape(-1) for k in p_names] + [grads["mu"].reshape(-1)])


def _worker(rank, port, supervised, out_dir, WORLD):
    B_LOCAL_ = B_LOCAL[WORLD]
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gccvae_b200.dp as dp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        torch.set_num_threads(2 if WORLD == 2 else 1)
        d = dp.dist_or_none()
        world, r = dp.world_and_rank(d)
        assert (world, r) == (WORLD, rank)
        assert dp.gate_seed(1234) == 1234 and dp.data_seed(1234, 0) != dp.data_seed(1234, 1)
        cfg = _cfg()
        p = O.init_params(0, dtype=torch.float64, trained_like=True)
        mu, _ = O.initialise_mu(cfg, dtype=torch.float64)
        x, y, noise = O.make_inputs(B_LOCAL_ * WORLD, k=6, dtype=torch.float64)      # the GLOBAL batch, same on all ranks
        sl = slice(rank * B_LOCAL_, (rank + 1) * B_LOCAL_)
        shard_noise = dict(eps=noise["eps"][sl], eps_k=noise["eps_k"][:, sl], U_y=noise["U_y"][sl],
                           U1=noise["U1"], U2=noise["U2"])                            # shared gate noise
        # local step with the kernels' normalisation: mean over the GLOBAL batch, L1 scaled by 1/world
        leaf = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        mu_leaf = mu.clone().requires_grad_(True)
        fn = O.sup_loss if supervised else O.unsup_loss
        args = (leaf, mu_leaf, x[sl], y[sl], shard_noise, dict(cfg, gate_type="fixed", gate_subtype="inferred"), 0.7) \
            if supervised else (leaf, mu_leaf, x[sl], shard_noise, dict(cfg, gate_type="fixed", gate_subtype="inferred"), 0.7)
        out = fn(*args)                                                               # no L1 inside (gate_type fixed)
        local = out["loss"] * B_LOCAL_ / dp.batch_global(B_LOCAL_, world)
        local = local + dp.l1_scale(world) * cfg["gating_reg"] * mu_leaf.abs().mean()
        local.backward()
        names = list(p.keys())
        flat = _flat(names, {**{k: leaf[k].grad for k in names}, "mu": mu_leaf.grad})
        pad = torch.cat([flat, torch.full((5,), 123.0, dtype=flat.dtype)])             # trailing non-trainable tail
        dp.allreduce_sum_(d, pad, flat.numel())
        assert torch.all(pad[flat.numel():] == 123.0)                                  # only the prefix is reduced
        loss = dp.global_scalar(d, local.detach())
        # every rank sampled the same c
        cs = [torch.zeros_like(out["c"]) for _ in range(world)]
        dist.all_gather(cs, out["c"].detach())
        assert all(torch.equal(cs[0], c) for c in cs[1:])
        if rank == 0:
            torch.save(dict(flat=pad[:flat.numel()], loss=loss), os.path.join(out_dir, "dp.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("supervised,WORLD", [(True, 2), (False, 2), (True, 8), (False, 8)])
def test_sharded_dp_equals_single_process(tmp_path, supervised, WORLD):
    port = _free_port()
    mp.spawn(_worker, args=(port, supervised, str(tmp_path), WORLD), nprocs=WORLD, join=True)
    got = torch.load(os.path.join(str(tmp_path), "dp.pt"))
    cfg = _cfg()
    p = O.init_params(0, dtype=torch.float64, trained_like=True)
    mu, _ = O.initialise_mu(cfg, dtype=torch.float64)
    x, y, noise = O.make_inputs(B_LOCAL[WORLD] * WORLD, k=6, dtype=torch.float64)
    out, g = O.loss_and_grads(p, mu, x, y, noise, cfg, 0.7, supervised)
    want = _flat(list(p.keys()), g)
    assert float((got["flat"] - want).abs().max() / want.abs().max()) < 1e-10
    assert abs(float(got["loss"]) - float(out["loss"])) < 1e-9 * abs(float(out["loss"]))


@pytest.mark.parametrize("world", [2, 3, 4, 8, 16])
def test_two_shot_partition_covers_the_range_once(world):
    """the shard partition of the fused exchange kernel (csrc/dp.cu restated in dp.shard_bounds): disjoint, ordered, complete -
    for ranges shorter than the world too - and the two-shot schedule leaves the plain sum in every rank's buffer."""
    import sys
    sys.path.insert(0, ROOT)
    import gccvae_b200.dp as dp
    for n4 in (0, 1, 3, world - 1, world, world + 1, 392, 251979, 251898):
        got = []
        for r in range(world):
            lo, hi = dp.shard_bounds(n4, world, r)
            assert 0 <= lo <= hi <= n4
            got += list(range(lo, hi)) if n4 < 1000 else []
            if r:
                assert lo == dp.shard_bounds(n4, world, r - 1)[1] or lo == n4
        if n4 < 1000:
            assert got == list(range(n4))
        assert dp.shard_bounds(n4, world, world - 1)[1] == n4
    g = torch.Generator().manual_seed(world)
    bufs = [torch.randn(4 * 37, generator=g) for _ in range(world)]
    want = bufs[0].clone()
    for b in bufs[1:]:
        want = want + b               # rank order, as the kernel sums
    dp.two_shot_allreduce_reference(bufs)
    for b in bufs:
        assert torch.equal(b, want)
