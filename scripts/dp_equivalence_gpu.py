"""Multi-GPU check (run under torchrun, 2 / 4 / 8 ranks) of the data-parallel train_step on real devices (SURVEY.md 8c
item 9):
 (1) fp32 engine, explicit noise: loss, gate sample and every all-reduced gradient of the sharded step (batch 8 per
     rank) equal those of ONE device on the whole batch of 8 x world - relative error <= 1e-4 per tensor (1e-5 on the
     loss);
 (2) fp32 engine, explicit noise, both exchange paths ("peer": the fused barrier + two-shot all-reduce + Adam kernel over
     NVLink peer memory, "nccl"): after a supervised and an unsupervised train_step the parameters equal those of ONE
     device stepping on the whole batch, and the returned losses agree to 1e-5;
 (3) bf16 engine, replayed graphs, Philox noise, both exchange paths: after three sup+unsup step pairs all ranks hold
     bit-identical, finite parameters (the gate sample and the reduced gradients are the same everywhere), and the step
     counter advanced once per train_step.
Exit code 0 = pass."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

import gccvae_b200 as G
import gccvae_oracle as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
assert world in (2, 4, 8), "run with --nproc-per-node 2, 4 or 8"
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
mu0 = np.load(os.path.join(ROOT, "data", "gating_matrix_0.5.npy"))
cfg = dict(gate_type="learnable", gate_subtype=None, mu_init=mu0, gating_reg=0.2, lr=1e-3, gating_init_temp=1.0,
           batch_size=8 * world, init_temp=0.1)
BL, K = 8, 10
p0 = O.init_params(0, trained_like=True)
x, y, noise = O.make_inputs(world * BL, k=K)
sl = slice(rank * BL, (rank + 1) * BL)
shard_noise = dict(eps=noise["eps"][sl], eps_k=noise["eps_k"][:, sl], U_y=noise["U_y"][sl], U1=noise["U1"], U2=noise["U2"])
ok = True
for supervised in (True, False):
    dp_l = G.Learner((64, 64, 3), 45, 18, 18, 1000, 1.0, cfg, device=dev, precision="fp32")
    dp_l.store.load_dict(p0)
    ref = G.Learner((64, 64, 3), 45, 18, 18, 1000, 1.0, cfg, device=dev, precision="fp32")
    ref._dist, ref.world, ref.rank = None, 1, 0          # one device, whole batch
    ref.store.load_dict(p0)
    loss_dp, c_dp = dp_l.loss_and_grads(x[sl], y[sl], supervised, noise=shard_noise, k=K)   # includes the all-reduce
    loss_ref, c_ref = ref.loss_and_grads(x, y, supervised, noise=noise, k=K)
    torch.cuda.synchronize()
    if not torch.equal(c_dp, c_ref):
        print("rank %d: gate sample differs between the sharded and the whole-batch step" % rank, flush=True)
        ok = False
    if abs(float(loss_dp) - float(loss_ref)) > 1e-5 * abs(float(loss_ref)):
        print("rank %d sup=%s loss %.6f vs %.6f" % (rank, supervised, float(loss_dp), float(loss_ref)), flush=True)
        ok = False
    for name in ref.store.names():
        g_ref, g_dp = ref.store.g(name).double(), dp_l.store.g(name).double()
        scale = float(g_ref.abs().max())
        err = float((g_dp - g_ref).abs().max()) / max(scale, 1e-30)
        if scale > 0 and err > 1e-4:       # fp32 summation order (two partial sums instead of one)
            print("rank %d sup=%s %-18s gradient rel err %.2e" % (rank, supervised, name, err), flush=True)
            ok = False
    if rank == 0:
        print("fp32 sharded step == whole-batch step (supervised=%s)" % supervised, flush=True)

for mode in ("peer", "nccl"):
    # ---- (2) train_step: exchange + Adam against one device on the whole batch --------------------------------------------
    dp_l = G.Learner((64, 64, 3), 45, 18, 18, 1000, 1.0, cfg, device=dev, precision="fp32", dp_exchange=mode)
    dp_l.store.load_dict(p0)
    ref = G.Learner((64, 64, 3), 45, 18, 18, 1000, 1.0, cfg, device=dev, precision="fp32", dp_exchange="nccl")
    ref._dist, ref.world, ref.rank = None, 1, 0
    ref.store.load_dict(p0)
    active = "peer" if dp_l._peer is not None else "nccl"
    for supervised in (True, False):
        l_dp, c_dp = dp_l.train_step(x[sl], y[sl] if supervised else None, supervised, noise=shard_noise, k=K)
        l_ref, c_ref = ref.train_step(x, y if supervised else None, supervised, noise=noise, k=K)
        torch.cuda.synchronize()
        # (mu is learnable: after the first update the two runs' mu - and so their c - agree to rounding, not bit for bit)
        if abs(float(l_dp) - float(l_ref)) > 1e-5 * abs(float(l_ref)) or float((c_dp - c_ref).abs().max()) > 1e-6:
            print("rank %d [%s] train_step sup=%s: loss %.6f vs %.6f" % (rank, active, supervised, float(l_dp), float(l_ref)), flush=True)
            ok = False
    d = (dp_l.store.flat - ref.store.flat).abs()
    # Adam's first steps move every parameter by ~lr = 1e-3 whatever the gradient's size: entries whose gradient is at the
    # level of the fp32 summation noise may differ by O(lr), all others agree to ~1e-8
    frac = float((d > 2e-5).float().mean())
    if frac > 0.01 or dp_l.optimiser.iterations != 2 or float(dp_l.store.grad.abs().max()) != 0.0:
        print("rank %d [%s] train_step: %.4f of the parameters differ, iterations %d, grad max %g" % (
            rank, active, frac, dp_l.optimiser.iterations, float(dp_l.store.grad.abs().max())), flush=True)
        ok = False
    elif rank == 0:
        print("fp32 DP train_step == whole-batch train_step (exchange requested %s, active %s; %.5f of the parameters differ by > 2e-5)" % (
            mode, active, frac), flush=True)
    # ---- (3) replayed bf16 graphs ----------------------------------------------------------------------------------------
    lrn = G.Learner((64, 64, 3), 45, 18, 18, 1000, 0.2, cfg, device=dev, precision="bf16", graphs=True, seed=7, dp_exchange=mode)
    xs = torch.randint(0, 256, (64, 64, 64, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(rank)).pin_memory()
    ys = (torch.rand(64, 18, generator=torch.Generator().manual_seed(10 + rank)) < 0.5).long().pin_memory()
    losses = []
    for i in range(3):
        losses.append(lrn.train_step(xs, ys, True)[0])
        losses.append(lrn.train_step(xs, None, False)[0])
    torch.cuda.synchronize()
    mine = lrn.store.flat.clone()
    other = mine.clone()
    dist.broadcast(other, src=0)
    lv = torch.stack([v.clone() for v in losses])
    lv0 = lv.clone()
    dist.broadcast(lv0, src=0)
    same = bool(torch.equal(mine, other)) and bool(torch.isfinite(mine).all()) and bool(torch.equal(lv, lv0))
    its = lrn.optimiser.iterations
    if not same or its != 6:
        print("rank %d [%s]: graphed DP steps: identical=%s iterations=%d" % (rank, mode, same, its), flush=True)
        ok = False
    elif rank == 0:
        print("bf16 graphed DP steps (exchange %s): parameters and losses bit-identical on all %d ranks, iterations = 6, losses %s" % (
            "peer" if lrn._peer is not None else "nccl", world, [round(float(v), 3) for v in lv]), flush=True)
flag = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(flag)
dist.destroy_process_group()
sys.exit(int(flag.item() != 0))
