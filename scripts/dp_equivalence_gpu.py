"""2-GPU check (run under torchrun, 2 ranks) of the data-parallel train_step on real devices:
 (1) fp32 engine, explicit noise: loss, gate sample and every all-reduced gradient of the sharded step (batch 8 per
     rank) equal those of ONE device on the whole batch of 16 - relative error <= 1e-4 per tensor (1e-5 on the loss);
 (2) bf16 engine, replayed graphs, Philox noise: after three sup+unsup step pairs both ranks hold bit-identical,
     finite parameters (the gate sample and the reduced gradients are the same everywhere), and the step counter
     advanced once per train_step.
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
assert world == 2, "run with --nproc-per-node 2"
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
mu0 = np.load(os.path.join(ROOT, "data", "gating_matrix_0.5.npy"))
cfg = dict(gate_type="learnable", gate_subtype=None, mu_init=mu0, gating_reg=0.2, lr=1e-3, gating_init_temp=1.0,
           batch_size=16, init_temp=0.1)
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

lrn = G.Learner((64, 64, 3), 45, 18, 18, 1000, 0.2, cfg, device=dev, precision="bf16", graphs=True, seed=7)
xs = torch.randint(0, 256, (64, 64, 64, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(rank)).pin_memory()
ys = (torch.rand(64, 18, generator=torch.Generator().manual_seed(10 + rank)) < 0.5).long().pin_memory()
for i in range(3):
    lrn.train_step(xs, ys, True)
    lrn.train_step(xs, None, False)
torch.cuda.synchronize()
mine = lrn.store.flat.clone()
other = mine.clone()
dist.broadcast(other, src=0)
same = bool(torch.equal(mine, other)) and bool(torch.isfinite(mine).all())
its = lrn.optimiser.iterations
if not same or its != 6:
    print("rank %d: graphed DP steps: identical=%s iterations=%d" % (rank, same, its), flush=True)
    ok = False
elif rank == 0:
    print("bf16 graphed DP steps: parameters bit-identical on both ranks, iterations = 6", flush=True)
flag = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(flag)
dist.destroy_process_group()
sys.exit(int(flag.item() != 0))
