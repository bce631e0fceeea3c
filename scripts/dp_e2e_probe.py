"""Diagnostic (run under torchrun): host-issue time vs device time of train_step with resident and with host inputs."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import gccvae_b200 as G

rank, world, lr_ = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr_)
dev = torch.device("cuda", lr_)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mu = np.load(os.path.join(ROOT, "data", "gating_matrix_0.2.npy"))
cfg = dict(gate_type="fixed", gate_subtype="inferred", mu_init=mu, gating_reg=0.2, lr=1e-4, gating_init_temp=0.3,
           batch_size=1024, init_temp=0.1)
B = int(os.environ.get("PROBE_B", "1024"))
lrn = G.Learner((64, 64, 3), 45, 18, 18, 1000, 0.2, cfg, device=dev, precision="bf16", graphs=True)
hx = [torch.randint(0, 256, (B, 64, 64, 3), dtype=torch.uint8).pin_memory() for _ in range(4)]
hy = [(torch.rand(B, 18) < 0.5).long().pin_memory() for _ in range(4)]
dx, dy = [t.to(dev) for t in hx], [t.to(dev) for t in hy]


def run(xs, ys, steps=200, label=""):
    for i in range(6):
        lrn.train_step(xs[i % 4], ys[i % 4], True); lrn.train_step(xs[(i + 1) % 4], None, False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(steps):
        lrn.train_step(xs[i % 4], ys[i % 4], True); lrn.train_step(xs[(i + 1) % 4], None, False)
    e1.record(); t_issue = time.perf_counter() - t0
    torch.cuda.synchronize(); t_all = time.perf_counter() - t0
    print("rank %d %-10s host issue %.3f ms/pair, wall %.3f ms/pair, device %.3f ms/pair" % (
        rank, label, 1e3 * t_issue / steps, 1e3 * t_all / steps, e0.elapsed_time(e1) / steps), flush=True)


run(dx, dy, label="resident")
run(hx, hy, label="host-u8")
run(dx, dy, label="resident")
if world > 1:
    dist.destroy_process_group()
