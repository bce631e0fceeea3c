"""In-graph timeline of one supervised train_step: TIMELINE_MARKERS=1 (marker after every op, default) or 2 (segments).
Marker kernels carry no PDL attribute, so they serialise the stream: compare segment sums, not absolute totals."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gccvae_b200 as G

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
# under torchrun: the data-parallel step (rank 0 prints its own timeline)
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mu = np.load(os.path.join(ROOT, "data", "gating_matrix_0.2.npy"))
cfg = dict(gate_type="fixed", gate_subtype="inferred", mu_init=mu, gating_reg=0.2, lr=1e-4, gating_init_temp=0.3,
           batch_size=B, init_temp=0.1)
lrn = G.Learner((64, 64, 3), 45, 18, 18, 1000, 0.2, cfg, precision="bf16", graphs=True,
                dp_exchange=os.environ.get("TIMELINE_DP_EXCHANGE", "peer"),
                engine_options=dict(markers=int(os.environ.get("TIMELINE_MARKERS", "1"))))
x = torch.randint(0, 256, (B, 64, 64, 3), dtype=torch.uint8, device="cuda") if os.environ.get("TIMELINE_U8", "1") != "0" else torch.rand(B, 64, 64, 3, device="cuda")
y = (torch.rand(B, 18, device="cuda") < 0.5).long()
for sup in (True, False):
    for _ in range(5):
        lrn.train_step(x, y if sup else None, sup)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        lrn.train_step(x, y if sup else None, sup)
    e1.record()
    torch.cuda.synchronize()
    if rank != 0:
        continue
    print("supervised=%s: %.1f us per train_step (with markers)" % (sup, e0.elapsed_time(e1) * 50))
    marks = lrn.engine.marks
    t = lrn.engine.mark_buf.cpu()[:len(marks)].double() / 1e3
    t0 = t[0]
    last = {"main": t0, "side": None, "side2": None, "side3": None, "wg1": None, "wg2": None}
    for (what, lane), ti in zip(marks, t):
        prev = last[lane]
        d = (ti - prev) if prev is not None else float("nan")
        print("%-6s %9.1f  +%7.1f  %s" % (lane, ti - t0, d, what))
        last[lane] = ti
if world > 1:
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
