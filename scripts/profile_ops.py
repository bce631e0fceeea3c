"""Per-op device time of one supervised + one unsupervised train_step (eager, CUDA events around every op)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gccvae_b200 as G

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mu = np.load(os.path.join(ROOT, "data", "gating_matrix_0.2.npy"))
cfg = dict(gate_type="fixed", gate_subtype="inferred", mu_init=mu, gating_reg=0.2, lr=1e-4, gating_init_temp=0.3,
           batch_size=B, init_temp=0.1)
eopts = {}
for kv in filter(None, os.environ.get("PROFILE_ENGINE_OPTIONS", "").replace(":", ",").split(",")):
    k_, v_ = kv.split("=")
    eopts[k_] = int(v_) if v_.lstrip("-").isdigit() else v_
lrn = G.Learner((64, 64, 3), 45, 18, 18, 1000, 0.2, cfg, precision="bf16", engine_options=eopts or None)
x = torch.rand(B, 64, 64, 3, device="cuda")
if os.environ.get("PROFILE_U8"):
    x = torch.randint(0, 256, (B, 64, 64, 3), dtype=torch.uint8, device="cuda")
y = (torch.rand(B, 18, device="cuda") < 0.5).long()
for _ in range(3):
    lrn.train_step(x, y, True)
    lrn.train_step(x, None, False)
prof = lrn.engine.profile_step(lrn, x, y, steps=5)
tot = sum(ms * n for ms, n, _ in prof.values())
print("profiled ops total per sup+unsup pair: %.3f ms" % tot)
for k, (ms, n, nb) in sorted(prof.items(), key=lambda kv: -kv[1][0] * kv[1][1]):
    print("%-22s %8.1f us x%.0f  %6.1f%%  %7.1f MB  %7.0f GB/s" % (k, ms * 1e3, n, 100 * ms * n / tot, nb / 1e6, nb / ms / 1e6))
# whole-step timing, eager vs graph
for graphs in (False, True):
    lrn.use_graphs = graphs
    for _ in range(3):
        lrn.train_step(x, y, True); lrn.train_step(x, None, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        lrn.train_step(x, y, True); lrn.train_step(x, None, False)
    e1.record(); torch.cuda.synchronize()
    print("graphs=%s: %.3f ms per pair" % (graphs, e0.elapsed_time(e1) / 20))
