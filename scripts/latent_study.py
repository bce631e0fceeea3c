"""BASELINE.json config 4: the latent stage alone at a large batch, fixed-noise mode (HBM roofline) and Philox mode
(issue / SFU bound): the stand-alone latent kernels (fp32 engine) and the fused dense-chain kernels of the bf16 engine
(heads + latent + fc1 + conv1t in one launch, csrc/chain.cu).  Prints one JSON line per kernel and mode with the
achieved GB/s against MEASURED_PEAKS.json."""
import ctypes as C
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gccvae_b200 as G
from gccvae_b200._lib import ptr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K = 100
mu = np.load(os.path.join(ROOT, "data", "gating_matrix_1.0.npy"))
cfg = dict(gate_type="learnable", gate_subtype=None, mu_init=mu, gating_reg=0.2, lr=1e-4, gating_init_temp=1.0,
           batch_size=B, init_temp=0.1)
lrn = G.Learner((64, 64, 3), 45, 18, 18, 1000, 1.0, cfg, precision="fp32")
dev = lrn.device
b, lb = lrn.engine.bufs(B), lrn._latent_bufs(B)
b["enc.locs.out"].normal_(); b["enc.std.out"].normal_(); b["dz"].normal_(); lb["log_pxz"].fill_(-11000.0)
y = (torch.rand(B, 18, device=dev) < 0.5).long()
peak = 6553.6
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
for mode in ("fixed-noise", "philox"):
    if mode == "fixed-noise":
        noise = dict(eps=torch.randn(B, 45, device=dev), eps_k=torch.randn(K, B, 18, device=dev),
                     U1=torch.rand(18, 18, device=dev), U2=torch.rand(18, 18, device=dev))
    else:
        noise = None
    n = lrn._noise(noise, B, True, K)
    lrn._gate(n)

    def fwd():
        lrn._latent_fwd(B, lb, b, y, n, True, K)

    def bwd():
        lrn._latent_bwd(B, lb, b, n, True, K)

    for name, fn in (("latent_fwd<sup>", fwd), ("latent_bwd<sup>", bwd)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        # algorithmic bytes per image: heads' pre-activations 2x45x4, y, outputs (fwd: loc,scale,z 3x45x4 + 6 terms +
        # logits 18x4; bwd: dz 45x4 in, 2x45x4 out) + fixed noise eps 45x4 and eps_k Kx18x4
        per_img = 2 * 45 * 4 + 18 * 8 + (3 * 45 * 4 + 6 * 4 + 18 * 4 if "fwd" in name else 3 * 45 * 4 + 7 * 4)
        if mode == "fixed-noise":
            per_img += 45 * 4 + K * 18 * 4
        gbs = per_img * B / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": name, "mode": mode, "batch": B, "K": K, "ms": round(ms, 4),
                          "bytes_per_image": per_img, "GB/s": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peak, 4),
                          "softplus_per_s": round(B * K * 18 / (ms * 1e-3) / 1e9, 2)}))


# ---- the fused dense-chain kernels of the bf16 engine (what the bench path runs) ---------------------------------------
lrn2 = G.Learner((64, 64, 3), 45, 18, 18, 1000, 1.0, cfg, precision="bf16")
b2, lb2 = lrn2.engine.bufs(B), lrn2._latent_bufs(B)
lrn2.engine.pack_weights()
b2["enc.conv5.out"].copy_(torch.rand(B, 1, 1, 256, device=dev).to(torch.bfloat16))
b2["dec.conv1t.dout"].copy_((0.01 * torch.randn(B, 4, 4, 128, device=dev)).to(torch.bfloat16))
lb2["log_pxz"].fill_(-11000.0)
for mode in ("fixed-noise", "philox"):
    noise = None
    if mode == "fixed-noise":
        noise = dict(eps=torch.randn(B, 45, device=dev), eps_k=torch.randn(K, B, 18, device=dev),
                     U1=torch.rand(18, 18, device=dev), U2=torch.rand(18, 18, device=dev))
    n = lrn2._noise(noise, B, True, K)
    lrn2._gate(n)
    for name, fn in (("chain_fwd<sup>", lambda: lrn2._chain_fwd(B, lb2, b2, y, n, True, K)),
                     ("chain_bwd<sup>", lambda: lrn2._chain_bwd(B, lb2, b2, n, True, K))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        # per image: fwd reads h5 512 + y 144, writes pre 384 + loc / scale / z 540 + terms 24 + logits 72 + y_out 72 + z16 128
        # + g0 128 + g1 4096; bwd reads dg1 4096 + g0 128 + h5 512 + pre 384 + y 72 + terms 28, writes dg0 128 + dpre16 192
        # + dh5 512; + the fixed noise (eps 180, eps_k K x 72); the weights (~0.6 MB per CTA, L2-resident) are not counted
        per_img = (512 + 144 + 384 + 540 + 24 + 72 + 72 + 128 + 128 + 4096) if "fwd" in name else \
                  (4096 + 128 + 512 + 384 + 72 + 28 + 128 + 192 + 512)
        if mode == "fixed-noise":
            per_img += 45 * 4 + K * 18 * 4
        gbs = per_img * B / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": name, "mode": mode, "batch": B, "K": K, "ms": round(ms, 4),
                          "bytes_per_image": per_img, "GB/s": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peak, 4),
                          "softplus_per_s": round(B * K * 18 / (ms * 1e-3) / 1e9, 2)}))
