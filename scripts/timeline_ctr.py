"""Per-tile pipeline timeline of block 0 of the fused conv5t + likelihood kernel (instrumented build, see timeline_probe.py):
GCCVAE_TIMELINE=1 python semi-supervised-gated-lt-vae_b200/build.py && python scripts/timeline_ctr.py"""
import os
import sys
os.environ.setdefault("GCCVAE_LIB", "libgccvae_tl.so")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gccvae_b200._lib as L

lib = L.load()
d = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
tl = torch.zeros(40 * 8 + 2048, dtype=torch.int64, device=d)
B = 1024
g4 = torch.relu(torch.randn(B, 32, 32, 32, device=d)).to(torch.bfloat16)
w8 = torch.randn(16 * 128, device=d).to(torch.bfloat16) * 0.05
b3 = torch.randn(3, device=d)
xb = torch.randint(0, 256, (B, 33, 33, 16), dtype=torch.uint8, device=d)
coef = -torch.rand(B, device=d) / B
lpx = torch.empty(B, device=d)
D2 = torch.empty(B, 33, 33, 16, dtype=torch.bfloat16, device=d)
db3 = torch.zeros(3, device=d)
fn = lambda: L.check(lib.gccvae_convt_recon_bf16(B, g4.data_ptr(), w8.data_ptr(), b3.data_ptr(), xb.data_ptr(), 2, coef.data_ptr(),
                                                 lpx.data_ptr(), D2.data_ptr(), None, db3.data_ptr(), 0, st))
for rep in range(3):
    tl.zero_()
    lib.gccvae_debug_set_timeline(tl.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.gccvae_debug_set_timeline(None)
t = tl.cpu()[:33 * 8].view(33, 8).double()
t0 = t[0, 0]
print("== conv5t fwd + likelihood, block 0 (us since first event; cols: slot-free, tma-issued, tmem-free, landed, acc-ready, acc-read, stored, mma-issued)")
for i in range(20):
    if t[i, 0] == 0:
        break
    print("%2d" % i, " ".join("%7.2f" % ((v - t0) / 1.9e3) if v > 0 else "      -" for v in t[i, :8]))
blk = tl.cpu()[40 * 8:].view(1024, 2)
used = blk[:, 0] > 0
if used.any():
    bs, be = blk[used, 0].double(), blk[used, 1].double()
    g0 = bs.min()
    dur = (be - bs) / 1e3
    print("   %d blocks: start spread %.1f us, last end - first start %.1f us, per-block duration min/median/max %.1f/%.1f/%.1f us"
          % (int(used.sum()), float(bs.max() - g0) / 1e3, float(be.max() - g0) / 1e3, float(dur.min()), float(dur.median()), float(dur.max())))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    fn()
e1.record()
torch.cuda.synchronize()
print("   20 back-to-back launches: %.1f us each (instrumented build)" % (e0.elapsed_time(e1) * 50))
