"""Debug probe: per-item pipeline timeline of block 0 of a tap-GEMM launch (clock64 ticks -> ns).
Needs the instrumented build: `GCCVAE_TIMELINE=1 python semi-supervised-gated-lt-vae_b200/build.py` -> csrc/libgccvae_tl.so
(the hooks are compiled out of the production library: they cost ~25 % of the c3conv epilogue even when idle)."""
import ctypes as C
import sys
import os
os.environ.setdefault("GCCVAE_LIB", "libgccvae_tl.so")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gccvae_b200._lib as L
from gccvae_b200._lib import Geom

lib = L.load()
d = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
tl = torch.zeros(40 * 8 + 2048, dtype=torch.int64, device=d)


def run(name, fn):
    for rep in range(3):
        tl.zero_()
        lib.gccvae_debug_set_timeline(tl.data_ptr())
        fn()
        torch.cuda.synchronize()
        lib.gccvae_debug_set_timeline(None)
    t = tl.cpu()[:33 * 8].view(33, 8).double()
    t0 = t[0, 0]
    print("==", name, "(us since first event; cols: slot-free, tma-issued, tmem-free, landed, acc-ready, acc-read, stored)")
    for i in range(12):
        if t[i, 0] == 0:
            break
        print(i, " ".join("%7.2f" % ((v - t0) / 1.9e3) if v > 0 else "      -" for v in t[i, :8]))
    a = t[32]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("   20 back-to-back launches: %.1f us each" % (e0.elapsed_time(e1) * 50))
    blk = tl.cpu()[40 * 8:].view(1024, 2)
    used = blk[:, 0] > 0
    if used.any():
        bs, be = blk[used, 0].double(), blk[used, 1].double()
        g0 = bs.min()
        dur = (be - bs) / 1e3
        print("   %d blocks: start spread %.1f us, end-first-start max %.1f us, per-block duration min/median/max %.1f/%.1f/%.1f us"
              % (int(used.sum()), float(bs.max() - g0) / 1e3, float(be.max() - g0) / 1e3, float(dur.min()), float(dur.median()), float(dur.max())))
    if a[4] > 0:
        print("   producer: waited %.0f of %.0f clk | mma thread: waited for data %.0f, for accumulator %.0f of %.0f clk; "
              "%d k-iters -> %.0f clk/k-iter; stages %d, grid %d" % (a[0], a[1], a[2], a[3], a[4], a[5], a[4] / max(a[5], 1), a[6], a[7]))


B = 1024
X64 = torch.randn(B * 1024, 64, device=d).to(torch.bfloat16)
wp = torch.randn(32 * 64, device=d).to(torch.bfloat16)
bias = torch.randn(32, device=d)
h1 = torch.empty(B, 32, 32, 32, dtype=torch.bfloat16, device=d)
gd = Geom(B * 1024, 1, 1, 64, 1, 1, 32, 1, 1, 1, 0)
run("conv1 fwd (dense K=64,N=32)", lambda: L.check(lib.gccvae_ls_bf16(C.byref(gd), X64.data_ptr(), wp.data_ptr(), bias.data_ptr(), 1, None, h1.data_ptr(), 0, st)))
# conv2 fwd
g2 = Geom(B, 32, 32, 32, 16, 16, 32, 4, 4, 2, 1)
w2 = torch.randn(32 * 512, device=d).to(torch.bfloat16)
h2 = torch.empty(B, 16, 16, 32, dtype=torch.bfloat16, device=d)
run("conv2 fwd (16 taps)", lambda: L.check(lib.gccvae_ls_bf16(C.byref(g2), h1.data_ptr(), w2.data_ptr(), bias.data_ptr(), 1, None, h2.data_ptr(), 0, st)))
# conv5t fwd
g5 = Geom(B, 64, 64, 3, 32, 32, 32, 4, 4, 2, 1)
w5 = torch.randn(4 * 16 * 128, device=d).to(torch.bfloat16)
xh = torch.empty(B, 64, 64, 4, device=d)
b3 = torch.randn(3, device=d)
run("conv5t fwd (4 phases x 4 taps, N=16)", lambda: L.check(lib.gccvae_sl_bf16(C.byref(g5), h1.data_ptr(), w5.data_ptr(), b3.data_ptr(), 2, None, xh.data_ptr(), 2, st)))


def pack_sl9(CL, CS):
    g = Geom(1, 2, 2, CL, 1, 1, CS, 4, 4, 2, 1)
    n = lib.gccvae_packed_weight_elems(C.byref(g), 2)
    return torch.randn(n, device=d).to(torch.bfloat16) * 0.05


def run_halo(name, geom, S, W, bias, act, mask, out, out_f32):
    run(name, lambda: L.check(lib.gccvae_sl_halo_bf16(C.byref(geom), S.data_ptr(), W.data_ptr(),
                                                      None if bias is None else bias.data_ptr(), act,
                                                      None if mask is None else mask.data_ptr(), out.data_ptr(),
                                                      out_f32, st)))
    print("   (halo cols: slot-free, tma-issued, tmem-free, landed, acc-ready, acc-read, stored, mma-issued)")
    t = tl.cpu()[:33 * 8].view(33, 8).double()
    print("   mma-issued:", " ".join("%7.2f" % ((t[i, 7] - t[0, 0]) / 1.9e3) for i in range(12)))


g4 = torch.randn(B, 32, 32, 32, device=d).to(torch.bfloat16)
run_halo("HALO conv5t fwd", Geom(B, 64, 64, 3, 32, 32, 32, 4, 4, 2, 1), g4, pack_sl9(3, 32), b3, 2, None, xh, 2)
g3 = torch.randn(B, 16, 16, 32, device=d).to(torch.bfloat16)
b32 = torch.randn(32, device=d)
o4 = torch.empty(B, 32, 32, 32, dtype=torch.bfloat16, device=d)
run_halo("HALO conv4t fwd", Geom(B, 32, 32, 32, 16, 16, 32, 4, 4, 2, 1), g3, pack_sl9(32, 32), b32, 1, None, o4, 0)
run_halo("HALO conv2 dgrad (mask)", Geom(B, 32, 32, 32, 16, 16, 32, 4, 4, 2, 1), g3, pack_sl9(32, 32), None, 0, g4, o4, 0)

# more tap-GEMM shapes
g3c = Geom(B, 16, 16, 32, 8, 8, 64, 4, 4, 2, 1)
w3 = torch.randn(64 * 512, device=d).to(torch.bfloat16)
b64 = torch.randn(64, device=d)
h3 = torch.empty(B, 8, 8, 64, dtype=torch.bfloat16, device=d)
run("conv3 fwd (16 taps, N=64)", lambda: L.check(lib.gccvae_ls_bf16(C.byref(g3c), g3.data_ptr(), w3.data_ptr(), b64.data_ptr(), 1, None, h3.data_ptr(), 0, st)))
g5c = Geom(B, 4, 4, 128, 1, 1, 256, 4, 4, 1, 0)
h4 = torch.randn(B, 4, 4, 128, device=d).to(torch.bfloat16)
w5c = torch.randn(256 * 2048, device=d).to(torch.bfloat16)
b256 = torch.randn(256, device=d)
h5 = torch.empty(B, 256, dtype=torch.bfloat16, device=d)
run("conv5 fwd (dense K=2048, N=64 slabs)", lambda: L.check(lib.gccvae_ls_bf16(C.byref(g5c), h4.data_ptr(), w5c.data_ptr(), b256.data_ptr(), 1, None, h5.data_ptr(), 0, st)))

# x2 end layers
X2 = torch.randn(B, 33, 33, 16, device=d).to(torch.bfloat16)
wx2 = torch.randn(32 * 64, device=d).to(torch.bfloat16)
run("C3CONV conv1 fwd (cols: slot-free, cp.async issued, tmem-free, landed, acc-ready, acc-read, stored, mma-issued)", lambda: L.check(lib.gccvae_c3conv_bf16(B, X2.data_ptr(), wx2.data_ptr(), 32, bias.data_ptr(), 1, None, h1.data_ptr(), st)))
run("C3CONV conv5t dgrad (mask)", lambda: L.check(lib.gccvae_c3conv_bf16(B, X2.data_ptr(), wx2.data_ptr(), 32, None, 0, g4.data_ptr(), h1.data_ptr(), st)))

w8 = torch.randn(16 * 128, device=d).to(torch.bfloat16) * 0.05
xf = torch.rand(B, 64, 64, 3, device=d)
coef = -torch.rand(B, device=d) / B
lpx = torch.empty(B, device=d)
D2 = torch.empty(B, 33, 33, 16, dtype=torch.bfloat16, device=d)
db3 = torch.zeros(3, device=d)
run("CTR conv5t fwd + recon fused (cols: slot-free, tma-issued, tmem-free, landed, acc-ready, acc-read, stored, mma-issued)",
    lambda: L.check(lib.gccvae_convt_recon_bf16(B, g4.data_ptr(), w8.data_ptr(), b3.data_ptr(), xf.data_ptr(), 0, coef.data_ptr(), lpx.data_ptr(), D2.data_ptr(), None, db3.data_ptr(), 0, st)))
