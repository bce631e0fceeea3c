"""Debug probe: per-item pipeline timeline of block 0 of a tap-GEMM launch (clock64 ticks -> ns)."""
import ctypes as C
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gccvae_b200._lib as L
from gccvae_b200._lib import Geom

lib = L.load()
d = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
tl = torch.zeros(32 * 8, dtype=torch.int64, device=d)


def run(name, fn):
    for rep in range(3):
        tl.zero_()
        lib.gccvae_debug_set_timeline(tl.data_ptr())
        fn()
        torch.cuda.synchronize()
        lib.gccvae_debug_set_timeline(None)
    t = tl.cpu().view(32, 8).double()
    t0 = t[0, 0]
    print("==", name, "(us since first event; cols: slot-free, tma-issued, tmem-free, landed, acc-ready, acc-read, stored)")
    for i in range(12):
        if t[i, 0] == 0:
            break
        print(i, " ".join("%7.2f" % ((v - t0) / 1.9e3) if v > 0 else "      -" for v in t[i, :7]))


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
