"""Per-tile pipeline timeline of block (0, 0) of the s2d weight-gradient kernel (instrumented build, see timeline_probe.py):
GCCVAE_TIMELINE=1 python semi-supervised-gated-lt-vae_b200/build.py && python scripts/timeline_wgrad.py
cols: slot-free, tma-issued, mma-thread-ready, landed, (epilogue: acc-ready), -, (epilogue: stored), mma-issued"""
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
for name, HS, CL, CS in (("enc.conv2 / dec.conv4t", 16, 32, 32), ("enc.conv3 / dec.conv3t", 8, 32, 64), ("enc.conv4 / dec.conv2t", 4, 64, 128)):
    in2 = torch.randn(B, HS + 1, HS + 1, 4 * CL, device=d).to(torch.bfloat16)
    S = torch.randn(B, HS, HS, CS, device=d).to(torch.bfloat16)
    dW = torch.zeros(4, 4, CL, CS, device=d)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=d)
    fn = lambda: L.check(lib.gccvae_wg_s2d_bf16(B, HS, HS, CL, in2.data_ptr(), S.data_ptr(), CS, dW.data_ptr(), st))
    for rep in range(3):
        flush.zero_()                      # cold L2, as in the step
        tl.zero_()
        lib.gccvae_debug_set_timeline(tl.data_ptr())
        fn()
        torch.cuda.synchronize()
        lib.gccvae_debug_set_timeline(None)
    t = tl.cpu()[:33 * 8].view(33, 8).double()
    t0 = t[0, 0]
    print("== weight gradient", name, "block (0,0), us since its first event")
    for i in range(32):
        if t[i].abs().sum() == 0:
            continue
        print("%2d" % i, " ".join("%7.2f" % ((v - t0) / 1.9e3) if v > 0 else "      -" for v in t[i, :8]))
    ts = []
    for rep in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print("   one launch, cold L2: %.1f us (min of 5; instrumented build)" % min(ts))
