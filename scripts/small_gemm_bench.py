"""Back-to-back timing of the small dense tensor-core GEMMs (launch-latency floor of the tap-GEMM kernel)."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gccvae_b200._lib as L

lib = L.load()
d = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
B = 1024


def bench(name, rows, K, N, out_f32=0, reps=200):
    A = torch.randn(rows, K, device=d).to(torch.bfloat16)
    W = torch.randn(N, K, device=d).to(torch.bfloat16)
    bias = torch.randn(N, device=d)
    out = torch.empty(rows, N, dtype=torch.float32 if out_f32 else torch.bfloat16, device=d)
    fn = lambda: L.check(lib.gccvae_gemm_bf16(rows, K, N, A.data_ptr(), W.data_ptr(), bias.data_ptr(), N, 0, 1, None,
                                              out.data_ptr(), out_f32, torch.cuda.current_stream().cuda_stream))
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("%-28s rows=%5d K=%5d N=%5d  %.2f us per launch (graph of %d)" % (name, rows, K, N, e0.elapsed_time(e1) * 1e3 / reps, reps))


bench("fc1 fwd", B, 64, 64)
bench("heads fwd", B, 256, 96, out_f32=1)
bench("conv1t fwd", B, 64, 2048)
bench("conv1t dgrad", B, 2048, 64)
bench("heads dgrad", B, 96, 256)
bench("conv5 fwd-like", B, 2048, 256)
bench("one tile", 128, 64, 64)
