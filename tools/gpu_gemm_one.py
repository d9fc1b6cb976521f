"""GPU: one GEMM shape in a loop (for `ncu --set full --import-source on -k regex:gemm16`).  Exploration tool.
    python tools/gpu_gemm_one.py M N K [trans_b] [iters]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gotennet_b200 import ops
dev = torch.device("cuda:0")
M, N, K = (int(x) for x in sys.argv[1:4])
tb = int(sys.argv[4]) if len(sys.argv) > 4 else 1
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 6
torch.manual_seed(0)
a = torch.randn(M, K, device=dev)
w = torch.randn(N, K, device=dev) / K ** 0.5 if tb else torch.randn(K, N, device=dev) / K ** 0.5
bias = torch.randn(N, device=dev)
out = torch.empty(M, N, device=dev)
am, bm = ops.absmax(a, K, M, K), ops.absmax(w, w.shape[1], w.shape[0], w.shape[1])
fn = lambda: ops.gemm(a, K, 0, w, w.shape[1], tb, out, N, M, N, K, bias=bias, impl=3, a_amax=am, b_amax=bm)
for _ in range(2):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    fn()
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / iters
print(f"M={M} N={N} K={K}: {t * 1e3:.1f} us  {2.0 * M * N * K / t / 1e9:.0f} TF/s  {4.0 * (M * K + M * N) / t / 1e6:.0f} GB/s")
