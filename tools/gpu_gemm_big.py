"""GPU: one pass of the three GEMM forms at the cfg2 edge shape, for `ncu --metrics gpu__time_duration.sum`
(exploration tool, not collected by pytest).   python tools/gpu_gemm_big.py [impl] [M N K]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gotennet_b200 import ops
dev = torch.device("cuda:0")
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 3
M, N, K = (int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (301491, 1792, 256)
torch.manual_seed(0)
a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5
b = torch.randn(N, device=dev); g = torch.randn(M, N, device=dev)
y = torch.empty(M, N, device=dev); da = torch.empty(M, K, device=dev); dw = torch.empty(N, K, device=dev); db = torch.empty(N, device=dev)
for _ in range(3):
    ops.gemm(a, K, 0, w, K, 1, y, N, M, N, K, bias=b, impl=impl)
    ops.gemm(g, N, 0, w, K, 0, da, K, M, K, N, impl=impl)
    ops.gemm(g, N, 1, a, K, 0, dw, K, N, K, M, colsum=db, impl=impl)
torch.cuda.synchronize()
