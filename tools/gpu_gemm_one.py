"""GPU: a few launches of ONE big GEMM form for ncu (exploration tool).  argv: fwd|dgrad|wgrad"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gotennet_b200 import ops
dev = torch.device("cuda:0")
form = sys.argv[1] if len(sys.argv) > 1 else "fwd"
M, N, K = 301491, 1792, 256
a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5
g = torch.randn(M, N, device=dev)
y = torch.empty(M, N, device=dev); da = torch.empty(M, K, device=dev); dw = torch.empty(N, K, device=dev)
for _ in range(3):
    if form == "fwd": ops.gemm(a, K, 0, w, K, 1, y, N, M, N, K)
    elif form == "dgrad": ops.gemm(g, N, 0, w, K, 0, da, K, M, K, N)
    else: ops.gemm(g, N, 1, a, K, 0, dw, K, N, K, M)
torch.cuda.synchronize()
