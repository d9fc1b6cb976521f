"""GPU: the data-gradient GEMM with the fused residual add (EQFF / HTR backward shape), for ncu source captures.
Exploration tool, not collected by pytest.   python tools/gpu_gemm_add.py [with_add=1]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gotennet_b200 import ops
dev = torch.device("cuda:0")
with_add = (sys.argv[1] != "0") if len(sys.argv) > 1 else True
M, N, K = 147768, 256, 256
torch.manual_seed(0)
g = torch.randn(M, K, device=dev); w = torch.randn(K, N, device=dev) / 16; add = torch.randn(M, N, device=dev)
out = torch.empty(M, N, device=dev)
am, bm = ops.absmax(g, K, M, K), ops.absmax(w, N, K, N)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(6):
    if i == 3:
        e0.record()
    ops.gemm(g, K, 0, w, N, 0, out, N, M, N, K, add_src=add if with_add else None, ld_add=N, impl=3, a_amax=am, b_amax=bm)
e1.record(); torch.cuda.synchronize()
print(f"add={with_add}: {e0.elapsed_time(e1) / 3:.3f} ms per call")
