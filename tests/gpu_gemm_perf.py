"""GPU: timing of goten_gemm variants (exploration)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gotennet_b200 import ops
dev = torch.device("cuda:0")

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for (M, N, K) in [(301491, 1792, 256), (301491, 256, 256), (301491, 256, 1792), (37686, 1792, 2048)]:
    a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev); add = torch.randn(M, N, device=dev)
    y = torch.empty(M, N, device=dev); act = torch.empty(M, N, device=dev)
    fl = 2.0 * M * N * K
    for impl in (2, 1):
        variants = {
            "plain": lambda: ops.gemm(a, K, 0, w, K, 1, y, N, M, N, K, impl=impl),
            "bias": lambda: ops.gemm(a, K, 0, w, K, 1, y, N, M, N, K, bias=b, impl=impl),
            "bias+add": lambda: ops.gemm(a, K, 0, w, K, 1, y, N, M, N, K, bias=b, add_src=add, ld_add=N, impl=impl),
            "bias+act": lambda: ops.gemm(a, K, 0, w, K, 1, y, N, M, N, K, bias=b, act_out=act, ld_act=N, act_lo=0, act_hi=N, impl=impl),
        }
        print(f"M={M} N={N} K={K} impl={impl}: " + "  ".join(f"{k} {timeit(f):.2f}ms ({fl / (timeit(f) * 1e-3) / 1e12:.0f} TF/s)" for k, f in variants.items()), flush=True)
