"""GPU: timing of the big edge GEMM forms only (exploration tool)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gotennet_b200 import ops
dev = torch.device("cuda:0")
def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
M, N, K = 301491, 1792, 256
a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5
g = torch.randn(M, N, device=dev)
y = torch.empty(M, N, device=dev); da = torch.empty(M, K, device=dev); dw = torch.empty(N, K, device=dev)
fl = 2.0 * M * N * K
t = [timeit(f) for f in (lambda: ops.gemm(a, K, 0, w, K, 1, y, N, M, N, K), lambda: ops.gemm(g, N, 0, w, K, 0, da, K, M, K, N),
                         lambda: ops.gemm(g, N, 1, a, K, 0, dw, K, N, K, M))]
print(os.environ.get("GOTEN_GEMM_NCTA", "-"), os.environ.get("GOTEN_GEMM_DBG", "-"),
      "  ".join(f"{nm} {x:.3f}ms ({fl / (x * 1e-3) / 1e12:.0f} TF/s)" for nm, x in zip(("fwd", "dgrad", "wgrad"), t)), flush=True)
