"""GPU: accuracy (vs fp64) and timing of goten_gemm at the shapes of the cfg2 step (exploration tool,
not collected by pytest).   python tools/gpu_gemm_perf.py [quick]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gotennet_b200 import ops
dev = torch.device("cuda:0")
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def rel(a, b):
    return ((a.double() - b).abs().max() / b.abs().max()).item()


E, Nn = 301491, 18471
shapes = [(E, 1792, 256), (E, 256, 1792), (Nn * 8, 256, 256), (Nn, 1024, 256), (Nn, 1280, 256), (1000, 130, 37)]
if quick:
    shapes = [(4096, 1792, 256), (1000, 130, 37), (777, 256, 256)]
torch.manual_seed(0)
for (M, N, K) in shapes:
    a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    g = torch.randn(M, N, device=dev)
    y = torch.empty(M, N, device=dev); da = torch.empty(M, K, device=dev); dw = torch.empty(N, K, device=dev); db = torch.empty(N, device=dev)
    fl = 2.0 * M * N * K
    chk = M * N * K < 3e11
    for impl in (0,):
        f_fwd = lambda: ops.gemm(a, K, 0, w, K, 1, y, N, M, N, K, bias=b, impl=impl)
        f_dgr = lambda: ops.gemm(g, N, 0, w, K, 0, da, K, M, K, N, impl=impl)
        f_wgr = lambda: ops.gemm(g, N, 1, a, K, 0, dw, K, N, K, M, colsum=db, impl=impl)
        t = [timeit(f) for f in (f_fwd, f_dgr, f_wgr)]
        msg = f"M={M} N={N} K={K} impl={impl}: " + "  ".join(f"{nm} {x:.3f}ms ({fl / (x * 1e-3) / 1e12:.0f} TF/s)" for nm, x in zip(("fwd", "dgrad", "wgrad"), t))
        if chk:
            ad, wd, gd = a.double(), w.double(), g.double()
            msg += f"  err fwd {rel(y, ad @ wd.T + b.double()):.1e} dgrad {rel(da, gd @ wd):.1e} wgrad {rel(dw, gd.T @ ad):.1e} db {rel(db, gd.sum(0)):.1e}"
        print(msg, flush=True)
