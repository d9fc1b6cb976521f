"""GPU: random-shape check of the split-fp16 GEMM arm (all three forms, fused epilogues, tails) against fp64.
Exploration tool, not collected by pytest.   python tools/gpu_gemm_fuzz.py [n_shapes] [seed]"""
import os, sys, random
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gotennet_b200 import ops
dev = torch.device("cuda:0")
n_shapes = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)


def rel(a, b):
    return ((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


worst = 0.0
for it in range(n_shapes):
    M = rng.choice([1, 7, 31, 128, 129, 255, 256, 257, 511, 1000, 2049, rng.randint(1, 5000)])
    N = rng.choice([16, 20, 64, 68, 128, 132, 256, 260, 512, 1792, 4 * rng.randint(4, 200)])
    K = rng.choice([4, 8, 60, 64, 68, 128, 256, 260, 1024, 4 * rng.randint(1, 600)])
    g = torch.Generator().manual_seed(it)
    a = torch.randn(M, K, generator=g).to(dev) * 10.0 ** rng.randint(-6, 3)
    w = torch.randn(N, K, generator=g).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    add = torch.randn(M, N, generator=g).to(dev)
    gr = torch.randn(M, N, generator=g).to(dev) * 10.0 ** rng.randint(-9, 0)
    y = torch.empty(M, N, device=dev); act = torch.empty(M, N, device=dev)
    da = torch.empty(M, K, device=dev); dw = torch.empty(N, K, device=dev); db = torch.empty(N, device=dev)
    ops.gemm(a, K, 0, w, K, 1, y, N, M, N, K, bias=b, add_src=add, ld_add=N, act_out=act, ld_act=N, act_lo=0, act_hi=N, impl=3)
    ops.gemm(gr, N, 0, w, K, 0, da, K, M, K, N, impl=3) if K >= 16 else None
    ok_tn = N % 32 == 0 and K >= 16
    if ok_tn:
        ops.gemm(gr, N, 1, a, K, 0, dw, K, N, K, M, colsum=db, impl=3)
    torch.cuda.synchronize()
    ad, wd, gd = a.double(), w.double(), gr.double()
    ref = ad @ wd.T + b.double() + add.double()
    errs = [rel(y, ref), rel(act, torch.nn.functional.silu(ref))]
    if K >= 16:
        errs.append(rel(da, gd @ wd))
    if ok_tn:
        errs += [rel(dw, gd.T @ ad), rel(db, gd.sum(0))]
    worst = max(worst, max(errs))
    flag = "" if max(errs) < 3e-5 else "   <-- HIGH"
    print(f"M={M:5d} N={N:5d} K={K:5d} " + " ".join(f"{e:.1e}" for e in errs) + flag, flush=True)
print("worst", worst)
