"""GPU: accuracy and speed of the tensor-core GEMM arms (2 = tcgen05 3xTF32, 3 = tcgen05 split-fp16) vs the
fp32 SIMT arm (1), all through goten_gemm_scaled.  Exploration tool, not collected by pytest.
    python tools/gpu_gemm_check.py [quick] [impls=1,2,3] [scale=1e-6]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gotennet_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def rel(a, b):
    return (a.double() - b).abs().max().item() / b.abs().max().item()


def timeit(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


IMPLS = (1, 2, 3)
SCALE = 1.0
for a_ in sys.argv[1:]:
    if a_.startswith("impls="):
        IMPLS = tuple(int(x) for x in a_[6:].split(","))
    if a_.startswith("scale="):
        SCALE = float(a_[6:])


def check(M, N, K, perf=False):
    g = torch.Generator(device="cpu").manual_seed(M + 3 * N + 7 * K)
    a = (torch.randn(M, K, generator=g)).to(dev)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    add = torch.randn(M, N, generator=g).to(dev)
    gr = (torch.randn(M, N, generator=g) * SCALE).to(dev)
    if SCALE != 1.0:  # wide dynamic range inside the tensors: a few large outliers
        a[::7, ::5] *= 300.0
        gr[::3, ::11] *= 1e-4
    exact = M * N * K < 3e10
    out = {}
    for impl in IMPLS:
        y = torch.empty(M, N, device=dev)
        act = torch.empty(M, N, device=dev)
        da = torch.empty(M, K, device=dev)
        dw = torch.empty(N, K, device=dev)
        db = torch.empty(N, device=dev)
        f_nt = lambda: ops.gemm(a, K, 0, w, K, 1, y, N, M, N, K, bias=b, add_src=add, ld_add=N, act_out=act, ld_act=N,
                                act_lo=0, act_hi=N, impl=impl)
        f_nn = lambda: ops.gemm(gr, N, 0, w, K, 0, da, K, M, K, N, impl=impl)
        f_tn = lambda: ops.gemm(gr, N, 1, a, K, 0, dw, K, N, K, M, colsum=db, impl=impl)
        try:
            f_nt(); f_nn(); f_tn()
            torch.cuda.synchronize()
        except Exception as e:
            print(f"  impl={impl} FAILED: {e}")
            continue
        out[impl] = (y, act, da, dw, db)
        msg = f"  impl={impl}"
        if exact:
            ad, wd = a.double(), w.double()
            ref = ad @ wd.T + b.double() + add.double()
            msg += (f" NT {rel(y, ref):.1e} act {rel(act, torch.nn.functional.silu(ref)):.1e}"
                    f" NN {rel(da, gr.double() @ wd):.1e} TN {rel(dw, gr.double().T @ ad):.1e}"
                    f" colsum {rel(db, gr.double().sum(0)):.1e}")
        if perf:
            fl = 2.0 * M * N * K
            msg += "  TF/s: " + " ".join(f"{nm} {fl / (timeit(f) * 1e-3) / 1e12:.1f}" for nm, f in
                                        (("NT", f_nt), ("NN", f_nn), ("TN", f_tn)))
        print(msg, flush=True)
    if not exact:
        for i in IMPLS[1:]:
            if IMPLS[0] in out and i in out:
                print(f"  impl {i} vs {IMPLS[0]}:", " ".join(f"{rel(x, y_.double()):.1e}" for x, y_ in zip(out[i], out[IMPLS[0]])))


if __name__ == "__main__":
    quick = "quick" in sys.argv[1:]
    shapes = [(128, 256, 32, False), (1000, 256, 256, False), (4096, 1792, 256, False), (333, 64, 64, False),
              (2048, 96, 160, False), (777, 256, 1000, False), (4099, 512, 72, False)]
    if not quick:
        shapes += [(301491, 1792, 256, True), (147768, 256, 256, True), (18471, 1024, 256, True)]
    for M, N, K, perf in shapes:
        print(f"M={M} N={N} K={K}", flush=True)
        check(M, N, K, perf)
