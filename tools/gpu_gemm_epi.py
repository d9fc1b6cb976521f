"""GPU: kernel-only timing + accuracy of the node-level GEMM shapes whose epilogue dominates (residual add, SiLU side
output, short K) at the cfg2 sizes.  Exploration tool, not collected by pytest.
    python tools/gpu_gemm_epi.py            (GOTEN_LIB_PATH=... to time another build of the library)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gotennet_b200 import ops
dev = torch.device("cuda:0")
Nn, E = 18471, 301491
torch.manual_seed(0)


def timeit(fn, n=8):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def rel(a, b):
    return ((a.double() - b).abs().max() / b.abs().max()).item()


cases = [  # (label, M, N, K, trans_b, add, act)
    ("eqff dgrad + residual", 8 * Nn, 256, 256, 0, True, False),
    ("htr dgrad + residual", 5 * Nn, 256, 512, 0, True, False),
    ("g_h dgrad + residual", Nn, 256, 1024, 0, True, False),
    ("edge dgrad + residual", E, 256, 1792, 0, True, False),
    ("node proj + silu side", Nn, 1024, 256, 1, False, True),
    ("eqff gamma_m.0 + silu", Nn, 256, 512, 1, False, True),
    ("eqff W_vu fwd (plain)", 8 * Nn, 256, 256, 1, False, False),
    ("edge fwd W_re|W_rs|g_t", E, 1792, 256, 1, False, False),
    ("gamma_s.1 fwd (plain)", Nn, 1280, 256, 1, False, False),
    ("htr EQ|EK fwd (plain)", 5 * Nn, 512, 256, 1, False, False),
]
for label, M, N, K, tb, add, act in cases:
    a = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5 if tb else torch.randn(K, N, device=dev) / K ** 0.5
    r = torch.randn(M, N, device=dev) if add else None
    out = torch.empty(M, N, device=dev)
    so = torch.empty(M, N, device=dev) if act else None
    am, bm = ops.absmax(a, K, M, K), ops.absmax(w, w.shape[1], w.shape[0], w.shape[1])
    inplace = add and os.environ.get("EPI_INPLACE", "1") == "1"   # C == add_src: TMA reduction-store epilogue
    if inplace:
        out.copy_(r)
    fn = lambda: ops.gemm(a, K, 0, w, w.shape[1], tb, out, N, M, N, K, add_src=out if inplace else r, ld_add=N, act_out=so,
                          ld_act=N, act_lo=0, act_hi=N if act else 0, impl=3, a_amax=am, b_amax=bm)
    t = timeit(fn)
    byt = 4.0 * (M * K + M * N * (1 + add + act))
    msg = f"{label:24s} M={M:7d} N={N:5d} K={K:5d}: {t * 1e3:7.1f} us  {2.0 * M * N * K / t / 1e9:6.0f} TF/s  {byt / t / 1e6:6.0f} GB/s"
    if M * N * K < 4e11:
        if inplace:
            out.copy_(r)
            fn()
        ref = a.double() @ (w.double().T if tb else w.double()) + (r.double() if add else 0)
        msg += f"  err {rel(out, ref):.1e}"
        if act:
            msg += f" silu {rel(so, torch.nn.functional.silu(ref)):.1e}"
    print(msg, flush=True)
