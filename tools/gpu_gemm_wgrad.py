"""GPU: weight-gradient GEMM shapes (C[M,N] = A[R,M]^T B[R,N], fused column sums), timing + error vs fp64.
    python tools/gpu_gemm_wgrad.py      (GOTEN_GEMM_CONV8=1: eight converter warps)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gotennet_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
for M, N, R in ((1792, 256, 301491), (1536, 256, 301491), (1280, 256, 18471), (1024, 256, 18471), (256, 256, 147768),
                (512, 256, 92355), (256, 1280, 18471), (512, 32, 301491)):
    a = torch.randn(R, M, device=dev); b = torch.randn(R, N, device=dev)
    out = torch.empty(M, N, device=dev); cs = torch.empty(M, device=dev)
    am, bm = ops.absmax(a, M, R, M), ops.absmax(b, N, R, N)
    fn = lambda: ops.gemm(a, M, 1, b, N, 0, out, N, M, N, R, colsum=cs, impl=3, a_amax=am, b_amax=bm)
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(6): fn()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 6
    msg = f"M={M:5d} N={N:5d} R={R:7d}: {t * 1e3:7.1f} us  {2.0 * M * N * R / t / 1e9:6.0f} TF/s"
    if M * N * R < 3e10:
        ref = a.double().T @ b.double()
        msg += "  err %.1e cs %.1e" % (((out.double() - ref).abs().max() / ref.abs().max()).item(),
                                       ((cs.double() - a.double().sum(0)).abs().max() / a.double().sum(0).abs().max()).item())
    print(msg, flush=True)
