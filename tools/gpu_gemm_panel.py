"""GPU: correctness + timing of the A-stationary panel GEMM (GOTEN_GEMM_PANEL=1) on the short-K forward shapes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gotennet_b200 import ops
dev = torch.device("cuda:0")
for M, N, K in ((1000, 512, 256), (4096, 1792, 256), (777, 1024, 64), (5000, 1280, 192), (513, 300, 100)):
    g = torch.Generator().manual_seed(M)
    a = torch.randn(M, K, generator=g); w = torch.randn(N, K, generator=g); b = torch.randn(N, generator=g)
    out = torch.empty(M, N, device=dev); act = torch.empty(M, N, device=dev)
    ops.gemm(a.to(dev), K, 0, w.to(dev), K, 1, out, N, M, N, K, bias=b.to(dev), act_out=act, ld_act=N, act_lo=0, act_hi=N, impl=3)
    ref = a.double() @ w.double().T + b.double()
    err = ((out.cpu().double() - ref).abs().max() / ref.abs().max()).item()
    erra = ((act.cpu().double() - torch.nn.functional.silu(ref)).abs().max() / ref.abs().max()).item()
    print(M, N, K, "err %.2e silu %.2e" % (err, erra), flush=True)
    assert err < 2e-5 and erra < 2e-5
