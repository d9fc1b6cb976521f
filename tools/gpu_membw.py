"""GPU: write-only, read-only and copy HBM bandwidth (torch fill_ / sum / copy_ over 4 GiB), for the store-bound
GEMM discussion in DESIGN.md.   python tools/gpu_membw.py"""
import torch
dev = torch.device("cuda:0")
n = 1 << 30
a = torch.empty(n, device=dev); b = torch.empty(n, device=dev)


def t(fn, k=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(k):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


print(f"fill_ (write only): {4 * n / t(lambda: a.fill_(1.0)) / 1e6:.0f} GB/s")
print(f"sum   (read only):  {4 * n / t(lambda: a.sum()) / 1e6:.0f} GB/s")
print(f"copy_ (read+write): {8 * n / t(lambda: b.copy_(a)) / 1e6:.0f} GB/s")
