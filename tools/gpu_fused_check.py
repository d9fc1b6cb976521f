"""GPU: fused edge kernel (edge_fused.cu) against the unfused sequence (GEMM + attention + message kernels) on the
cfg2 model: outputs, all gradients, timing of forward-only and forward+backward.  Exploration / A-B tool.
    python tools/gpu_fused_check.py [n_mol]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gotennet_b200 as g
from gotennet_b200.synthetic import synth_batch
import bench

dev = torch.device("cuda:0")
n_mol = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
model = g.GotenNetWrapper(cutoff_fn=g.CosineCutoff(5.0), max_num_neighbors=32, activation="swish", **bench.MODEL).to(dev)
z, pos, batch = synth_batch("qm9", n_mol, seed=7)


class D:
    pass


def run(fused, grad=True):
    os.environ["GOTEN_EDGE_FUSED"] = "1" if fused else "0"
    d = D()
    d.z, d.pos, d.batch = z.to(dev), pos.to(dev).requires_grad_(grad), batch.to(dev)
    for p in model.parameters():
        p.grad = None
    if not grad:
        with torch.no_grad():
            h, X = model(d)
        return h, X, None, None
    h, X = model(d)
    (h.sum() + X.pow(2).sum()).backward()
    return h.detach(), X.detach(), d.pos.grad.clone(), {k: p.grad.clone() for k, p in model.named_parameters()}


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp(min=1e-30)).item()


h0, X0, gp0, g0 = run(False)
h1, X1, gp1, g1 = run(True)
print(f"fused vs unfused: h {rel(h1, h0):.2e}  X {rel(X1, X0):.2e}  dpos {rel(gp1, gp0):.2e}  "
      f"worst param grad {max(rel(g1[k], g0[k]) for k in g0):.2e}")
hi, Xi, _, _ = run(True, grad=False)
print(f"inference (gamma_t-only Ze store) vs training forward: h {rel(hi, h1):.2e}  X {rel(Xi, X1):.2e}")
a, b, _, _ = run(True, grad=False)
print("bit-reproducible:", bool(torch.equal(a, hi) and torch.equal(b, Xi)))


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for fused in (False, True, False, True):
    tf = timeit(lambda: run(fused, grad=False))
    tb = timeit(lambda: run(fused, grad=True)) if not os.environ.get("GOTEN_FUSED_DBG") else 0.0
    print(f"fused={int(fused)}  forward-only {tf:.2f} ms   forward+backward {tb:.2f} ms   ({n_mol} molecules)")
