"""Times the other BASELINE.json configurations at their full sizes on one GPU (exploration / DESIGN.md numbers;
bench.py keeps the contract workload, configs[1]).

    python tools/run_workload.py rmd17 [--batch 4096] [--steps 5]     # configs[2]: 21 atoms, 6 layers, force head
    python tools/run_workload.py md22  [--batch 64]   [--steps 5]     # configs[3]: 370 atoms, lmax=3, K=160
    python tools/run_workload.py qm9   [--batch 1024]

A step = radius graph + geometry + forward + backward (all parameter gradients).  For rmd17 the loss is the summed
Atomwise energy and positions require grad, so pos.grad is minus the forces (first-order; outputs.py:365-375).
Prints one JSON line.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

BASE = dict(n_atom_basis=256, n_rbf=32, num_heads=8, sep_htr=True, sep_dir=True, sep_tensor=True, scale_edge=False,
            edge_updates=True)
WORKLOADS = {
    "qm9": dict(kind="qm9", batch=1024, max_nbr=32, head=False, model=dict(n_interactions=4, lmax=2)),
    "rmd17": dict(kind="aspirin", batch=4096, max_nbr=32, head=True, model=dict(n_interactions=6, lmax=2)),
    "md22": dict(kind="md22", batch=64, max_nbr=160, head=False, model=dict(n_interactions=4, lmax=3)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--profile", action="store_true", help="per-entry-point CUDA-event table of one extra step (stderr)")
    args = ap.parse_args()
    import gotennet_b200 as g
    from gotennet_b200._lib import lib
    from gotennet_b200.synthetic import synth_batch

    w = WORKLOADS[args.workload]
    B = args.batch or w["batch"]
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = g.GotenNetWrapper(cutoff_fn=g.CosineCutoff(5.0), max_num_neighbors=w["max_nbr"], activation="swish",
                              **BASE, **w["model"]).to(dev)
    head = g.Atomwise(n_in=256, n_out=1, aggregation_mode="sum", activation="swish").to(dev) if w["head"] else None
    z, pos, batch = synth_batch(w["kind"], B, seed=1000)
    zd, pd, bd = z.to(dev), pos.to(dev), batch.to(dev)
    params = list(model.parameters()) + (list(head.parameters()) if head else [])

    class Data:
        pass

    def step():
        d = Data()
        d.z, d.pos, d.batch = zd, (pd.clone().requires_grad_(True) if head is not None else pd), bd
        d.num_graphs = B
        for p in params:
            p.grad = None
        h, X = model(d)
        if head is not None:
            d.representation, d.vector_representation = h, X
            loss = head(d)["y"].sum()
        else:
            loss = h.sum() + X.pow(2).sum()
        loss.backward()
        return loss, d.pos.grad

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    L = lib()
    n0 = L.cdll.goten_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss, gpos = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if args.profile:
        L.profile = []
        step()
        torch.cuda.synchronize()
        prof, L.profile = L.profile, None
        agg = {}
        for name, a, p0, p1 in prof:
            d = agg.setdefault(name, [0, 0.0])
            d[0] += 1
            d[1] += p0.elapsed_time(p1)
        for name, (cnt, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
            print(f"  {name:30s} {cnt:4d}x {t:8.3f} ms {100 * t / ms:5.1f}%", file=sys.stderr)
    plan = model.last_plan
    finite = bool(torch.isfinite(loss).item()) and (gpos is None or bool(torch.isfinite(gpos).all().item()))
    print(json.dumps({
        "workload": args.workload, "molecules": B, "atoms": plan.N, "edges": plan.E, "max_in_degree": plan.max_deg_in,
        "model": {**BASE, **w["model"]}, "max_num_neighbors": w["max_nbr"], "force_head": w["head"],
        "ms_per_step": ms, "molecules_per_s": B / (ms * 1e-3), "launches_per_step": (L.cdll.goten_launch_count() - n0) // args.steps,
        "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9, "finite": finite}), flush=True)


if __name__ == "__main__":
    main()
