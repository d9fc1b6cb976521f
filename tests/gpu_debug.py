"""Stage-by-stage comparison of the CUDA path with the oracle (run on the GPU box):
    python tests/gpu_debug.py [case ...]
Prints relative errors per stage so a failing kernel can be localised in one call."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gotennet_b200 as g  # noqa: E402
from oracle import gotennet_oracle as orc  # noqa: E402
from oracle.golden_cases import CASES, blob, grad_fingerprint  # noqa: E402


def rel(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def build(cfg, sd, dev):
    m = g.GotenNetWrapper(n_atom_basis=cfg.n_atom_basis, n_interactions=cfg.n_interactions, n_rbf=cfg.n_rbf,
                          cutoff_fn=g.CosineCutoff(cfg.cutoff), max_z=cfg.max_z, epsilon=cfg.epsilon,
                          num_heads=cfg.num_heads, edge_updates=cfg.edge_updates, scale_edge=cfg.scale_edge,
                          lmax=cfg.lmax, sep_htr=cfg.sep_htr, sep_dir=cfg.sep_dir, sep_tensor=cfg.sep_tensor,
                          max_num_neighbors=cfg.max_num_neighbors)
    m.load_state_dict(orc.expand_aliases(sd), strict=True)
    return m.to(dev)


class Data:
    pass


def run(name):
    spec = CASES[name]
    cfg = spec["cfg"]
    dev = torch.device("cuda:0")
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    z, pos, batch = blob(spec["atoms"], spec["seed"])
    sd = orc.make_state_dict(cfg, seed=spec["seed"])
    m = build(cfg, sd, dev)
    m._capture = {}
    d = Data()
    d.z, d.pos, d.batch = z.to(dev), pos.to(dev).requires_grad_(True), batch.to(dev)
    h, X = m(d)
    torch.cuda.synchronize()
    plan = m.last_plan
    print(f"== {name}: N={plan.N} E={plan.E} (gold E={gold['edge_index'].shape[1]})")
    ok_ei = plan.E == gold["edge_index"].shape[1] and np.array_equal(plan.edge_index.cpu().numpy(), gold["edge_index"])
    print("   edge_index bit-exact:", ok_ei)
    inter = {}
    sdo = {k: v.clone().requires_grad_("radial_basis" not in k) for k, v in sd.items()}
    pos_o = pos.clone().requires_grad_(True)
    ho, Xo = orc.wrapper_forward(sdo, cfg, z, pos_o, batch, inter)
    cap = m._capture
    for k in ["phi", "Y", "h0", "t0"] + [f"{s}{i + 1}" for i in range(cfg.n_interactions) for s in ("h", "X", "t")]:
        if ok_ei or k[0] in "hX":
            print(f"   {k:4s} rel err vs oracle {rel(cap[k], inter[k].detach()):.2e}")
    print(f"   h vs golden {rel(h.detach(), gold['h']):.2e}   X vs golden {rel(X.detach(), gold['X']):.2e}")
    (h.sum() + X.pow(2).sum()).backward()
    torch.cuda.synchronize()
    print(f"   grad_pos vs golden {rel(d.pos.grad, gold['grad_pos']):.2e}")
    worst = []
    params = dict(m.named_parameters())
    for k in gold.files:
        if k.startswith("grad_") and k != "grad_pos":
            p = params[k[5:]]
            gr = p.grad if p.grad is not None else torch.zeros_like(p)
            worst.append((rel(grad_fingerprint(gr.cpu()), gold[k]), k[5:]))
    worst.sort(reverse=True)
    for e, k in worst[:8]:
        print(f"   dparam {k:50s} {e:.2e}")


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for n in names:
        try:
            run(n)
        except Exception as e:  # keep going: one call should report on every case
            import traceback
            traceback.print_exc()
            print(f"== {n}: FAILED {type(e).__name__}: {e}")
