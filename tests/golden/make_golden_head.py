"""Golden vectors for the read-out head from the VERBATIM reference:
reference GotenNetWrapper + reference Atomwise(derivative="forces") executed under oracle/ref_standins.py.

    python tests/golden/make_golden_head.py        (build container only: needs /root/reference)

Stored per case: inputs, energy [n_mol,1], forces [N,3], per-atom contributions, and gradient fingerprints of
loss = (E * w).sum() for every head parameter and every representation parameter.  The generator asserts the
oracle restatement (oracle.gotennet_oracle.atomwise_forward / energy_and_forces) against the reference."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import gotennet_oracle as orc  # noqa: E402
from oracle.golden_cases import HEAD_CASES, blob, grad_fingerprint, probe_vector  # noqa: E402
from oracle.ref_standins import import_reference  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import build_reference  # noqa: E402


def run_case(ref, name, spec, save=True):
    from gotennet.models.components.layers import shifted_softplus
    from gotennet.models.components.outputs import Atomwise

    cfg, act = spec["cfg"], spec["activation"]
    z, pos, batch = blob(spec["atoms"], spec["seed"])
    n_mol = len(spec["atoms"])
    sd = orc.make_state_dict(cfg, seed=spec["seed"])
    sdh = orc.make_head_state_dict(cfg.n_atom_basis, seed=spec["seed"])
    rep = build_reference(ref, cfg)
    rep.load_state_dict(orc.expand_aliases(sd), strict=True)
    head = Atomwise(n_in=cfg.n_atom_basis, activation=F.silu if act == "silu" else shifted_softplus,
                    mean=sdh["standardize.mean"].clone(), stddev=sdh["standardize.stddev"].clone(),
                    atomref=sdh["atomref.weight"].clone(), property="property", contributions="contrib",
                    derivative="forces")
    head.load_state_dict(sdh, strict=True)
    head.atomref.weight.requires_grad_(False)

    class Data:  # PyG Data offers attribute AND item access (GetItem uses the latter, layers.py:223)
        def __getitem__(self, k):
            return getattr(self, k)

    d = Data()
    d.z, d.pos, d.batch = z, pos.clone().requires_grad_(True), batch
    h, X = rep(d)                                   # goten_model.py:289
    d.representation, d.vector_representation = h, X
    res = head(d)                                   # outputs.py:323-376
    E, Fo, yi = res["property"], res["forces"], res["contrib"]
    w = probe_vector(n_mol).unsqueeze(1)
    ((E * w).sum()).backward()
    out = dict(z=z.numpy(), pos=pos.numpy(), batch=batch.numpy(), energy=E.detach().numpy(), forces=Fo.detach().numpy(),
               contrib=yi.detach().numpy(), h=h.detach().numpy())
    for k, p in head.named_parameters():
        if p.grad is not None:
            out["gradh_" + k] = grad_fingerprint(p.grad).numpy()
    seen = set()
    for k, p in rep.named_parameters():
        key = k.replace(".layers.", ".dense_layers.") if ("W_ndp" in k or "W_nrd_nru" in k or "gamma_t" in k) else k
        if key in seen:
            continue
        seen.add(key)
        out["grad_" + key] = grad_fingerprint(p.grad if p.grad is not None else torch.zeros_like(p)).numpy()

    # oracle restatement vs the reference
    sdo = {k: v.clone().requires_grad_(v.is_floating_point() and "radial_basis" not in k) for k, v in sd.items()}
    sdho = {k: v.clone().requires_grad_(k.startswith("out_net")) for k, v in sdh.items()}
    Eo, Fo_o, ho = orc.energy_and_forces(sdo, sdho, cfg, z, pos, batch, n_mol, act)
    ((Eo * w).sum()).backward()

    def rel(a, b):
        return (a.detach() - b.detach()).abs().max().item() / max(b.detach().abs().max().item(), 1e-30)

    worst = 0.0
    for k in out:
        if k.startswith("gradh_"):
            worst = max(worst, rel(grad_fingerprint(sdho[k[6:]].grad), torch.from_numpy(out[k])))
        elif k.startswith("grad_"):
            g = sdo[k[5:]].grad
            worst = max(worst, rel(grad_fingerprint(g if g is not None else torch.zeros_like(sdo[k[5:]])), torch.from_numpy(out[k])))
    print(f"{name}: N={z.numel()} oracle-vs-reference rel err E {rel(Eo, E):.2e} F {rel(Fo_o, Fo):.2e} dparam {worst:.2e}")
    assert rel(Eo, E) < 2e-6 and rel(Fo_o, Fo) < 2e-5 and worst < 1e-4, name
    if save:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    return out


def main():
    torch.set_num_threads(4)
    ref = import_reference()
    for name, spec in HEAD_CASES.items():
        run_case(ref, name, spec)


if __name__ == "__main__":
    main()
