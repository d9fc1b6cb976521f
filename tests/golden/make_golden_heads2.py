"""Golden vectors for the remaining QM9 read-out heads from the VERBATIM reference: reference
GotenNetWrapper + reference Dipole / ElectronicSpatialExtentV2 (models/components/outputs.py:379-542) executed under
oracle/ref_standins.py.

    python tests/golden/make_golden_heads2.py        (build container only: needs /root/reference)

Stored per case: inputs, the head's outputs, d/dpos and gradient fingerprints of loss = (y * w).sum() (+ the vector
output for Dipole) for every head and representation parameter.  The generator asserts the oracle restatement
(oracle.gotennet_oracle.dipole_forward / spatial_extent_forward) against the reference while it runs."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import gotennet_oracle as orc  # noqa: E402
from oracle.golden_cases import HEAD2_CASES, blob, grad_fingerprint, head2_loss  # noqa: E402
from oracle.ref_standins import import_reference  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import build_reference  # noqa: E402


def head_loss(kind, res, n_mol):
    return head2_loss(kind, res["property"], res.get("property_vector"), n_mol)


def run_case(ref, name, spec, save=True):
    from gotennet.models.components.outputs import Dipole, ElectronicSpatialExtentV2

    cfg, kind = spec["cfg"], spec["kind"]
    z, pos, batch = blob(spec["atoms"], spec["seed"])
    n_mol = len(spec["atoms"])
    C = cfg.n_atom_basis
    sd = orc.make_state_dict(cfg, seed=spec["seed"])
    rep = build_reference(ref, cfg)
    rep.load_state_dict(orc.expand_aliases(sd), strict=True)
    if kind == "dipole":
        sdh = orc.make_dipole_state_dict(C, seed=spec["seed"])
        mean = None if spec["mean"] is None else torch.tensor(spec["mean"])
        std = None if spec["stddev"] is None else torch.tensor(spec["stddev"])
        head = Dipole(n_in=C, predict_magnitude=spec["predict_magnitude"], property="property", mean=mean, stddev=std)
        head.load_state_dict(sdh, strict=True)
    else:
        sdh = {k: v for k, v in orc.make_head_state_dict(C, seed=spec["seed"], atomref=False).items()}
        head = ElectronicSpatialExtentV2(n_in=C, property="property", contributions="contrib")
        missing = head.load_state_dict(sdh, strict=False)
        assert missing.missing_keys == ["atomic_mass"] and not missing.unexpected_keys, missing

    class Data:
        def __getitem__(self, k):
            return getattr(self, k)

    d = Data()
    d.z, d.pos, d.batch = z, pos.clone().requires_grad_(True), batch
    h, X = rep(d)
    d.representation, d.vector_representation = h, X
    res = head(d)
    head_loss(kind, res, n_mol).backward()
    out = dict(z=z.numpy(), pos=pos.numpy(), batch=batch.numpy(), y=res["property"].detach().numpy(),
               grad_pos=d.pos.grad.numpy())
    if kind == "dipole":
        out["y_vector"] = res["property_vector"].detach().numpy()
    else:
        out["contrib"] = res["contrib"].detach().numpy()
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
    sdho = {k: v.clone().requires_grad_(k.startswith("out_net") or k.startswith("equivariant")) for k, v in sdh.items()}
    pos_o = pos.clone().requires_grad_(True)
    ho, Xo = orc.wrapper_forward(sdo, cfg, z, pos_o, batch)
    if kind == "dipole":
        y, yv = orc.dipole_forward(sdho, ho, Xo, pos_o, batch, n_mol, mean=spec["mean"], stddev=spec["stddev"],
                                   predict_magnitude=spec["predict_magnitude"])
        res_o = {"property": y, "property_vector": yv}
    else:
        from gotennet_b200.atomic_data import ATOMIC_MASSES
        y, x = orc.spatial_extent_forward(sdho, ho, z, pos_o, batch, n_mol, torch.tensor(ATOMIC_MASSES))
        res_o = {"property": y}
    head_loss(kind, res_o, n_mol).backward()

    def rel(a, b):
        return (a.detach() - b.detach()).abs().max().item() / max(b.detach().abs().max().item(), 1e-30)

    worst = rel(pos_o.grad, d.pos.grad)
    for k in out:
        if k.startswith("gradh_"):
            worst = max(worst, rel(grad_fingerprint(sdho[k[6:]].grad), torch.from_numpy(out[k])))
        elif k.startswith("grad_") and k != "grad_pos":
            g = sdo[k[5:]].grad
            worst = max(worst, rel(grad_fingerprint(g if g is not None else torch.zeros_like(sdo[k[5:]])),
                                   torch.from_numpy(out[k])))
    print(f"{name}: N={z.numel()} oracle-vs-reference rel err y {rel(res_o['property'], res['property']):.2e} "
          f"grads {worst:.2e}")
    assert rel(res_o["property"], res["property"]) < 5e-6 and worst < 1e-4, name
    if save:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    return out


def main():
    torch.set_num_threads(4)
    ref = import_reference()
    for name, spec in HEAD2_CASES.items():
        run_case(ref, name, spec)


if __name__ == "__main__":
    main()
