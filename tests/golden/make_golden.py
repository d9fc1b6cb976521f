"""Generate golden vectors by executing the VERBATIM reference code.

Run in the build container only (needs /root/reference, absent on the GPU box):

    python tests/golden/make_golden.py

For every case below it
  1. builds the unmodified reference `GotenNetWrapper` (imported under
     oracle/ref_standins.py), loads the deterministic synthetic weights from
     oracle.gotennet_oracle.make_state_dict (biases / LayerNorm affine non-zero),
  2. runs forward + backward of  loss = h.sum() + X.pow(2).sum()  in float32,
  3. stores inputs, edge_index, outputs, per-layer states and gradient
     fingerprints in tests/golden/<case>.npz.

The fixtures are what pins the oracle restatement (tests/test_oracle_golden.py)
and the CUDA path (tests/test_gpu_parity.py).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import gotennet_oracle as orc  # noqa: E402
from oracle.golden_cases import CASES, NORM_CASES, blob as _blob, grad_fingerprint as _gfp  # noqa: E402
from oracle.ref_standins import import_reference  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def build_reference(ref, cfg: orc.OracleConfig, **extra):
    from gotennet.models.components.layers import CosineCutoff

    return ref.GotenNetWrapper(
        evec_dim=cfg.evec_dim, edge_ln=cfg.edge_ln, **extra,
        n_atom_basis=cfg.n_atom_basis, n_interactions=cfg.n_interactions, n_rbf=cfg.n_rbf,
        cutoff_fn=CosineCutoff(cfg.cutoff), max_z=cfg.max_z, epsilon=cfg.epsilon, num_heads=cfg.num_heads,
        attn_dropout=0.0, edge_updates=cfg.edge_updates, scale_edge=cfg.scale_edge, lmax=cfg.lmax,
        sep_htr=cfg.sep_htr, sep_dir=cfg.sep_dir, sep_tensor=cfg.sep_tensor,
        max_num_neighbors=cfg.max_num_neighbors, layernorm=cfg.layernorm, steerable_norm=cfg.steerable_norm,
        radial_basis=cfg.radial_basis, emlp_dim=cfg.emlp_dim,
    )


def run_case(ref, name, spec, save=True):
    """Runs one case on the verbatim reference; returns the fixture dict (and writes <name>.npz when `save`)."""
    cfg = spec["cfg"]
    z, pos, batch = _blob(spec["atoms"], spec["seed"])
    sd = orc.make_state_dict(cfg, seed=spec["seed"])
    model = build_reference(ref, cfg, **({"activation": spec["activation"]} if "activation" in spec else {}))
    missing = model.load_state_dict(orc.expand_aliases(sd, cfg), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model.eval()

    class Data:
        pass

    # capture per-layer states via forward hooks on the reference modules
    states = {}
    def gata_hook(i):
        def fn(mod, args, out):
            states[f"t{i + 1}"] = out[2].detach().clone()
        return fn

    def eqff_hook(i):
        def fn(mod, args, out):
            states[f"h{i + 1}"] = out[0].detach().squeeze(1).clone()
            states[f"X{i + 1}"] = out[1].detach().clone()
        return fn

    for i, (g, e) in enumerate(zip(model.gata_list, model.eqff_list)):
        g.register_forward_hook(gata_hook(i))
        e.register_forward_hook(eqff_hook(i))
    captured = {}
    orig_dist = model.distance.forward

    def dist_hook(p, b):
        ei, w, v = orig_dist(p, b)
        captured.update(edge_index=ei.clone(), edge_weight=w.detach().clone())
        return ei, w, v

    model.distance.forward = dist_hook
    d = Data()
    d.z, d.pos, d.batch = z, pos.clone().requires_grad_(True), batch
    h, X = model(d)
    loss = h.sum() + X.pow(2).sum()
    loss.backward()

    out = dict(z=z.numpy(), pos=pos.numpy(), batch=batch.numpy(),
               edge_index=captured["edge_index"].numpy(), edge_weight=captured["edge_weight"].numpy(),
               h=h.detach().numpy(), X=X.detach().numpy(), loss=np.float64(loss.item()),
               grad_pos=d.pos.grad.numpy())
    for k, v in states.items():
        out["state_" + k] = v.numpy()
    seen = set()
    for k, p in model.named_parameters():  # named_parameters de-duplicates aliases
        key = k.replace(".layers.", ".dense_layers.") if ("W_ndp" in k or "W_nrd_nru" in k or "gamma_t" in k) else k
        if ".gamma_w." in key and key not in sd:   # W_edp seen through the gamma_w Sequential first: canonical name
            key = key[:key.index(".gamma_w.")] + ".W_edp." + key.split(".gamma_w.")[1].split(".", 1)[1]
        if key in seen:
            continue
        seen.add(key)
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        out["grad_" + key] = _gfp(g).numpy()

    # cross-check the oracle restatement right here (float32 and float64)
    inter = {}
    pos_o = pos.clone().requires_grad_(True)
    sd_o = {k: v.clone().requires_grad_(v.is_floating_point() and "radial_basis" not in k) for k, v in sd.items()}
    ho, Xo = orc.wrapper_forward(sd_o, cfg, z, pos_o, batch, inter)
    (ho.sum() + Xo.pow(2).sum()).backward()
    assert torch.equal(inter["edge_index"], captured["edge_index"]), name
    err_h = (ho - h).abs().max().item() / h.abs().max().item()
    err_X = (Xo - X).abs().max().item() / max(X.abs().max().item(), 1e-30)
    err_p = (pos_o.grad - d.pos.grad).abs().max().item() / d.pos.grad.abs().max().item()
    worst_g = 0.0
    for k in seen:
        ref_fp = torch.from_numpy(out["grad_" + k])
        got = _gfp(sd_o[k].grad if sd_o[k].grad is not None else torch.zeros_like(sd_o[k]))
        worst_g = max(worst_g, (got - ref_fp).abs().max().item() / max(ref_fp.abs().max().item(), 1e-30))
    print(f"{name}: N={z.numel()} E={captured['edge_index'].shape[1]} oracle-vs-reference rel err "
          f"h {err_h:.2e} X {err_X:.2e} dpos {err_p:.2e} dparam {worst_g:.2e}")
    assert max(err_h, err_X, err_p) < 2e-5 and worst_g < 1e-4, name
    if save:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    return out


def main():
    torch.set_num_threads(4)
    ref = import_reference()
    only = sys.argv[1:]
    for name, spec in {**CASES, **NORM_CASES}.items():
        if not only or name in only:
            run_case(ref, name, spec)


if __name__ == "__main__":
    main()
