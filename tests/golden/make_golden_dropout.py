"""Golden vectors for training-mode attention dropout, from the VERBATIM reference (build container only).

    python tests/golden/make_golden_dropout.py

The reference applies `F.dropout(attn, p, training)` to the scaled attention weights [E, H, 1]
(representation/gotennet.py:513).  Its random stream cannot be reproduced by another implementation, so the
generator lets the reference draw its own masks (the real F.dropout runs; a wrapper only RECORDS which entries
survived), stores them next to inputs / outputs / gradients, and checks at generation time that the oracle fed
with the same masks reproduces the reference."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import gotennet_oracle as orc  # noqa: E402
from oracle.golden_cases import DROPOUT_CASES, blob, grad_fingerprint  # noqa: E402
from oracle.ref_standins import import_reference  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    torch.set_num_threads(4)
    ref = import_reference()
    import gotennet.models.representation.gotennet as ref_mod
    from gotennet.models.components.layers import CosineCutoff

    for name, spec in DROPOUT_CASES.items():
        cfg, p = spec["cfg"], spec["p"]
        z, pos, batch = blob(spec["atoms"], spec["seed"])
        sd = orc.make_state_dict(cfg, seed=spec["seed"])
        model = ref.GotenNetWrapper(
            n_atom_basis=cfg.n_atom_basis, n_interactions=cfg.n_interactions, n_rbf=cfg.n_rbf,
            cutoff_fn=CosineCutoff(cfg.cutoff), max_z=cfg.max_z, epsilon=cfg.epsilon, num_heads=cfg.num_heads,
            attn_dropout=p, edge_updates=cfg.edge_updates, scale_edge=cfg.scale_edge, lmax=cfg.lmax,
            sep_htr=cfg.sep_htr, sep_dir=cfg.sep_dir, sep_tensor=cfg.sep_tensor, max_num_neighbors=cfg.max_num_neighbors)
        model.load_state_dict(orc.expand_aliases(sd), strict=True)
        model.train()
        masks = []
        real_dropout = ref_mod.F.dropout

        def recording_dropout(x, p=0.5, training=True, inplace=False):
            out = real_dropout(x, p=p, training=training, inplace=False)
            assert bool((x != 0).all()), "softmax weights are strictly positive"
            masks.append((out != 0).reshape(x.shape[0], x.shape[1]).clone())
            return out

        ref_mod.F.dropout = recording_dropout
        try:
            torch.manual_seed(1234 + spec["seed"])

            class Data:
                pass

            d = Data()
            d.z, d.pos, d.batch = z, pos.clone().requires_grad_(True), batch
            h, X = model(d)
            loss = h.sum() + X.pow(2).sum()
            loss.backward()
        finally:
            ref_mod.F.dropout = real_dropout
        assert len(masks) == cfg.n_interactions
        kept = float(torch.stack([m.float().mean() for m in masks]).mean())
        out = dict(z=z.numpy(), pos=pos.numpy(), batch=batch.numpy(), h=h.detach().numpy(), X=X.detach().numpy(),
                   loss=np.float64(loss.item()), grad_pos=d.pos.grad.numpy(), p=np.float64(p),
                   masks=np.stack([m.numpy() for m in masks]))
        seen = set()
        for k, prm in model.named_parameters():
            key = k.replace(".layers.", ".dense_layers.") if ("W_ndp" in k or "W_nrd_nru" in k or "gamma_t" in k) else k
            if key in seen:
                continue
            seen.add(key)
            g = prm.grad if prm.grad is not None else torch.zeros_like(prm)
            out["grad_" + key] = grad_fingerprint(g).numpy()
        # oracle with the recorded masks
        drop = [m.float() / (1.0 - p) for m in masks]
        pos_o = pos.clone().requires_grad_(True)
        sd_o = {k: v.clone().requires_grad_(v.is_floating_point() and "radial_basis" not in k) for k, v in sd.items()}
        ho, Xo = orc.wrapper_forward(sd_o, cfg, z, pos_o, batch, drop_masks=drop)
        (ho.sum() + Xo.pow(2).sum()).backward()
        eh = (ho - h).abs().max().item() / h.abs().max().item()
        eX = (Xo - X).abs().max().item() / X.abs().max().item()
        ep = (pos_o.grad - d.pos.grad).abs().max().item() / d.pos.grad.abs().max().item()
        wg = 0.0
        for k in seen:
            fp = torch.from_numpy(out["grad_" + k])
            got = grad_fingerprint(sd_o[k].grad if sd_o[k].grad is not None else torch.zeros_like(sd_o[k]))
            wg = max(wg, (got - fp).abs().max().item() / max(fp.abs().max().item(), 1e-30))
        # and the masks matter: eval-mode oracle must differ
        h_eval, _ = orc.wrapper_forward(sd, cfg, z, pos, batch)
        diff_eval = (h_eval - h).abs().max().item() / h.abs().max().item()
        print(f"{name}: E={masks[0].shape[0]} kept {kept:.3f} oracle-vs-reference h {eh:.2e} X {eX:.2e} dpos {ep:.2e} "
              f"dparam {wg:.2e}; eval-mode differs by {diff_eval:.2e}")
        assert max(eh, eX, ep) < 2e-5 and wg < 1e-4 and diff_eval > 1e-3, name
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    main()
