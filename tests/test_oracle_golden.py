"""CPU: the oracle restatement reproduces the golden vectors that were produced by
the verbatim reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import gotennet_oracle as orc
from oracle.golden_cases import CASES, NORM_CASES, blob, grad_fingerprint

TOL = 1e-4  # north_star: within 1e-4 relative, fp32


def rel(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.parametrize("name", list(CASES) + list(NORM_CASES))
def test_oracle_matches_reference_golden(name, golden_dir):
    spec = {**CASES, **NORM_CASES}[name]
    cfg = spec["cfg"]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    z, pos, batch = blob(spec["atoms"], spec["seed"])
    assert np.array_equal(z.numpy(), gold["z"]) and np.array_equal(pos.numpy(), gold["pos"])
    sd = orc.make_state_dict(cfg, seed=spec["seed"])
    sd = {k: v.clone().requires_grad_("radial_basis" not in k) for k, v in sd.items()}
    pos = pos.clone().requires_grad_(True)
    inter = {}
    h, X = orc.wrapper_forward(sd, cfg, z, pos, batch, inter)
    # integer work: bit exact
    assert np.array_equal(inter["edge_index"].numpy(), gold["edge_index"])
    assert rel(inter["edge_weight"].detach(), gold["edge_weight"]) < 1e-6
    assert rel(h.detach(), gold["h"]) < TOL and rel(X.detach(), gold["X"]) < TOL
    for i in range(cfg.n_interactions):
        for s in ("h", "X", "t"):
            assert rel(inter[f"{s}{i + 1}"].detach(), gold[f"state_{s}{i + 1}"]) < TOL, (s, i)
    (h.sum() + X.pow(2).sum()).backward()
    assert rel(pos.grad, gold["grad_pos"]) < TOL
    n_checked = 0
    for k in gold.files:
        if k.startswith("grad_") and k != "grad_pos":
            g = sd[k[5:]].grad
            g = torch.zeros_like(sd[k[5:]]) if g is None else g
            assert rel(grad_fingerprint(g), gold[k]) < TOL, k
            n_checked += 1
    assert n_checked == len([k for k, _, _ in orc.state_dict_spec(cfg) if "tensor_layernorm" not in k])  # (a buffer)


def test_fp64_tiebreak(golden_dir):
    """fp32 oracle sits within ~1e-6 of its own fp64 evaluation (noise floor)."""
    spec = CASES["yaml_l2"]
    cfg = spec["cfg"]
    z, pos, batch = blob(spec["atoms"], spec["seed"])
    sd = orc.make_state_dict(cfg, seed=spec["seed"])
    h32, X32 = orc.wrapper_forward(sd, cfg, z, pos, batch)
    sd64 = {k: v.double() for k, v in sd.items()}
    h64, X64 = orc.wrapper_forward(sd64, cfg, z, pos.double(), batch)
    assert rel(h32, h64) < 2e-5 and rel(X32, X64) < 2e-5


def test_radius_graph_properties():
    z, pos, batch = orc.synth_batch("qm9", 16, seed=7)
    ei = orc.radius_graph(pos, batch, 5.0, 32)
    src, tgt = ei
    assert (batch[src] == batch[tgt]).all()
    # sorted by (target, source), self loops present for every node
    key = tgt * pos.size(0) + src
    assert (key[1:] > key[:-1]).all()
    assert ((src == tgt).sum() == pos.size(0))
    # truncation keeps the first K sources
    ei8 = orc.radius_graph(pos, batch, 5.0, 4)
    deg = torch.bincount(ei8[1], minlength=pos.size(0))
    assert deg.max() <= 4
    # empty input
    assert orc.radius_graph(pos[:0], batch[:0], 5.0).shape == (2, 0)


def test_rotation_equivariance_lmax2():
    """h invariant, l=1 block of X rotates as a vector (valid for lmax<=2)."""
    spec = CASES["yaml_l2"]
    cfg = spec["cfg"]
    z, pos, batch = blob(spec["atoms"], spec["seed"])
    sd = {k: v.double() for k, v in orc.make_state_dict(cfg, seed=5).items()}
    q, _ = torch.linalg.qr(torch.randn(3, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(0)))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    h1, X1 = orc.wrapper_forward(sd, cfg, z, pos.double(), batch)
    h2, X2 = orc.wrapper_forward(sd, cfg, z, pos.double() @ q.T, batch)
    assert rel(h2, h1) < 1e-9
    assert rel(X2[:, :3], torch.einsum("ab,nbc->nac", q, X1[:, :3])) < 1e-9
    assert rel(X2[:, 3:].pow(2).sum(1), X1[:, 3:].pow(2).sum(1)) < 1e-9


# ------------------------------------------------ read-out head --
from oracle.golden_cases import HEAD_CASES, probe_vector  # noqa: E402


@pytest.mark.parametrize("name", list(HEAD_CASES))
def test_oracle_head_matches_reference_golden(name, golden_dir):
    """Energy, forces (-dE/dpos), per-atom contributions and all parameter gradients of the oracle's
    Atomwise restatement against the verbatim reference head (tests/golden/make_golden_head.py)."""
    spec = HEAD_CASES[name]
    cfg = spec["cfg"]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    z, pos, batch = blob(spec["atoms"], spec["seed"])
    n_mol = len(spec["atoms"])
    sd = {k: v.clone().requires_grad_("radial_basis" not in k) for k, v in orc.make_state_dict(cfg, seed=spec["seed"]).items()}
    sdh = {k: v.clone().requires_grad_(k.startswith("out_net")) for k, v in
           orc.make_head_state_dict(cfg.n_atom_basis, seed=spec["seed"]).items()}
    E, Fo, h = orc.energy_and_forces(sd, sdh, cfg, z, pos, batch, n_mol, spec["activation"])
    assert rel(E.detach(), gold["energy"]) < TOL and rel(Fo.detach(), gold["forces"]) < TOL
    _, yi = orc.atomwise_forward(sdh, h, z, batch, n_mol, spec["activation"])
    assert rel(yi.detach(), gold["contrib"]) < TOL
    ((E * probe_vector(n_mol).unsqueeze(1)).sum()).backward()
    n = 0
    for k in gold.files:
        if k.startswith("gradh_"):
            assert rel(grad_fingerprint(sdh[k[6:]].grad), gold[k]) < TOL, k
            n += 1
        elif k.startswith("grad_"):
            g = sd[k[5:]].grad
            assert rel(grad_fingerprint(g if g is not None else torch.zeros_like(sd[k[5:]])), gold[k]) < TOL, k
            n += 1
    assert n == 4 + len(orc.state_dict_spec(cfg))


def test_oracle_head_modes():
    """mean aggregation = sum / atom count; None returns the per-atom values (outputs.py:354-357)."""
    g = torch.Generator().manual_seed(0)
    h = torch.randn(11, 16, generator=g)
    z = torch.randint(1, 9, (11,), generator=g)
    batch = torch.tensor([0] * 4 + [1] * 1 + [2] * 6)
    sdh = orc.make_head_state_dict(16, seed=1)
    ys, yi = orc.atomwise_forward(sdh, h, z, batch, 3, "ssp", "sum")
    ym, _ = orc.atomwise_forward(sdh, h, z, batch, 3, "ssp", "mean")
    yn, _ = orc.atomwise_forward(sdh, h, z, batch, 3, "ssp", None)
    assert torch.allclose(ym, ys / torch.tensor([[4.0], [1.0], [6.0]]), atol=1e-6) and torch.equal(yn, yi)
    assert torch.allclose(ys[1], yi[4], atol=1e-6)


def test_oracle_dropout_matches_reference_golden(golden_dir):
    """Training-mode attention dropout (gotennet.py:513): the oracle fed with the masks the reference's own F.dropout
    drew reproduces the reference's outputs and gradients."""
    from oracle.golden_cases import DROPOUT_CASES
    for name, spec in DROPOUT_CASES.items():
        gold = np.load(os.path.join(golden_dir, name + ".npz"))
        cfg, p = spec["cfg"], spec["p"]
        z, pos, batch = blob(spec["atoms"], spec["seed"])
        sd = {k: v.clone().requires_grad_(v.is_floating_point() and "radial_basis" not in k)
              for k, v in orc.make_state_dict(cfg, seed=spec["seed"]).items()}
        drop = [torch.from_numpy(m).float() / (1.0 - p) for m in gold["masks"]]
        pos_o = pos.clone().requires_grad_(True)
        h, X = orc.wrapper_forward(sd, cfg, z, pos_o, batch, drop_masks=drop)
        (h.sum() + X.pow(2).sum()).backward()
        assert rel(h, torch.from_numpy(gold["h"])) < 1e-6 and rel(X, torch.from_numpy(gold["X"])) < 1e-6
        assert rel(pos_o.grad, torch.from_numpy(gold["grad_pos"])) < 1e-5
        for k in gold.files:
            if k.startswith("grad_") and k != "grad_pos":
                g = sd[k[5:]].grad
                g = g if g is not None else torch.zeros_like(sd[k[5:]])
                assert rel(grad_fingerprint(g), torch.from_numpy(gold[k])) < 1e-4, k


# ------------------------------------------------ remaining QM9 heads --
from oracle.golden_cases import HEAD2_CASES, head2_loss, head2_state  # noqa: E402


@pytest.mark.parametrize("name", list(HEAD2_CASES))
def test_oracle_heads2_match_reference_golden(name, golden_dir):
    """Dipole / ElectronicSpatialExtentV2 restatements against the verbatim reference's outputs and gradients."""
    from gotennet_b200.atomic_data import ATOMIC_MASSES
    spec = HEAD2_CASES[name]
    cfg, kind = spec["cfg"], spec["kind"]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    z, pos, batch = blob(spec["atoms"], spec["seed"])
    n_mol = len(spec["atoms"])
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "radial_basis" not in k)
          for k, v in orc.make_state_dict(cfg, seed=spec["seed"]).items()}
    sdh = {k: v.clone().requires_grad_(k.startswith(("out_net", "equivariant"))) for k, v in head2_state(spec).items()}
    pos = pos.clone().requires_grad_(True)
    h, X = orc.wrapper_forward(sd, cfg, z, pos, batch)
    yv = None
    if kind == "dipole":
        y, yv = orc.dipole_forward(sdh, h, X, pos, batch, n_mol, mean=spec["mean"], stddev=spec["stddev"],
                                   predict_magnitude=spec["predict_magnitude"])
        assert rel(yv.detach(), gold["y_vector"]) < TOL
    else:
        y, x = orc.spatial_extent_forward(sdh, h, z, pos, batch, n_mol, torch.tensor(ATOMIC_MASSES))
        assert rel(x.detach(), gold["contrib"]) < TOL
    assert rel(y.detach(), gold["y"]) < TOL
    head2_loss(kind, y, yv, n_mol).backward()
    assert rel(pos.grad, gold["grad_pos"]) < TOL
    for k in gold.files:
        if k.startswith("gradh_"):
            assert rel(grad_fingerprint(sdh[k[6:]].grad), gold[k]) < TOL, k
        elif k.startswith("grad_") and k != "grad_pos":
            g = sd[k[5:]].grad
            assert rel(grad_fingerprint(g if g is not None else torch.zeros_like(sd[k[5:]])), gold[k]) < TOL, k
