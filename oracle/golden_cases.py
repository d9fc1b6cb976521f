"""Golden-vector case table shared by tests/golden/make_golden.py (which runs the
verbatim reference) and the parity tests (which rebuild the identical inputs and
weights from the seeds).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import torch

from .gotennet_oracle import OracleConfig

CASES = {
    # BASELINE.json configs[0]: class defaults, C=64, 2 interactions, lmax=1, one 20-atom molecule
    "cfg1": dict(cfg=OracleConfig(n_atom_basis=64, n_interactions=2, lmax=1), atoms=[20], seed=1),
    # configs/model/gotennet.yaml flags (lmax=2, sep_* on, scale_edge off) at reduced width,
    # ragged batch incl. a single-atom and a two-atom molecule
    "yaml_l2": dict(cfg=OracleConfig(n_atom_basis=64, n_interactions=3, lmax=2, sep_dir=True,
                                     sep_tensor=True, scale_edge=False),
                    atoms=[17, 1, 23, 2, 9], seed=2),
    # lmax=3 with neighbour truncation (asymmetric graph)
    "l3_trunc": dict(cfg=OracleConfig(n_atom_basis=32, n_interactions=2, lmax=3, num_heads=4, sep_dir=True,
                                      sep_tensor=True, scale_edge=True, max_num_neighbors=8),
                     atoms=[40, 5], seed=3),
    # all sep_* off at lmax=2 (shared W_vk, single direction / tensor chunk)
    "nosep_l2": dict(cfg=OracleConfig(n_atom_basis=32, n_interactions=2, lmax=2, sep_htr=False,
                                      sep_dir=False, sep_tensor=False, scale_edge=True),
                     atoms=[12, 14], seed=4),
}

# training-mode attention dropout (configs/model/gotennet.yaml:30 ships attn_dropout 0.1): the generator records the
# masks the reference's own F.dropout drew (gotennet.py:513) and stores them with the outputs
DROPOUT_CASES = {
    "dropout_l2": dict(cfg=OracleConfig(n_atom_basis=64, n_interactions=3, lmax=2, sep_dir=True, sep_tensor=True,
                                        scale_edge=False), atoms=[15, 3, 19], seed=7, p=0.25),
}

# optional pre-norms of the GATA block: nn.LayerNorm on h + TensorLayerNorm on X
NORM_CASES = {
    "norms_l2": dict(cfg=OracleConfig(n_atom_basis=64, n_interactions=3, lmax=2, sep_dir=True, sep_tensor=True,
                                      scale_edge=False, layernorm="layer", steerable_norm="tensor"),
                     atoms=[14, 2, 21, 1], seed=8),
    "norms_l3": dict(cfg=OracleConfig(n_atom_basis=32, n_interactions=2, lmax=3, num_heads=4, scale_edge=True,
                                      steerable_norm="tensor"), atoms=[11, 9], seed=9),
}

# the other registered radial bases (layers.py:749-777); merged into NORM_CASES' test parametrisation
NORM_CASES.update({
    "bessel_l1": dict(cfg=OracleConfig(n_atom_basis=32, n_interactions=2, lmax=1, num_heads=4, n_rbf=20,
                                       radial_basis="BesselBasis"), atoms=[13, 6], seed=10),
    "gauss_l2": dict(cfg=OracleConfig(n_atom_basis=32, n_interactions=2, lmax=2, num_heads=4, n_rbf=24, sep_dir=True,
                                      sep_tensor=True, scale_edge=False, radial_basis="GaussianRBF"),
                     atoms=[10, 12], seed=11),
})

# edge_updates="norej": HTR without the vector rejection (gotennet.py:176-177, :583-599)
NORM_CASES["norej_l2"] = dict(cfg=OracleConfig(n_atom_basis=32, n_interactions=2, lmax=2, num_heads=4, sep_dir=True,
                                               sep_tensor=True, scale_edge=False, edge_updates="norej"),
                              atoms=[12, 8], seed=12)

# gated edge updates: gamma_w = Sigmoid ("gated"), Tanh ("gatedt"), SiLU ("act") on the scalar HTR weight (:283-289)
for _i, _flag in enumerate(["gated", "gatedt_norej", "act"]):
    NORM_CASES["eu_" + _flag] = dict(cfg=OracleConfig(n_atom_basis=32, n_interactions=2, lmax=2 if _i != 1 else 1,
                                                      num_heads=4, sep_dir=True, sep_tensor=True, scale_edge=bool(_i % 2),
                                                      edge_updates=_flag), atoms=[11, 7], seed=13 + _i)

# two-layer gamma_t: "mlp" (no final activation, emlp_dim 48) and "mlpa" combined with a gate
NORM_CASES["eu_mlp"] = dict(cfg=OracleConfig(n_atom_basis=32, n_interactions=2, lmax=2, num_heads=4, sep_dir=True,
                                             sep_tensor=True, scale_edge=False, edge_updates="mlp", emlp_dim=48),
                            atoms=[10, 9], seed=16)
NORM_CASES["eu_mlpa_gated"] = dict(cfg=OracleConfig(n_atom_basis=32, n_interactions=3, lmax=1, num_heads=4,
                                                    edge_updates="mlpa_gated"), atoms=[13, 4], seed=17)

# gamma_w as a network on the HTR weight ("linw" / "linwa" with "ln" / "postln", evec_dim) and LayerNorm inside a
# two-layer gamma_t (edge_ln) - reference gotennet.py:178-185, :236, :249, :270-282.  "linwa" needs a Module
# activation in the reference (a string one; its functional default fails in nn.Sequential): `activation` below is
# what the generator passes to the reference constructor.
NORM_CASES["eu_linw_ln_ev"] = dict(cfg=OracleConfig(n_atom_basis=32, n_interactions=3, lmax=2, num_heads=4, sep_dir=True,
                                                    sep_tensor=True, scale_edge=False, edge_updates="linw_ln",
                                                    evec_dim=24), atoms=[12, 7], seed=18)
NORM_CASES["eu_linwa_postln_gated"] = dict(cfg=OracleConfig(n_atom_basis=32, n_interactions=2, lmax=1, num_heads=4,
                                                            edge_updates="linwa_postln_gated_norej", evec_dim=48),
                                           atoms=[9, 11], seed=19, activation="silu")
NORM_CASES["eu_linw_nosep"] = dict(cfg=OracleConfig(n_atom_basis=32, n_interactions=2, lmax=2, num_heads=4,
                                                    sep_htr=False, edge_updates="linw_mlp", emlp_dim=40),
                                   atoms=[13, 5], seed=20)
NORM_CASES["eu_mlpa_edgeln"] = dict(cfg=OracleConfig(n_atom_basis=32, n_interactions=2, lmax=2, num_heads=4, sep_dir=True,
                                                     sep_tensor=True, scale_edge=False, edge_updates="mlpa_gatedt",
                                                     emlp_dim=48, edge_ln="layer"), atoms=[10, 8], seed=22)

# read-out head cases: representation + Atomwise energy head with forces.  `rep` names the
# representation config; head = Atomwise(n_in=C, activation=..., mean, stddev, atomref, derivative="forces")
HEAD_CASES = {
    # BASELINE configs[2] flags at reduced size: aspirin-shape molecules (21 atoms), lmax=2, yaml flags, SiLU head
    "head_forces_l2": dict(cfg=OracleConfig(n_atom_basis=64, n_interactions=3, lmax=2, sep_dir=True, sep_tensor=True,
                                            scale_edge=False), atoms=[21, 21, 7], seed=5, activation="silu"),
    # class defaults (lmax=1), shifted-softplus head (the reference's Atomwise default activation)
    "head_forces_ssp": dict(cfg=OracleConfig(n_atom_basis=32, n_interactions=2, lmax=1), atoms=[9, 12], seed=6,
                            activation="ssp"),
}


# remaining QM9 heads: Dipole (two gated equivariant blocks, magnitude output, standardised charges as the
# QM9 task builds it, QM9Task.py:172-179) and ElectronicSpatialExtentV2 (QM9Task.py:181-185)
HEAD2_CASES = {
    "dipole_l2": dict(kind="dipole", cfg=OracleConfig(n_atom_basis=64, n_interactions=2, lmax=2, sep_dir=True,
                                                      sep_tensor=True, scale_edge=False), atoms=[14, 1, 9, 20], seed=31,
                      mean=0.3, stddev=1.7, predict_magnitude=True),
    "dipole_vec_l1": dict(kind="dipole", cfg=OracleConfig(n_atom_basis=32, n_interactions=2, lmax=1, num_heads=4),
                          atoms=[11, 6], seed=32, mean=None, stddev=None, predict_magnitude=False),
    "ese_l2": dict(kind="ese", cfg=OracleConfig(n_atom_basis=64, n_interactions=2, lmax=2, sep_dir=True, sep_tensor=True,
                                                scale_edge=False), atoms=[16, 2, 12], seed=33),
}

def blob(n_atoms, seed, sigma0=1.45):
    """Gaussian-blob molecules of the given sizes: z [N] i64, pos [N,3] f32, batch [N] i64."""
    g = torch.Generator().manual_seed(seed)
    sizes = torch.tensor(n_atoms)
    batch = torch.repeat_interleave(torch.arange(len(n_atoms)), sizes)
    sigma = sigma0 * (sizes.double() / 18.0).pow(1 / 3).float()
    pos = torch.randn(int(sizes.sum()), 3, generator=g) * sigma[batch].unsqueeze(-1)
    species = torch.tensor([1, 6, 7, 8, 9])
    z = species[torch.randint(0, 5, (int(sizes.sum()),), generator=g)]
    return z, pos, batch


def probe_vector(n):
    g = torch.Generator().manual_seed(12345 + n)
    return torch.randn(n, generator=g)


def grad_fingerprint(g):
    """1-D gradients are kept whole; matrices as a seeded random projection of every row."""
    if g.dim() == 1:
        return g
    return g @ probe_vector(g.shape[1]).to(g.dtype).to(g.device)


def head2_loss(kind, y, yv, n_mol):
    """Probe loss of the HEAD2 cases: seeded weights on the head output (+ the vector output for Dipole)."""
    w = probe_vector(n_mol * y.shape[1]).view(n_mol, y.shape[1]).to(y)
    loss = (y * w).sum()
    if kind == "dipole":
        loss = loss + 0.5 * (yv.squeeze(-1) * probe_vector(n_mol * 3 + 1)[:n_mol * 3].view(n_mol, 3).to(y)).sum()
    return loss


def head2_state(spec):
    from . import gotennet_oracle as orc
    C = spec["cfg"].n_atom_basis
    if spec["kind"] == "dipole":
        return orc.make_dipole_state_dict(C, seed=spec["seed"])
    return orc.make_head_state_dict(C, seed=spec["seed"], atomref=False)
