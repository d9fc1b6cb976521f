"""Deterministic synthetic molecule batches of the shapes BASELINE.json names:
QM9-shape (n ~ round(N(18, 2.94)) clamped 3..29, Gaussian blob sigma = 1.45 (n/18)^(1/3) A),
rMD17-aspirin-shape (21 atoms, sigma0 1.6) and MD22-shape (370 atoms, sigma0 1.5)."""
from __future__ import annotations

import torch

_SHAPES = {"qm9": (None, 1.45), "aspirin": (21, 1.6), "md22": (370, 1.5)}


def synth_batch(kind: str, n_mol: int, seed: int = 0):
    """Returns z [N] int64, pos [N,3] float32, batch [N] int64 (CPU tensors)."""
    fixed, sigma0 = _SHAPES[kind]
    g = torch.Generator().manual_seed(seed)
    if fixed is None:
        n = (18.0 + 2.94 * torch.randn(n_mol, generator=g)).round().clamp(3, 29).long()
    else:
        n = torch.full((n_mol,), fixed, dtype=torch.long)
    N = int(n.sum())
    batch = torch.repeat_interleave(torch.arange(n_mol), n)
    sigma = sigma0 * (n.double() / 18.0).pow(1.0 / 3.0).float()
    pos = torch.randn(N, 3, generator=g) * sigma[batch].unsqueeze(-1)
    species = torch.tensor([1, 6, 7, 8, 9])
    probs = torch.tensor([0.51, 0.35, 0.06, 0.078, 0.002])
    z = species[torch.multinomial(probs, N, replacement=True, generator=g)]
    return z, pos, batch
