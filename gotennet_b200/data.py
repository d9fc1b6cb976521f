"""Batch ingestion for the hot path (first half): variable-size molecules -> the
`z / pos / batch` triple `GotenNetWrapper.forward` reads (reference representation/gotennet.py:1043;
PyG `Batch.from_data_list` collation in the reference's datamodule, data/datamodule.py:181-219), plus
the molecule pointer array and `num_graphs` our read-out kernels want (so no host read of
`batch[-1]` is needed, cf. outputs.py:355).  One pinned host staging buffer per field and one
asynchronous H2D copy each."""
from __future__ import annotations

from typing import Iterable, Optional, Sequence, Tuple

import torch


class MoleculeBatch:
    """Attribute bag with the PyG `Batch` fields the reference model touches."""

    def __init__(self, z, pos, batch, ptr, num_graphs):
        self.z, self.pos, self.batch, self.ptr, self.num_graphs = z, pos, batch, ptr, num_graphs

    def to(self, device, non_blocking: bool = True) -> "MoleculeBatch":
        mv = lambda t: t.to(device, non_blocking=non_blocking)  # noqa: E731
        return MoleculeBatch(mv(self.z), mv(self.pos), mv(self.batch), mv(self.ptr), self.num_graphs)

    def __getitem__(self, key):  # PyG batches allow item access too
        return getattr(self, key)


def collate(molecules: Iterable[Tuple[Sequence[int], Sequence[Sequence[float]]]], device: Optional[torch.device] = None,
            pin: bool = True) -> MoleculeBatch:
    """[(atomic_numbers [n_i], positions [n_i, 3]), ...] -> MoleculeBatch (int64 z / batch / ptr, float32 pos).
    With `device` given the fields are staged in pinned memory and copied asynchronously."""
    zs, ps = [], []
    for z, pos in molecules:
        z = torch.as_tensor(z, dtype=torch.int64).reshape(-1)
        pos = torch.as_tensor(pos, dtype=torch.float32).reshape(-1, 3)
        if z.numel() != pos.shape[0]:
            raise ValueError(f"molecule with {z.numel()} atomic numbers but {pos.shape[0]} positions")
        zs.append(z)
        ps.append(pos)
    counts = torch.tensor([z.numel() for z in zs], dtype=torch.int64)
    ptr = torch.zeros(len(zs) + 1, dtype=torch.int64)
    torch.cumsum(counts, 0, out=ptr[1:])
    z = torch.cat(zs) if zs else torch.zeros(0, dtype=torch.int64)
    pos = torch.cat(ps) if ps else torch.zeros(0, 3)
    batch = torch.repeat_interleave(torch.arange(len(zs)), counts)
    out = MoleculeBatch(z, pos, batch, ptr, len(zs))
    if device is not None:
        if pin and torch.cuda.is_available():
            out = MoleculeBatch(z.pin_memory(), pos.pin_memory(), batch.pin_memory(), ptr.pin_memory(), len(zs))
        out = out.to(device)
    return out
