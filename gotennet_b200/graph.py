"""Graph plan: the target-sorted edge list plus its CSR / transposed views.

Replaces torch_cluster.radius_graph + the implicit PyG gather/scatter indexing of
the reference (components/layers.py:1588-1590, representation/gotennet.py:412-424).
The plan is what every kernel of the path indexes with; it carries no floats.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from ._lib import GotenError, lib
from .ops import _ptr, _stream


@dataclass
class GraphPlan:
    N: int
    E: int
    n_mol: int
    src: torch.Tensor        # [E] int32, sorted by (tgt, src)
    tgt: torch.Tensor        # [E] int32
    edge_index: torch.Tensor  # [2,E] int64 (row0 = source, row1 = target)
    tgt_ptr: torch.Tensor    # [N+1] int32
    src_ptr: torch.Tensor    # [N+1] int32
    src_perm: torch.Tensor   # [E] int32 edge ids grouped by source
    deg_out: torch.Tensor    # [N] int32
    max_deg_in: int
    order: Optional[torch.Tensor] = None  # permutation applied to a caller-supplied edge list (None = identity)


def radius_graph_plan(pos: torch.Tensor, batch: Optional[torch.Tensor], cutoff: float, max_num_neighbors: int = 32,
                      loop: bool = True) -> GraphPlan:
    """torch_cluster.radius_graph semantics (CUDA build): strict d^2 < r^2, same molecule,
    first-K sources in ascending index per target, edges grouped by target."""
    if not pos.is_cuda:
        raise GotenError("radius_graph_plan needs CUDA tensors (there is no CPU path)")
    pos = pos.detach().contiguous().float()
    N = pos.shape[0]
    dev = pos.device
    if batch is None:
        batch = torch.zeros(N, dtype=torch.long, device=dev)
    batch = batch.contiguous().long()
    L = lib()
    st = _stream()
    import ctypes

    mol_ptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    mol_of = torch.empty(max(N, 1), dtype=torch.int32, device=dev)
    tgt_ptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    scratch = torch.empty(2 * N + 4096, dtype=torch.int32, device=dev)
    nE, nM = ctypes.c_int64(0), ctypes.c_int32(0)
    L.call("goten_radius_graph_count", _ptr(pos), _ptr(batch), N, float(cutoff), int(max_num_neighbors), int(loop),
           _ptr(mol_ptr), _ptr(mol_of), _ptr(tgt_ptr), _ptr(scratch), ctypes.addressof(nE), ctypes.addressof(nM), st)
    E = int(nE.value)
    src = torch.empty(E, dtype=torch.int32, device=dev)
    tgt = torch.empty(E, dtype=torch.int32, device=dev)
    edge_index = torch.empty(2, E, dtype=torch.int64, device=dev)
    deg_out = torch.empty(max(N, 1), dtype=torch.int32, device=dev)
    src_ptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    src_perm = torch.empty(E, dtype=torch.int32, device=dev)
    L.call("goten_radius_graph_fill", _ptr(pos), _ptr(mol_ptr), _ptr(mol_of), _ptr(tgt_ptr), N, E, float(cutoff),
           int(max_num_neighbors), int(loop), _ptr(src), _ptr(tgt), _ptr(edge_index), _ptr(deg_out), _ptr(src_ptr),
           _ptr(src_perm), _ptr(scratch), st)
    return GraphPlan(N=N, E=E, n_mol=int(nM.value), src=src, tgt=tgt, edge_index=edge_index, tgt_ptr=tgt_ptr,
                     src_ptr=src_ptr, src_perm=src_perm, deg_out=deg_out, max_deg_in=int(max_num_neighbors))


def plan_from_edge_index(edge_index: torch.Tensor, num_nodes: int) -> GraphPlan:
    """Plan for a caller-supplied edge list (GotenNet.forward / GATA.forward signature).  Index
    preparation (stable sorts) is plumbing and uses torch; `order` records the permutation that
    brings the caller's per-edge tensors into plan order."""
    if not edge_index.is_cuda:
        raise GotenError("plan_from_edge_index needs CUDA tensors (there is no CPU path)")
    dev = edge_index.device
    E = edge_index.shape[1]
    src64, tgt64 = edge_index[0], edge_index[1]
    order = None
    if E > 1 and not bool((tgt64[1:] >= tgt64[:-1]).all()):
        order = torch.sort(tgt64, stable=True).indices
        src64, tgt64 = src64[order], tgt64[order]
    src = src64.to(torch.int32).contiguous()
    tgt = tgt64.to(torch.int32).contiguous()
    by_src = torch.sort(src64, stable=True).indices.to(torch.int32).contiguous()
    N = int(num_nodes)
    tgt_ptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    src_ptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    deg_out = torch.empty(max(N, 1), dtype=torch.int32, device=dev)
    src_perm = torch.empty(E, dtype=torch.int32, device=dev)
    scratch = torch.empty(N + 4096 + 2, dtype=torch.int32, device=dev)
    lib().call("goten_csr_from_sorted", _ptr(src), _ptr(tgt), _ptr(by_src), N, E, _ptr(tgt_ptr), _ptr(deg_out),
               _ptr(src_ptr), _ptr(src_perm), _ptr(scratch), _stream())
    max_deg = int((tgt_ptr[1:] - tgt_ptr[:-1]).max().item()) if N > 0 and E > 0 else 1
    ei = torch.stack([src64, tgt64], 0).contiguous() if order is not None else edge_index.contiguous()
    return GraphPlan(N=N, E=E, n_mol=0, src=src, tgt=tgt, edge_index=ei, tgt_ptr=tgt_ptr, src_ptr=src_ptr,
                     src_perm=src_perm, deg_out=deg_out, max_deg_in=max(max_deg, 1), order=order)
