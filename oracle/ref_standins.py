"""Stand-in modules that let the UNMODIFIED reference import in this container.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (gotennet_b200/) may
import this file.  It exists so that `tests/golden/make_golden.py` and
`tests/test_oracle_vs_reference.py` can execute the verbatim reference code
under /root/reference (which is absent on the GPU box) to pin the oracle.

The reference imports four third-party packages that are not installed here
(no network): torch_geometric, torch_cluster, pytorch_lightning, omegaconf.
Only a handful of their symbols are touched on the hot path:

  torch_geometric.nn.MessagePassing        gotennet.py:11, layers.py:16
  torch_geometric.utils.scatter / softmax  gotennet.py:13
  torch_geometric.typing.OptTensor         gotennet.py:12
  torch_geometric.nn.inits.glorot_orthogonal  layers.py:17
  torch_cluster.radius_graph               layers.py:15
  pytorch_lightning.utilities.rank_zero_only  utils/__init__.py:6
  omegaconf.DictConfig / OmegaConf         utils/__init__.py:5
  torch_scatter.scatter, ase               components/outputs.py:3,6 (Atomwise read-out head)

The stand-ins below restate the *published* semantics of those symbols
(PyG 2.x, torch_cluster 1.6 CUDA build):
  * MessagePassing, flow source_to_target: `foo_j = foo[edge_index[0]]`,
    `foo_i = foo[edge_index[1]]`, aggregation index = edge_index[1],
    dim_size = number of nodes.
  * utils.softmax: subtract (detached) per-segment max, exp, divide by
    (segment sum + 1e-16).
  * utils.scatter(reduce='sum'): zero-initialised index_add.
  * radius_graph(loop=True): strict d^2 < r^2, same batch id only, at most
    K neighbours per target kept in ascending source index (CUDA build),
    edges emitted grouped by target.
"""
from __future__ import annotations

import inspect
import sys
import types
from typing import Optional

import torch


def _scatter(src, index, dim=0, dim_size=None, reduce="sum"):
    if dim < 0:
        dim += src.dim()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    if reduce in ("sum", "add"):
        out = src.new_zeros(shape)
        return out.index_add_(dim, index, src)
    if reduce == "mean":
        out = src.new_zeros(shape).index_add_(dim, index, src)
        cnt = src.new_zeros(dim_size).index_add_(0, index, src.new_ones(index.numel()))
        view = [1] * src.dim()
        view[dim] = dim_size
        return out / cnt.clamp(min=1).view(view)
    if reduce in ("max", "min"):
        view = [1] * src.dim()
        view[dim] = index.numel()
        idx = index.view(view).expand_as(src)
        out = src.new_zeros(shape)
        return out.scatter_reduce_(dim, idx, src, "a" + reduce, include_self=False)
    raise ValueError(reduce)


def _segment_softmax(src, index, ptr=None, num_nodes=None, dim=0):
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    view = [1] * src.dim()
    view[dim] = index.numel()
    idx = index.view(view).expand_as(src)
    shape = list(src.shape)
    shape[dim] = n
    mx = src.new_full(shape, float("-inf")).scatter_reduce_(
        dim, idx, src.detach(), "amax", include_self=True
    )
    out = (src - mx.index_select(dim, index)).exp()
    den = src.new_zeros(shape).index_add_(dim, index, out) + 1e-16
    return out / den.index_select(dim, index)


class _MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2, **kw):
        super().__init__()
        self.aggr = aggr
        self.flow = flow
        self.node_dim = node_dim

    # -- helpers -----------------------------------------------------------
    def _collect(self, fn, edge_index, size, kwargs):
        j, i = (0, 1) if self.flow == "source_to_target" else (1, 0)
        params = [p for p in inspect.signature(fn).parameters]
        out = {}
        for name in params:
            if name in ("index", "ptr", "dim_size", "size", "edge_index"):
                continue
            if name.endswith("_i") or name.endswith("_j"):
                base = kwargs[name[:-2]]
                sel = edge_index[i] if name.endswith("_i") else edge_index[j]
                nd = self.node_dim if self.node_dim >= 0 else base.dim() + self.node_dim
                out[name] = base.index_select(nd, sel)
            else:
                out[name] = kwargs[name]
        if "edge_index" in params:
            out["edge_index"] = edge_index
        if "index" in params:
            out["index"] = edge_index[i]
        if "ptr" in params:
            out["ptr"] = None
        if "dim_size" in params:
            out["dim_size"] = size
        return out

    def _num_nodes(self, kwargs, edge_index):
        nd = self.node_dim
        for k, v in kwargs.items():
            if torch.is_tensor(v) and v.dim() > 0 and k not in ("edge_index",):
                # first node-level tensor decides; edge-level tensors are passed
                # without suffix and are never gathered.
                pass
        return None

    def propagate(self, edge_index, size=None, **kwargs):
        j, i = (0, 1) if self.flow == "source_to_target" else (1, 0)
        # dim_size: the node count, taken from any gathered (suffix) argument
        n = None
        for name in inspect.signature(self.message).parameters:
            if name.endswith("_i") or name.endswith("_j"):
                base = kwargs[name[:-2]]
                nd = self.node_dim if self.node_dim >= 0 else base.dim() + self.node_dim
                n = base.size(nd)
                break
        msg_kwargs = self._collect(self.message, edge_index, n, kwargs)
        out = self.message(**msg_kwargs)
        agg_params = inspect.signature(self.aggregate).parameters
        agg_kwargs = {}
        if "index" in agg_params:
            agg_kwargs["index"] = edge_index[i]
        if "ptr" in agg_params:
            agg_kwargs["ptr"] = None
        if "dim_size" in agg_params:
            agg_kwargs["dim_size"] = n
        out = self.aggregate(out, **agg_kwargs)
        return self.update(out)

    def edge_updater(self, edge_index, size=None, **kwargs):
        e_kwargs = self._collect(self.edge_update, edge_index, None, kwargs)
        return self.edge_update(**e_kwargs)

    def aggregate(self, inputs, index, ptr=None, dim_size=None):
        return _scatter(inputs, index, dim=self.node_dim, dim_size=dim_size, reduce=self.aggr)

    def message(self, x_j):
        return x_j

    def update(self, inputs):
        return inputs


def _radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow="source_to_target", **kw):
    """Dense restatement of torch_cluster.radius_graph (CUDA build semantics)."""
    assert flow == "source_to_target"
    n = x.size(0)
    if batch is None:
        batch = x.new_zeros(n, dtype=torch.long)
    diff = x.unsqueeze(1) - x.unsqueeze(0)  # [target, source, 3]
    d2 = (diff * diff).sum(-1)
    ok = (d2 < r * r) & (batch.unsqueeze(1) == batch.unsqueeze(0))
    if not loop:
        ok = ok & ~torch.eye(n, dtype=torch.bool, device=x.device)
    # first-K by ascending source index per target
    rank = ok.long().cumsum(1)
    ok = ok & (rank <= max_num_neighbors)
    tgt, src = ok.nonzero(as_tuple=True)  # row-major -> sorted by target then source
    return torch.stack([src, tgt], dim=0)


def _glorot_orthogonal(tensor, scale):
    torch.nn.init.orthogonal_(tensor.data)
    s = scale / ((tensor.size(-2) + tensor.size(-1)) * tensor.var())
    tensor.data *= s.sqrt()


def install() -> None:
    """Register the stand-ins in sys.modules (idempotent)."""
    if "torch_geometric" in sys.modules and getattr(sys.modules["torch_geometric"], "_goten_standin", False):
        return

    tg = types.ModuleType("torch_geometric")
    tg._goten_standin = True
    tg_nn = types.ModuleType("torch_geometric.nn")
    tg_nn.MessagePassing = _MessagePassing
    tg_inits = types.ModuleType("torch_geometric.nn.inits")
    tg_inits.glorot_orthogonal = _glorot_orthogonal
    tg_typing = types.ModuleType("torch_geometric.typing")
    tg_typing.OptTensor = Optional[torch.Tensor]
    tg_utils = types.ModuleType("torch_geometric.utils")
    tg_utils.scatter = _scatter
    tg_utils.softmax = _segment_softmax
    tg.nn, tg.typing, tg.utils = tg_nn, tg_typing, tg_utils
    tg_nn.inits = tg_inits
    sys.modules.update({
        "torch_geometric": tg,
        "torch_geometric.nn": tg_nn,
        "torch_geometric.nn.inits": tg_inits,
        "torch_geometric.typing": tg_typing,
        "torch_geometric.utils": tg_utils,
    })

    tc = types.ModuleType("torch_cluster")
    tc.radius_graph = _radius_graph
    sys.modules["torch_cluster"] = tc

    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = torch.nn.Module
    pl.Trainer = object
    pl.Callback = object
    pl_util = types.ModuleType("pytorch_lightning.utilities")
    pl_util.rank_zero_only = lambda fn: fn
    pl.utilities = pl_util
    sys.modules["pytorch_lightning"] = pl
    sys.modules["pytorch_lightning.utilities"] = pl_util

    oc = types.ModuleType("omegaconf")

    class DictConfig(dict):
        pass

    class OmegaConf:  # noqa: D401 - placeholder
        pass

    oc.DictConfig, oc.OmegaConf = DictConfig, OmegaConf
    sys.modules["omegaconf"] = oc

    # read-out heads (models/components/outputs.py:3,6): torch_scatter.scatter and an `ase` placeholder
    ts = types.ModuleType("torch_scatter")
    ts.scatter = lambda src, index, dim=0, out=None, dim_size=None, reduce="sum": _scatter(
        src, index, dim=dim, dim_size=dim_size, reduce=reduce)
    sys.modules["torch_scatter"] = ts
    ase = types.ModuleType("ase")
    ase_data = types.ModuleType("ase.data")
    import numpy as _np

    # ASE itself is absent offline: the table the product ships (IUPAC 2016, gotennet_b200/atomic_data.py) stands in
    # for ase.data.atomic_masses, so the reference's ElectronicSpatialExtentV2 (outputs.py:513) runs verbatim
    from gotennet_b200.atomic_data import ATOMIC_MASSES as _MASSES

    ase_data.atomic_masses = _np.asarray(_MASSES, dtype=_np.float64)
    ase.data = ase_data
    sys.modules["ase"] = ase
    sys.modules["ase.data"] = ase_data


def import_reference(path: str = "/root/reference"):
    """Return the verbatim reference package (raises if the tree is absent)."""
    import os

    if not os.path.isdir(os.path.join(path, "gotennet")):
        raise FileNotFoundError(path)
    install()
    if path not in sys.path:
        sys.path.insert(0, path)
    import gotennet  # noqa: WPS433

    return gotennet
