"""Primitive layers of the GotenNet path with the reference's class names, constructor
signatures and state_dict keys (reference components/layers.py), so reference
checkpoints load unchanged.

On the hot path these modules are *parameter containers*: GotenNet/GATA/EQFF read
their weights and hand them to the fused sm_100a kernels (ops.py).  Their own
`forward` methods exist for stand-alone use and also run on our kernels (Dense ->
goten_gemm); the few scalar utilities (CosineCutoff, ExpNormalSmearing, TensorInit
called directly) are closed-form element-wise formulas evaluated with torch
tensor ops on whatever device the input lives on — GotenNet itself never calls
them, it uses the fused geometry kernel.
"""
from __future__ import annotations

import inspect
import math
from functools import partial
from typing import Callable, List, Optional, Union

import torch
import torch.nn.functional as F
from torch import Tensor, nn
from torch.nn.init import constant_, xavier_uniform_

from . import ops
from .graph import radius_graph_plan

zeros_initializer = partial(constant_, val=0.0)


# ------------------------------------------------------------------ registries
def _norm_name(s: str) -> str:
    return s.lower().replace("-", "").replace("_", "").replace(" ", "")


class ShiftedSoftplus(nn.Module):
    def forward(self, x):
        return F.softplus(x) - math.log(2.0)


def str2act(name, *args, **kwargs):
    """Activation by name (reference layers.py:685-701); None / '' -> None."""
    if not name:
        return None
    table = {
        _norm_name(k): v for k, v in vars(torch.nn.modules.activation).items()
        if isinstance(v, type) and issubclass(v, nn.Module)
    }
    table.update({"relu": nn.ReLU, "elu": nn.ELU, "sigmoid": nn.Sigmoid, "silu": nn.SiLU, "mish": nn.Mish,
                  "swish": nn.SiLU, "selu": nn.SELU, "softplus": ShiftedSoftplus})
    if name not in table:
        raise ValueError(f'Invalid choice "{name}", choose one from {", ".join(table)}')
    return table[name]()


def is_silu(act) -> bool:
    return act is F.silu or isinstance(act, nn.SiLU) or act is nn.SiLU


def get_weight_init_by_string(init_str: str) -> Callable:
    """reference layers.py:427-452."""
    if init_str == "":
        return lambda x: x
    if init_str == "zeros":
        return torch.nn.init.zeros_
    if init_str == "xavier_uniform":
        return torch.nn.init.xavier_uniform_
    if init_str == "glo_orthogonal":
        def glorot_orthogonal_(t, scale=2.0):
            torch.nn.init.orthogonal_(t.data)
            t.data *= (scale / ((t.size(-2) + t.size(-1)) * t.var())).sqrt()
            return t
        return glorot_orthogonal_
    if init_str == "he_orthogonal":
        def he_orthogonal_(t):
            torch.nn.init.orthogonal_(t)
            with torch.no_grad():
                var, mean = torch.var_mean(t.data, dim=1, unbiased=True, keepdim=True)
                t.data = (t.data - mean) / (var + 1e-6) ** 0.5
                t.data *= (1 / t.shape[1]) ** 0.5
            return t
        return he_orthogonal_
    raise ValueError(f"Unknown initialization {init_str}")


# ---------------------------------------------------------------------- Dense
class _LinearFn(torch.autograd.Function):
    """y = x W^T + b through goten_gemm (stand-alone Dense calls)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        ctx.save_for_backward(x2, w)
        ctx.has_bias = b is not None
        ctx.shape = x.shape
        return ops.linear_fwd(x2, w.contiguous(), b).view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, g):
        x2, w = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1]).contiguous()
        da, dw, db = ops.linear_bwd(g2, x2, w.contiguous(), need_bias=ctx.has_bias)
        return da.view(ctx.shape), dw, db


class Dense(nn.Linear):
    """Linear -> optional norm -> optional activation (reference layers.py:457-529)."""

    def __init__(self, in_features, out_features, bias=True, activation=None, weight_init=xavier_uniform_,
                 bias_init=zeros_initializer, norm=None, gain=None):
        self.weight_init, self.bias_init, self.gain = weight_init, bias_init, gain
        super().__init__(in_features, out_features, bias)
        self.activation = activation() if inspect.isclass(activation) else activation
        if norm == "layer":
            self.norm = nn.LayerNorm(out_features)
        elif norm == "batch":
            self.norm = nn.BatchNorm1d(out_features)
        elif norm == "instance":
            self.norm = nn.InstanceNorm1d(out_features)
        else:
            self.norm = None

    def reset_parameters(self):
        if self.gain:
            self.weight_init(self.weight, gain=self.gain)
        else:
            self.weight_init(self.weight)
        if self.bias is not None:
            self.bias_init(self.bias)

    def forward(self, inputs):
        y = _LinearFn.apply(inputs, self.weight, self.bias)
        if self.norm is not None:
            y = self.norm(y)
        if self.activation:
            y = self.activation(y)
        return y


class MLP(nn.Module):
    """Stack of Dense layers; every layer is registered under `dense_layers.i` and `layers.i`
    (reference layers.py:533-581) — the aliased keys are part of the checkpoint format."""

    def __init__(self, hidden_dims: List[int], bias=True, activation=None, last_activation=None,
                 weight_init=xavier_uniform_, bias_init=zeros_initializer, norm=""):
        super().__init__()
        mk = partial(Dense, bias=bias, weight_init=weight_init, bias_init=bias_init)
        n = len(hidden_dims)
        self.dense_layers = nn.ModuleList(
            [mk(hidden_dims[i], hidden_dims[i + 1], activation=activation, norm=norm) for i in range(n - 2)]
            + [mk(hidden_dims[-2], hidden_dims[-1], activation=last_activation)]
        )
        self.layers = nn.Sequential(*self.dense_layers)
        self.reset_parameters()

    def reset_parameters(self):
        for m in self.dense_layers:
            m.reset_parameters()

    def forward(self, x):
        return self.layers(x)


# ------------------------------------------------------------- scalar utilities
class CosineCutoff(nn.Module):
    """0.5 (cos(pi d / rc) + 1) [d < rc]   (reference layers.py:133-152)."""

    def __init__(self, cutoff):
        super().__init__()
        self.cutoff = cutoff.item() if isinstance(cutoff, torch.Tensor) else cutoff

    def forward(self, distances):
        return 0.5 * (torch.cos(distances * math.pi / self.cutoff) + 1.0) * (distances < self.cutoff).float()


class ExpNormalSmearing(nn.Module):
    """Exponential-normal radial basis (reference layers.py:703-746).  GotenNet feeds
    `means`/`betas` to the fused geometry kernel; trainable=True is not supported there."""

    def __init__(self, cutoff=5.0, n_rbf=50, trainable=False):
        super().__init__()
        self.cutoff = cutoff.item() if isinstance(cutoff, torch.Tensor) else cutoff
        self.n_rbf, self.trainable = n_rbf, trainable
        self.cutoff_fn = CosineCutoff(self.cutoff)
        self.alpha = 5.0 / self.cutoff
        means, betas = self._initial_params()
        if trainable:
            self.register_parameter("means", nn.Parameter(means))
            self.register_parameter("betas", nn.Parameter(betas))
        else:
            self.register_buffer("means", means)
            self.register_buffer("betas", betas)

    def _initial_params(self):
        start = torch.exp(torch.scalar_tensor(-self.cutoff))
        means = torch.linspace(start, 1, self.n_rbf)
        betas = torch.tensor([(2 / self.n_rbf * (1 - start)) ** -2] * self.n_rbf)
        return means, betas

    def reset_parameters(self):
        means, betas = self._initial_params()
        self.means.data.copy_(means)
        self.betas.data.copy_(betas)

    def forward(self, dist):
        d = dist.unsqueeze(-1)
        return self.cutoff_fn(d) * torch.exp(-self.betas * (torch.exp(self.alpha * (-d)) - self.means) ** 2)

    def kernel_args(self):
        """(basis id, p0, p1) for goten_edge_geometry_*."""
        return 0, self.means, self.betas


class GaussianRBF(nn.Module):
    """Gaussian radial basis (reference layers.py:276-325): exp(-0.5 (d - offset_k)^2 / width_k^2); no cutoff factor."""

    def __init__(self, n_rbf: int, cutoff: float, start: float = 0.0, trainable: bool = False):
        super().__init__()
        if trainable:
            raise NotImplementedError("the fused geometry kernel has no gradient for trainable basis parameters")
        self.n_rbf = n_rbf
        offset = torch.linspace(start, cutoff, n_rbf)
        self.register_buffer("widths", torch.abs(offset[1] - offset[0]) * torch.ones_like(offset))
        self.register_buffer("offsets", offset)

    def forward(self, inputs):
        coeff = -0.5 / torch.pow(self.widths, 2)
        return torch.exp(coeff * torch.pow(inputs[..., None] - self.offsets, 2))

    def kernel_args(self):
        return 2, self.offsets, self.widths


class BesselBasis(nn.Module):
    """0th-order Bessel radial basis sin(k pi d / r_c) / d (reference layers.py:328-358); d = 0 divides by 1."""

    def __init__(self, cutoff=5.0, n_rbf=None, trainable=False):
        super().__init__()
        if n_rbf is None:
            raise ValueError("n_rbf must be specified for BesselBasis")
        self.n_rbf = n_rbf
        self.register_buffer("freqs", torch.arange(1, n_rbf + 1) * math.pi / cutoff)
        self.register_buffer("norm1", torch.tensor(1.0))

    def forward(self, inputs):
        x = inputs[..., None]
        return torch.sin(x * self.freqs[None, :]) / torch.where(x == 0, self.norm1, x)

    def kernel_args(self):
        return 1, self.freqs, self.freqs


def str2basis(name):
    """reference layers.py:749-777 (same quirk: 'GaussianRBF' is matched case-sensitively)."""
    if not isinstance(name, str):
        return name
    if _norm_name(name) == "besselbasis":
        return BesselBasis
    if name == "GaussianRBF":
        return GaussianRBF
    if name.lower() == "expnorm":
        return ExpNormalSmearing
    raise ValueError("Unknown radial basis: {}".format(name))


class TensorInit(nn.Module):
    """Real spherical harmonics of degree 1..l of unit edge vectors, no l=0 term
    (reference layers.py:783-869).  Kept for its `.l` / `.tensor_size` attributes."""

    def __init__(self, l=2):
        super().__init__()
        self.l = l

    @property
    def tensor_size(self):
        return (self.l + 1) ** 2 - 1

    def forward(self, edge_vec):
        if self.l > 3:
            raise NotImplementedError("degrees above 3 are outside the accelerated path")
        x, y, z = edge_vec[..., 0], edge_vec[..., 1], edge_vec[..., 2]
        out = [x, y, z]
        if self.l >= 2:
            s3, y2, x2z2 = math.sqrt(3.0), y * y, x * x + z * z
            a, e = s3 * x * z, s3 / 2.0 * (z * z - x * x)
            out += [a, s3 * x * y, y2 - 0.5 * x2z2, s3 * y * z, e]
            if self.l >= 3:
                c42, c7, c168 = math.sqrt(42.0) / 6.0, math.sqrt(7.0), math.sqrt(168.0) / 8.0
                out += [c42 * (a * z + e * x), c7 * a * y, c168 * (4.0 * y2 - x2z2) * x,
                        0.5 * c7 * y * (2.0 * y2 - 3.0 * x2z2), c168 * z * (4.0 * y2 - x2z2), c7 * e * y,
                        c42 * (e * z - a * x)]
        return torch.stack(out, dim=-1)


# ------------------------------------------------------------------ init blocks
class TensorLayerNorm(nn.Module):
    """Max-min layer normalisation of the steerable features, independently per degree (reference
    components/layers.py:1497-1563).  forward(X [N,L,C]) -> [N,L,C]; the GATA block calls the kernel directly on the
    degree-major layout.  `weight` is a buffer unless `trainable` (the reference's GATA uses trainable=False)."""

    def __init__(self, hidden_channels, trainable, lmax=1, **kwargs):
        super().__init__()
        if trainable:
            raise NotImplementedError("TensorLayerNorm(trainable=True) is not used by the reference model and has no "
                                      "weight gradient here")
        self.hidden_channels, self.eps, self.lmax = hidden_channels, 1e-12, lmax
        self.register_buffer("weight", torch.ones(hidden_channels))

    def reset_parameters(self):
        self.weight.data.fill_(1.0)

    def forward(self, tensor):
        if tensor.shape[1] != (self.lmax + 1) ** 2 - 1:
            raise ValueError(f"TensorLayerNorm received unsupported feature dimension {tensor.shape[1]}")
        Xd = ops.PermuteFn.apply(tensor, True)
        return ops.PermuteFn.apply(ops.TensorLayerNormFn.apply(Xd, self.weight, self.lmax), False)


class NodeInit(nn.Module):
    """Parameters of the node initialisation (reference layers.py:1607-1675):
    A_nbr embedding, W_ndp (rbf -> C), W_nrd_nru (2C -> C -> C with LayerNorm).
    Evaluated inside ops.InitBlockFn by GotenNet.forward."""

    def __init__(self, hidden_channels, num_rbf, cutoff, max_z=100, activation=F.silu, proj_ln="",
                 weight_init=nn.init.xavier_uniform_, bias_init=nn.init.zeros_):
        super().__init__()
        if isinstance(hidden_channels, int):
            hidden_channels = [hidden_channels]
        last = hidden_channels[-1]
        self.A_nbr = nn.Embedding(max_z, last)
        self.W_ndp = MLP([num_rbf, last], activation=None, norm="", weight_init=weight_init, bias_init=bias_init,
                         last_activation=None)
        self.W_nrd_nru = MLP([2 * last] + hidden_channels, activation=activation, norm=proj_ln,
                             weight_init=weight_init, bias_init=bias_init, last_activation=None)
        self.cutoff = CosineCutoff(cutoff)
        self.reset_parameters()

    def reset_parameters(self):
        self.A_nbr.reset_parameters()
        self.W_ndp.reset_parameters()
        self.W_nrd_nru.reset_parameters()


class EdgeInit(nn.Module):
    """Parameters of the edge initialisation, t_ij = (h_i + h_j) * W_erp(phi)
    (reference layers.py:1677-1714).  Evaluated inside ops.InitBlockFn."""

    def __init__(self, num_rbf, hidden_channels, activation=None):
        super().__init__()
        self.W_erp = nn.Linear(num_rbf, hidden_channels)
        self.activation = activation
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.W_erp.weight)
        self.W_erp.bias.data.fill_(0)


class Distance(nn.Module):
    """Radius graph + edge vectors (reference layers.py:1566-1604): returns
    (edge_index [2,E] int64, edge_weight [E], edge_vec [E,3]); edges sorted by (target, source)."""

    def __init__(self, cutoff, max_num_neighbors=32, loop=True, direction="source_to_target"):
        super().__init__()
        if direction not in ["source_to_target", "target_to_source"]:
            raise ValueError(f"Unknown direction '{direction}'. Choose 'source_to_target' or 'target_to_source'.")
        self.direction, self.cutoff, self.max_num_neighbors, self.loop = direction, cutoff, max_num_neighbors, loop

    def plan(self, pos, batch):
        return radius_graph_plan(pos, batch, self.cutoff, self.max_num_neighbors, self.loop)

    def forward(self, pos, batch):
        plan = self.plan(pos, batch)
        ei = plan.edge_index
        edge_vec = pos[ei[0]] - pos[ei[1]]
        mask = ei[0] != ei[1]
        safe = torch.where(mask.unsqueeze(-1), edge_vec, torch.ones_like(edge_vec))
        edge_weight = torch.where(mask, safe.norm(dim=-1), torch.zeros_like(edge_vec[:, 0]))
        return ei, edge_weight, edge_vec
