"""Read-out head of the path's next row: `Atomwise` energy head with optional
forces, reference models/components/outputs.py:232-376 (+ SchnetMLP components/layers.py:225-273,
ScaleShift :172-202, GetItem :205-222, shifted_softplus :69-81).

Same constructor signature, attribute names and state_dict keys as the reference
(`atomref.weight`, `out_net.1.out_net.{i}.{weight,bias}`, `standardize.{mean,stddev}`), so a
reference checkpoint's `output_modules.0.*` entries load unchanged.  The per-atom MLP runs on
the tcgen05 / SIMT GEMM kernels (SiLU fused in the epilogue), standardisation + atomref +
the per-molecule sum is one deterministic segmented-reduction kernel (the reference scatters
with atomics), and forces come from the hand-written backward of the interaction block.

`GatedEquivariantBlock`, `Dipole` and `ElectronicSpatialExtentV2` (the remaining QM9 heads, SURVEY §8 f4,
reference outputs.py:24-104, :379-542) follow at the bottom of the file on the same kernels.

Not implemented (raises): custom out-nets passed to Atomwise, and second-order autograd through the block -- `create_graph=True`
(the reference default, outputs.py:248) still yields correct first-order forces, but
back-propagating *through* them (force-loss training) raises instead of silently dropping terms.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn
from torch.autograd import grad

from . import ops
from ._lib import GotenError
from .layers import Dense, ShiftedSoftplus, is_silu, str2act


def shifted_softplus(x: torch.Tensor):
    """ln(1 + exp(x)) - ln 2 (reference layers.py:69-81)."""
    return F.softplus(x) - math.log(2.0)


def _act_kind(act) -> int:
    if act is None:
        return ops.ACT_NONE
    if is_silu(act):
        return ops.ACT_SILU
    if act is shifted_softplus or isinstance(act, ShiftedSoftplus):
        return ops.ACT_SSP
    raise NotImplementedError(f"head activation {act!r}: the kernels implement SiLU and shifted softplus")


class ScaleShift(nn.Module):
    """y = x * stddev + mean (reference layers.py:172-202); fused into the read-out kernel by Atomwise."""

    def __init__(self, mean, stddev):
        super().__init__()
        if isinstance(mean, float):
            mean = torch.FloatTensor([mean])
        if isinstance(stddev, float):
            stddev = torch.FloatTensor([stddev])
        self.register_buffer("mean", mean)
        self.register_buffer("stddev", stddev)

    def forward(self, input):
        return input * self.stddev + self.mean


class GetItem(nn.Module):
    """inputs[key] / inputs.key (reference layers.py:205-222)."""

    def __init__(self, key):
        super().__init__()
        self.key = key

    def forward(self, inputs):
        return inputs[self.key] if isinstance(inputs, dict) else getattr(inputs, self.key)


class SchnetMLP(nn.Module):
    """Pyramidal MLP of Dense layers (reference layers.py:225-273)."""

    def __init__(self, n_in, n_out, n_hidden=None, n_layers=2, activation=shifted_softplus):
        super().__init__()
        if n_hidden is None:
            c, self.n_neurons = n_in, []
            for _ in range(n_layers):
                self.n_neurons.append(c)
                c = c // 2
            self.n_neurons.append(n_out)
        else:
            if type(n_hidden) is int:
                n_hidden = [n_hidden] * (n_layers - 1)
            self.n_neurons = [n_in] + n_hidden + [n_out]
        layers = [Dense(self.n_neurons[i], self.n_neurons[i + 1], activation=activation) for i in range(n_layers - 1)]
        layers.append(Dense(self.n_neurons[-2], self.n_neurons[-1], activation=None))
        self.out_net = nn.Sequential(*layers)

    def forward(self, inputs):
        x = inputs
        for layer in self.out_net:
            if layer.norm is not None:
                raise NotImplementedError("normalised Dense layers are not part of the read-out path")
            x = ops.DenseActFn.apply(x, layer.weight, layer.bias, _act_kind(layer.activation))
        return x


class _FirstOrderOnly(torch.autograd.Function):
    """Identity on a derivative computed by the first-order kernels; differentiating through it raises."""

    @staticmethod
    def forward(ctx, dy, pos, name):
        ctx.name = name
        return dy.clone()

    @staticmethod
    def backward(ctx, g):
        raise GotenError(
            f"back-propagating through '{ctx.name}' needs second-order kernels, which gotennet_b200 does not have: "
            "train force-matching losses with gotennet_b200.force_matching_backward(model, head, batch, loss_fn) "
            "(head built with derivative=None), or detach the derivative")


_MODES = {None: 0, "sum": 1, "add": 1, "mean": 2, "avg": 2}


class Atomwise(nn.Module):
    """Per-atom property -> per-molecule property (reference outputs.py:232-376)."""

    def __init__(self, n_in: int, n_out: int = 1, aggregation_mode: Optional[str] = "sum", n_layers: int = 2,
                 n_hidden: Optional[int] = None, activation=shifted_softplus, property: str = "y",
                 contributions: Optional[str] = None, derivative: Optional[str] = None, negative_dr: bool = True,
                 create_graph: bool = True, mean: Optional[torch.Tensor] = None, stddev: Optional[torch.Tensor] = None,
                 atomref: Optional[torch.Tensor] = None, outnet: Optional[nn.Module] = None,
                 return_vector: Optional[str] = None, standardize: bool = True):
        super().__init__()
        if aggregation_mode not in _MODES:
            raise NotImplementedError(f"aggregation_mode {aggregation_mode!r}: sum / mean / None are implemented")
        self.return_vector, self.n_layers, self.create_graph = return_vector, n_layers, create_graph
        self.property, self.contributions, self.derivative = property, contributions, derivative
        self.negative_dr = negative_dr
        mean = torch.FloatTensor([0.0]) if mean is None else mean
        stddev = torch.FloatTensor([1.0]) if stddev is None else stddev
        if type(activation) is str:
            activation = str2act(activation)
        self.atomref = nn.Embedding.from_pretrained(atomref.type(torch.float32)) if atomref is not None else None
        self.equivariant = False
        if outnet is not None:
            raise NotImplementedError("custom / equivariant out-nets are outside the accelerated read-out path")
        self.out_net = nn.Sequential(GetItem("representation"), SchnetMLP(n_in, n_out, n_hidden, n_layers, activation))
        self.standardize = ScaleShift(mean, stddev) if standardize else nn.Identity()
        self.aggregation_mode = aggregation_mode

    def forward(self, inputs):
        z = inputs.z
        with ops.device_of(z):
            return self._forward(inputs, z)

    def _forward(self, inputs, z):
        result = {}
        raw = self.out_net(inputs)  # [N, n_out]
        if not raw.is_cuda:
            raise GotenError("gotennet_b200 kernels need CUDA tensors (there is no CPU path)")
        mode = _MODES[self.aggregation_mode]
        mol_ptr, n_mol = None, 0
        if mode != 0:
            n_mol = getattr(inputs, "num_graphs", None)
            if n_mol is None:  # same host read torch_scatter does to size its output
                n_mol = int(inputs.batch[-1].item()) + 1 if inputs.batch.numel() else 0
            mol_ptr = ops.mol_ptr_from_batch(inputs.batch, int(n_mol))
        sc = self.standardize if isinstance(self.standardize, ScaleShift) else None
        yi, y = ops.AtomwiseReduceFn.apply(raw, z.contiguous().long() if self.atomref is not None else None,
                                           self.atomref.weight if self.atomref is not None else None,
                                           sc.mean.float() if sc is not None else None,
                                           sc.stddev.float() if sc is not None else None, mol_ptr, int(n_mol), mode)
        result[self.property] = y
        if self.contributions:
            result[self.contributions] = yi
        if self.derivative:
            # The kernels are first-order: the derivative is taken WITHOUT building a second-order graph, whatever
            # `create_graph` says (reference outputs.py:365-375 defaults to True).  The returned tensor still joins
            # the autograd graph through _FirstOrderOnly, whose backward raises a pointed error instead of the
            # anonymous "once_differentiable" failure deep inside a training step.
            sign = -1.0 if self.negative_dr else 1.0
            dy = grad(outputs=result[self.property], inputs=[inputs.pos],
                      grad_outputs=torch.ones_like(result[self.property]), create_graph=False, retain_graph=True)[0]
            dy = sign * dy
            if self.create_graph and torch.is_grad_enabled() and inputs.pos.requires_grad:
                dy = _FirstOrderOnly.apply(dy, inputs.pos, self.derivative)
            result[self.derivative] = dy
        return result


def _n_mol_of(inputs) -> int:
    n_mol = getattr(inputs, "num_graphs", None)
    if n_mol is None:  # same host read torch_scatter does to size its output
        n_mol = int(inputs.batch[-1].item()) + 1 if inputs.batch.numel() else 0
    return int(n_mol)


class GatedEquivariantBlock(nn.Module):
    """Invariant + equivariant feature mixing for tensorial read-outs (reference outputs.py:24-104); same constructor,
    attributes and state_dict keys (`mix_vectors.weight`, `scalar_net.{0,1}.{weight,bias}`)."""

    def __init__(self, n_sin: int, n_vin: int, n_sout: int, n_vout: int, n_hidden: int, activation=F.silu,
                 sactivation=None):
        super().__init__()
        self.n_sin, self.n_vin, self.n_sout, self.n_vout, self.n_hidden = n_sin, n_vin, n_sout, n_vout, n_hidden
        self.mix_vectors = Dense(n_vin, 2 * n_vout, activation=None, bias=False)
        self.scalar_net = nn.Sequential(Dense(n_sin + n_vout, n_hidden, activation=activation),
                                        Dense(n_hidden, n_sout + n_vout, activation=None))
        self.sactivation = sactivation

    def forward(self, scalars: torch.Tensor, vectors: torch.Tensor):
        """scalars [N, n_sin], vectors [N, 3, n_vin] -> (s_out [N, n_sout], v_out [N, 3, n_vout])."""
        with ops.device_of(scalars):
            d0, d1 = self.scalar_net[0], self.scalar_net[1]
            return ops.GatedEquivariantFn.apply(scalars, vectors, self.mix_vectors.weight, d0.weight, d0.bias, d1.weight,
                                                d1.bias, self.n_sout, self.n_vout, _act_kind(d0.activation),
                                                _act_kind(self.sactivation))


class Dipole(nn.Module):
    """Dipole-moment read-out (reference outputs.py:379-468): two gated equivariant blocks on (h, X[:, :3]) -> latent
    atomic charges and dipoles; mu = sum_atoms (mu_atom + q_atom * pos); optional magnitude.  Same constructor, result
    keys (`property`, `property + "_vector"`) and state_dict keys (`equivariant_layers.{0,1}.*`)."""

    def __init__(self, n_in: int, n_hidden: Optional[int] = None, activation=F.silu, property: str = "dipole",
                 predict_magnitude: bool = False, output_v: bool = True, mean: Optional[torch.Tensor] = None,
                 stddev: Optional[torch.Tensor] = None):
        super().__init__()
        self.stddev, self.mean, self.output_v = stddev, mean, output_v
        if n_hidden is None:
            n_hidden = n_in
        if type(activation) is str:
            activation = str2act(activation)
        self.property, self.derivative, self.predict_magnitude = property, None, predict_magnitude
        self.equivariant_layers = nn.ModuleList([
            GatedEquivariantBlock(n_sin=n_in, n_vin=n_in, n_sout=n_hidden, n_vout=n_hidden, n_hidden=n_hidden,
                                  activation=activation, sactivation=activation),
            GatedEquivariantBlock(n_sin=n_hidden, n_vin=n_hidden, n_sout=1, n_vout=1, n_hidden=n_hidden,
                                  activation=activation)])
        self.requires_dr = False
        self.requires_stress = False
        self.aggregation_mode = "sum"

    def forward(self, inputs):
        with ops.device_of(inputs.pos):
            l0 = inputs.representation
            l1 = inputs.vector_representation[:, :3, :]
            for eqlayer in self.equivariant_layers:
                l0, l1 = eqlayer(l0, l1)
            standardised = self.stddev is not None
            sd = float(self.stddev) if standardised else 1.0
            mu = float(self.mean) if standardised and self.mean is not None else 0.0
            n_mol = _n_mol_of(inputs)
            mol_ptr = ops.mol_ptr_from_batch(inputs.batch, n_mol)
            yi = ops.DipoleAtomFn.apply(l1.reshape(-1, 3), l0.reshape(-1), inputs.pos, sd, mu)
            _, y = ops.AtomwiseReduceFn.apply(yi, None, None, None, None, mol_ptr, n_mol, 1)
            result = {}
            if self.output_v:
                _, yv = ops.AtomwiseReduceFn.apply(l1.reshape(-1, 3), None, None, None, None, mol_ptr, n_mol, 1)
                result[self.property + "_vector"] = yv.unsqueeze(-1)
            if self.predict_magnitude:
                y = ops.RowNormFn.apply(y)
            result[self.property] = y
            return result


class ElectronicSpatialExtentV2(Atomwise):
    """<r^2> read-out (reference outputs.py:471-542): per-atom scalar from the Atomwise MLP weighted by the squared
    distance to the molecule's centre of mass, summed per molecule.  `atomic_mass` is a buffer as in the reference."""

    def __init__(self, n_in: int, n_layers: int = 2, n_hidden: Optional[int] = None, activation=shifted_softplus,
                 property: str = "y", contributions: Optional[str] = None, mean: Optional[torch.Tensor] = None,
                 stddev: Optional[torch.Tensor] = None, outnet: Optional[nn.Module] = None):
        super().__init__(n_in, 1, "sum", n_layers, n_hidden, activation=activation, mean=mean, stddev=stddev,
                         outnet=outnet, property=property, contributions=contributions)
        from .atomic_data import ATOMIC_MASSES
        self.register_buffer("atomic_mass", torch.tensor(ATOMIC_MASSES, dtype=torch.float32))

    def forward(self, inputs):
        with ops.device_of(inputs.pos):
            x = self.out_net(inputs)
            if not x.is_cuda:
                raise GotenError("gotennet_b200 kernels need CUDA tensors (there is no CPU path)")
            n_mol = _n_mol_of(inputs)
            mol_ptr = ops.mol_ptr_from_batch(inputs.batch, n_mol)
            y = ops.SpatialExtentFn.apply(x, inputs.pos, inputs.z.contiguous().long(), self.atomic_mass.float(), mol_ptr,
                                          n_mol)
            result = {self.property: y}
            if self.contributions:
                result[self.contributions] = x
            return result
