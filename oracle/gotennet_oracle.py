"""CPU oracle: a plain-PyTorch restatement of the GotenNet interaction stack.

TEST INFRASTRUCTURE ONLY.  The product (gotennet_b200/) never imports this
module; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs do, and there only as the checker / the timed CPU
baseline.  The product path fails loudly when the CUDA library is missing.

Parity pin: the reference ships NO tests and NO golden vectors,
so this restatement is pinned against outputs of the *verbatim reference code*
executed in the build container under `oracle/ref_standins.py`
(tests/golden/make_golden.py -> tests/golden/*.npz, and
tests/test_oracle_vs_reference.py which re-runs the comparison whenever
/root/reference is present).

Everything is functional: parameters come from a `state_dict` that uses the
reference's own key names, so a reference checkpoint can be
fed to the oracle and to the CUDA path unchanged.  All functions are dtype
generic (float32 for parity, float64 as a tie-breaker) and differentiable with
torch.autograd (that is the backward oracle).

Citations are `file:line` relative to /root/reference/gotennet/models/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------
@dataclass
class OracleConfig:
    """Constructor surface of GotenNet that changes the arithmetic
    (representation/gotennet.py:767-793, :1018)."""

    n_atom_basis: int = 128
    n_interactions: int = 8
    n_rbf: int = 32
    cutoff: float = 5.0
    max_z: int = 100
    epsilon: float = 1e-8
    num_heads: int = 8
    edge_updates: bool = True
    scale_edge: bool = True
    lmax: int = 1
    sep_htr: bool = True
    sep_dir: bool = False
    sep_tensor: bool = False
    max_num_neighbors: int = 32
    emlp_dim: Optional[int] = None  # width of gamma_t's hidden layer with the "mlp" / "mlpa" edge updates (:236-242)
    evec_dim: Optional[int] = None  # width of the W_vq / W_vk projections (:236, :252-268); != C needs "linw" / "linwa"
    edge_ln: str = ""               # "layer": LayerNorm inside a two-layer gamma_t (MLP norm, :249, layers.py:563-566)
    radial_basis: str = "expnorm"   # "expnorm" | "BesselBasis" | "GaussianRBF" (layers.py:749-777)
    layernorm: str = ""        # != "": nn.LayerNorm on h at the top of every GATA block (gotennet.py:308-310, :397)
    steerable_norm: str = ""   # != "": TensorLayerNorm on X (gotennet.py:311-315, :398)

    @property
    def L(self) -> int:
        return (self.lmax + 1) ** 2 - 1

    @property
    def S(self) -> int:  # gotennet.py:197-203
        s = 3
        if self.sep_dir:
            s += self.lmax - 1
        if self.sep_tensor:
            s += self.lmax - 1
        return s


def degree_slices(lmax: int) -> List[Tuple[int, int]]:
    """[start, stop) of each degree-l block in the L axis (gotennet.py:37-51)."""
    out, s = [], 0
    for l in range(1, lmax + 1):
        out.append((s, s + 2 * l + 1))
        s += 2 * l + 1
    return out


# --------------------------------------------------------------------------
# state_dict helpers
# --------------------------------------------------------------------------
def state_dict_spec(cfg: OracleConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) for every *unique* tensor, in the reference's
    registration order.  kind in {weight,bias,ln_w,ln_b,emb,buf}."""
    C, R, S = cfg.n_atom_basis, cfg.n_rbf, cfg.S
    sp: List[Tuple[str, Tuple[int, ...], str]] = []
    sp.append(("node_init.A_nbr.weight", (cfg.max_z, C), "emb"))
    sp.append(("node_init.W_ndp.dense_layers.0.weight", (C, R), "weight"))
    sp.append(("node_init.W_ndp.dense_layers.0.bias", (C,), "bias"))
    sp.append(("node_init.W_nrd_nru.dense_layers.0.weight", (C, 2 * C), "weight"))
    sp.append(("node_init.W_nrd_nru.dense_layers.0.bias", (C,), "bias"))
    sp.append(("node_init.W_nrd_nru.dense_layers.0.norm.weight", (C,), "ln_w"))
    sp.append(("node_init.W_nrd_nru.dense_layers.0.norm.bias", (C,), "ln_b"))
    sp.append(("node_init.W_nrd_nru.dense_layers.1.weight", (C, C), "weight"))
    sp.append(("node_init.W_nrd_nru.dense_layers.1.bias", (C,), "bias"))
    sp.append(("edge_init.W_erp.weight", (C, R), "weight"))
    sp.append(("edge_init.W_erp.bias", (C,), "bias"))
    sp.append(("A_na.weight", (cfg.max_z, C), "emb0"))
    for i in range(cfg.n_interactions):
        p = f"gata_list.{i}."
        last = i == cfg.n_interactions - 1
        for g in ("gamma_s", "gamma_v"):
            sp.append((p + f"{g}.0.weight", (C, C), "weight"))
            sp.append((p + f"{g}.0.bias", (C,), "bias"))
            sp.append((p + f"{g}.1.weight", (S * C, C), "weight"))
            sp.append((p + f"{g}.1.bias", (S * C,), "bias"))
        for g in ("W_q", "W_k", "W_re"):
            sp.append((p + f"{g}.weight", (C, C), "weight"))
            sp.append((p + f"{g}.bias", (C,), "bias"))
        sp.append((p + "W_rs.weight", (S * C, C), "weight"))
        sp.append((p + "W_rs.bias", (S * C,), "bias"))
        if cfg.layernorm:
            sp.append((p + "layernorm.weight", (C,), "ln_w"))
            sp.append((p + "layernorm.bias", (C,), "ln_b"))
        if cfg.steerable_norm:
            sp.append((p + "tensor_layernorm.weight", (C,), "ln_w"))   # a buffer in the reference (trainable=False)
        if not last and cfg.edge_updates:
            eu = cfg.edge_updates.split("_") if isinstance(cfg.edge_updates, str) else []
            if "mlp" in eu or "mlpa" in eu:                            # two-layer gamma_t, gotennet.py:239-250
                Em = cfg.emlp_dim or C
                sp.append((p + "gamma_t.dense_layers.0.weight", (Em, C), "weight"))
                sp.append((p + "gamma_t.dense_layers.0.bias", (Em,), "bias"))
                sp.append((p + "gamma_t.dense_layers.1.weight", (C, Em), "weight"))
                sp.append((p + "gamma_t.dense_layers.1.bias", (C,), "bias"))
            else:
                sp.append((p + "gamma_t.dense_layers.0.weight", (C, C), "weight"))
                sp.append((p + "gamma_t.dense_layers.0.bias", (C,), "bias"))
            if cfg.edge_ln and ("mlp" in eu or "mlpa" in eu):          # Dense(norm="layer") on the hidden layer
                sp.append((p + "gamma_t.dense_layers.0.norm.weight", (cfg.emlp_dim or C,), "ln_w"))
                sp.append((p + "gamma_t.dense_layers.0.norm.bias", (cfg.emlp_dim or C,), "ln_b"))
            Ev = cfg.evec_dim or C
            sp.append((p + "W_vq.weight", (Ev, C), "weight"))
            if cfg.sep_htr:
                for l in range(cfg.lmax):
                    sp.append((p + f"W_vk.{l}.weight", (Ev, C), "weight"))
            else:
                sp.append((p + "W_vk.weight", (Ev, C), "weight"))
            lin_w, lin_ln = edge_lin_flags(cfg)
            if lin_w:                                                  # gamma_w network, gotennet.py:270-282
                if lin_ln == 1:
                    sp.append((p + "gamma_w.0.weight", (Ev,), "ln_w"))
                    sp.append((p + "gamma_w.0.bias", (Ev,), "ln_b"))
                sp.append((p + "W_edp.weight", (C, Ev), "weight"))
                sp.append((p + "W_edp.bias", (C,), "bias"))
                if lin_ln == 2:
                    sp.append((p + "W_edp.norm.weight", (C,), "ln_w"))
                    sp.append((p + "W_edp.norm.bias", (C,), "ln_b"))
    for i in range(cfg.n_interactions):
        p = f"eqff_list.{i}."
        sp.append((p + "gamma_m.0.weight", (C, 2 * C), "weight"))
        sp.append((p + "gamma_m.0.bias", (C,), "bias"))
        sp.append((p + "gamma_m.1.weight", (2 * C, C), "weight"))
        sp.append((p + "gamma_m.1.bias", (2 * C,), "bias"))
        sp.append((p + "W_vu.weight", (C, C), "weight"))
    return sp


def edge_lin_flags(cfg: OracleConfig) -> Tuple[int, int]:
    """(lin_w, lin_ln) of GATA.update_info (gotennet.py:178-185): lin_w 1 = "linw", 2 = "linwa" (activation before
    W_edp); lin_ln 1 = "ln" (LayerNorm on w before W_edp), 2 = "postln" (LayerNorm after W_edp)."""
    parts = cfg.edge_updates.split("_") if isinstance(cfg.edge_updates, str) else []
    lin_w = 2 if "linwa" in parts else (1 if "linw" in parts else 0)
    lin_ln = 2 if "postln" in parts else (1 if "ln" in parts else 0)
    return lin_w, lin_ln


def rbf_buffers(cfg: OracleConfig, dtype=torch.float32) -> Tuple[Tensor, Tensor]:
    """ExpNormalSmearing._initial_params (components/layers.py:733-737)."""
    start = torch.exp(torch.scalar_tensor(-cfg.cutoff))
    means = torch.linspace(start, 1, cfg.n_rbf)
    betas = torch.tensor([(2 / cfg.n_rbf * (1 - start)) ** -2] * cfg.n_rbf)
    return means.to(dtype), betas.to(dtype)


def make_state_dict(cfg: OracleConfig, seed: int = 0, bias_scale: float = 0.1,
                    dtype=torch.float32) -> Dict[str, Tensor]:
    """Deterministic synthetic weights shared by the golden generator, the tests
    and the benchmark.  Weights are xavier-uniform *shaped* (same bounds as the
    reference initialiser, layers.py:503-509) but drawn from one seeded CPU
    generator in spec order, and biases / LayerNorm affine are made NON-trivial
    (the reference zero-initialises them, which would hide bias bugs)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for key, shape, kind in state_dict_spec(cfg):
        if kind == "weight":
            bound = math.sqrt(6.0 / (shape[0] + shape[1]))
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound
        elif kind == "bias":
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bias_scale
        elif kind == "ln_w":
            t = 1.0 + (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bias_scale
        elif kind == "ln_b":
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bias_scale
        else:  # embeddings ~ N(0,1); A_na has padding_idx=0 (gotennet.py:856)
            t = torch.randn(shape, generator=g, dtype=torch.float64)
            if kind == "emb0":
                t[0].zero_()
        sd[key] = t.to(dtype)
    means, betas = rbf_buffers(cfg, dtype)
    if cfg.radial_basis == "expnorm":
        sd["radial_basis.means"], sd["radial_basis.betas"] = means, betas
    elif cfg.radial_basis == "BesselBasis":                        # layers.py:345-347
        sd["radial_basis.freqs"] = (torch.arange(1, cfg.n_rbf + 1) * math.pi / cfg.cutoff).to(dtype)
        sd["radial_basis.norm1"] = torch.tensor(1.0, dtype=dtype)
    elif cfg.radial_basis == "GaussianRBF":                        # layers.py:313-316
        offset = torch.linspace(0.0, cfg.cutoff, cfg.n_rbf)
        sd["radial_basis.widths"] = (torch.abs(offset[1] - offset[0]) * torch.ones_like(offset)).to(dtype)
        sd["radial_basis.offsets"] = offset.to(dtype)
    else:
        raise ValueError(cfg.radial_basis)
    return sd


def expand_aliases(sd: Dict[str, Tensor], cfg: Optional[OracleConfig] = None) -> Dict[str, Tensor]:
    """MLP registers each Dense under `dense_layers.i` AND `layers.i`
    (layers.py:566-571): add the aliased duplicate keys the reference expects.  With the "linw" edge updates `W_edp`
    is both an attribute and an element of the `gamma_w` Sequential (gotennet.py:277-292): its position there depends
    on the switches, hence `cfg`."""
    out = dict(sd)
    for k, v in sd.items():
        if ".dense_layers." in k:
            out[k.replace(".dense_layers.", ".layers.")] = v
        if ".W_edp." in k:
            if cfg is None:
                raise ValueError("expand_aliases needs cfg for the gamma_w aliases of W_edp")
            lin_w, lin_ln = edge_lin_flags(cfg)
            pos = (1 if lin_ln == 1 else 0) + (1 if lin_w == 2 else 0)
            out[k.replace(".W_edp.", f".gamma_w.{pos}.")] = v
    return out


# --------------------------------------------------------------------------
# graph + geometry  (components/layers.py:1566-1604; torch_cluster semantics)
# --------------------------------------------------------------------------
def radius_graph(pos: Tensor, batch: Tensor, r: float, max_num_neighbors: int = 32,
                 loop: bool = True) -> Tensor:
    """torch_cluster.radius_graph, CUDA-build semantics (call at layers.py:1589):
    strict d^2 < r^2 in the dtype of `pos`, same molecule only, at most K sources
    per target kept in ascending source index, output sorted by (target, source).
    Row 0 = source j, row 1 = target i.  Done per molecule (dense)."""
    n = pos.size(0)
    if n == 0:
        return torch.zeros(2, 0, dtype=torch.long)
    # molecule boundaries (batch is sorted, PyG convention)
    change = torch.ones(n, dtype=torch.bool)
    change[1:] = batch[1:] != batch[:-1]
    starts = change.nonzero().flatten().tolist() + [n]
    r2 = torch.tensor(r, dtype=pos.dtype) ** 2
    srcs, tgts = [], []
    for a, b in zip(starts[:-1], starts[1:]):
        p = pos[a:b]
        d = p.unsqueeze(1) - p.unsqueeze(0)          # [target, source, 3]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        ok = d2 < r2
        if not loop:
            ok &= ~torch.eye(b - a, dtype=torch.bool)
        ok &= ok.long().cumsum(1) <= max_num_neighbors
        t, s = ok.nonzero(as_tuple=True)
        srcs.append(s + a)
        tgts.append(t + a)
    return torch.stack([torch.cat(srcs), torch.cat(tgts)], 0)


def edge_geometry(pos: Tensor, edge_index: Tensor) -> Tuple[Tensor, Tensor]:
    """Distance.forward (layers.py:1591-1604): edge_vec = pos[src] - pos[tgt];
    edge_weight = |edge_vec| on non-loop edges, exactly 0 on self loops."""
    src, tgt = edge_index[0], edge_index[1]
    vec = pos[src] - pos[tgt]
    mask = src != tgt
    # norm only where it is differentiable (self-loop rows stay 0, as the
    # reference's masked assignment does)
    safe = torch.where(mask.unsqueeze(-1), vec, torch.ones_like(vec))
    w = torch.where(mask, safe.norm(dim=-1), torch.zeros_like(vec[:, 0]))
    return w, vec


def cosine_cutoff(d: Tensor, rc: float) -> Tensor:
    """CosineCutoff.forward (layers.py:149-152)."""
    return 0.5 * (torch.cos(d * math.pi / rc) + 1.0) * (d < rc).to(d.dtype)


def expnorm_rbf(d: Tensor, means: Tensor, betas: Tensor, rc: float) -> Tensor:
    """ExpNormalSmearing.forward (layers.py:744-746), alpha = 5/rc (:725)."""
    d = d.unsqueeze(-1)
    alpha = 5.0 / rc
    return cosine_cutoff(d, rc) * torch.exp(-betas * (torch.exp(alpha * (-d)) - means) ** 2)


def radial_basis(sd, cfg: OracleConfig, d: Tensor) -> Tensor:
    """self.radial_basis(edge_diff) (gotennet.py:974) for the three registered bases (layers.py:749-777)."""
    if cfg.radial_basis == "expnorm":
        return expnorm_rbf(d, sd["radial_basis.means"], sd["radial_basis.betas"], cfg.cutoff)
    x = d.unsqueeze(-1)
    if cfg.radial_basis == "BesselBasis":                          # layers.py:349-358
        norm = torch.where(x == 0, sd["radial_basis.norm1"], x)
        return torch.sin(x * sd["radial_basis.freqs"][None, :]) / norm
    coeff = -0.5 / torch.pow(sd["radial_basis.widths"], 2)         # layers.py:288-291
    return torch.exp(coeff * torch.pow(x - sd["radial_basis.offsets"], 2))


def sph_harm(lmax: int, u: Tensor) -> Tensor:
    """TensorInit._calculate_components (layers.py:805-869), degrees 1..lmax, no l=0."""
    x, y, z = u[..., 0], u[..., 1], u[..., 2]
    out = [x, y, z]
    if lmax >= 2:
        s3 = math.sqrt(3.0)
        y2 = y * y
        x2z2 = x * x + z * z
        s20, s21, s22, s23, s24 = s3 * x * z, s3 * x * y, y2 - 0.5 * x2z2, s3 * y * z, s3 / 2.0 * (z * z - x * x)
        out += [s20, s21, s22, s23, s24]
    if lmax >= 3:
        c42, c7, c168 = math.sqrt(42.0) / 6.0, math.sqrt(7.0), math.sqrt(168.0) / 8.0
        out += [
            c42 * (s20 * z + s24 * x),
            c7 * s20 * y,
            c168 * (4.0 * y2 - x2z2) * x,
            0.5 * c7 * y * (2.0 * y2 - 3.0 * x2z2),
            c168 * z * (4.0 * y2 - x2z2),
            c7 * s24 * y,
            c42 * (s24 * z - s20 * x),
        ]
    if lmax >= 4:
        raise NotImplementedError("oracle restates degrees 1..3 (layers.py:822-869)")
    return torch.stack(out, dim=-1)


# --------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------
def _lin(sd, key, x, bias=True):
    return F.linear(x, sd[key + ".weight"], sd[key + ".bias"] if bias else None)


def node_init(sd, cfg: OracleConfig, z, h0, edge_index, r, phi) -> Tensor:
    """NodeInit.forward/message (layers.py:1658-1675): self loops dropped, cosine
    cutoff applied on top of phi (which already holds one), sum at targets, then
    Dense(2C->C)+LayerNorm+SiLU, Dense(C->C) (:1646-1649, Dense order :523-528)."""
    src, tgt = edge_index[0], edge_index[1]
    m = src != tgt
    src, tgt, r, phi = src[m], tgt[m], r[m], phi[m]
    feat = _lin(sd, "node_init.W_ndp.dense_layers.0", phi) * cosine_cutoff(r, cfg.cutoff).unsqueeze(-1)
    msg = sd["node_init.A_nbr.weight"][z][src] * feat
    agg = torch.zeros_like(h0).index_add_(0, tgt, msg)
    y = _lin(sd, "node_init.W_nrd_nru.dense_layers.0", torch.cat([h0, agg], dim=1))
    y = F.layer_norm(y, (cfg.n_atom_basis,), sd["node_init.W_nrd_nru.dense_layers.0.norm.weight"],
                     sd["node_init.W_nrd_nru.dense_layers.0.norm.bias"], 1e-5)
    return _lin(sd, "node_init.W_nrd_nru.dense_layers.1", F.silu(y))


def edge_init(sd, edge_index, phi, h) -> Tensor:
    """EdgeInit.message (layers.py:1704-1714): t_ij = (h_i + h_j) * W_erp(phi), loops kept."""
    return (h[edge_index[1]] + h[edge_index[0]]) * _lin(sd, "edge_init.W_erp", phi)


def segment_softmax(a: Tensor, index: Tensor, n: int) -> Tensor:
    """torch_geometric.utils.softmax (call at gotennet.py:503): detached max shift,
    denominator + 1e-16."""
    idx = index.view(-1, *([1] * (a.dim() - 1))).expand_as(a)
    mx = a.new_full((n,) + a.shape[1:], float("-inf")).scatter_reduce_(0, idx, a.detach(), "amax")
    e = (a - mx[index]).exp()
    den = a.new_zeros((n,) + a.shape[1:]).index_add_(0, index, e) + 1e-16
    return e / den[index]


def tensor_layernorm(X, weight, lmax: int, eps: float = 1e-12):
    """TensorLayerNorm.forward / max_min_norm (components/layers.py:1523-1563) on X [N,L,C].  The reference's
    `(dist == 0).all()` early-out returns zeros, which is what the formula below yields for an all-zero part."""
    parts = []
    for a0, a1 in degree_slices(lmax):
        part = X[:, a0:a1]
        dist = torch.norm(part, dim=1, keepdim=True)                  # :1525
        dist = dist.clamp(min=eps)                                     # :1530
        direct = part / dist                                           # :1531
        max_val, _ = torch.max(dist, dim=-1)                           # :1533
        min_val, _ = torch.min(dist, dim=-1)
        delta = (max_val - min_val).view(-1)
        delta = torch.where(delta == 0, torch.ones_like(delta), delta)  # :1536
        dist = (dist - min_val.view(-1, 1, 1)) / delta.view(-1, 1, 1)   # :1537
        parts.append(F.relu(dist) * direct)                            # :1539
    return torch.cat(parts, dim=1) * weight.unsqueeze(0).unsqueeze(0)  # :1557-1563


def gata_layer(sd, cfg: OracleConfig, i: int, edge_index, h, X, Y, t, r, n_edges, drop=None):
    """GATA.forward / message / aggregate / edge_update (gotennet.py:366-640).
    h [N,C], X [N,L,C], Y=rl_ij [E,L], t [E,C], r [E], n_edges [E].
    drop [E,H] (optional): the realised attention-dropout factors mask / (1 - p) of F.dropout at :513 (training
    mode); None = eval mode / p = 0."""
    p = f"gata_list.{i}."
    C, H, S, lmax = cfg.n_atom_basis, cfg.num_heads, cfg.S, cfg.lmax
    last = i == cfg.n_interactions - 1
    src, tgt = edge_index[0], edge_index[1]
    N, E = h.size(0), t.size(0)
    if cfg.layernorm:                                                 # :397
        h = F.layer_norm(h, (C,), sd[p + "layernorm.weight"], sd[p + "layernorm.bias"], 1e-5)
    if cfg.steerable_norm:                                            # :398
        X = tensor_layernorm(X, sd[p + "tensor_layernorm.weight"], lmax)

    q = _lin(sd, p + "W_q", h).view(N, H, C // H)                     # :400
    k = _lin(sd, p + "W_k", h).view(N, H, C // H)                     # :401
    x = _lin(sd, p + "gamma_s.1", F.silu(_lin(sd, p + "gamma_s.0", h)))  # :404
    v = _lin(sd, p + "gamma_v.1", F.silu(_lin(sd, p + "gamma_v.0", h)))  # :405
    ta = F.silu(_lin(sd, p + "W_re", t)).view(E, H, C // H)           # :406, :497
    tf = _lin(sd, p + "W_rs", t)                                      # :407

    a = (q[tgt] * k[src] * ta).sum(-1)                                # :502  [E,H]
    alpha = segment_softmax(a, tgt, N)                                # :503
    if cfg.scale_edge:                                                # :506-511
        alpha = alpha * (torch.sqrt(n_edges).view(-1, 1) / math.sqrt(C))
    else:
        alpha = alpha * (1.0 / math.sqrt(C))
    if drop is not None:                                              # :513 F.dropout(attn, p, training)
        alpha = alpha * drop
    sea = (alpha.unsqueeze(-1) * v[src].view(E, H, S * C // H)).reshape(E, S * C)  # :516-519
    spatial = tf * x[src] * cosine_cutoff(r, cfg.cutoff).unsqueeze(-1)  # :522-526
    o = (spatial + sea).view(E, S, C)                                 # :529-532

    blocks = degree_slices(lmax)
    kd = [1 + l for l in range(lmax)] if cfg.sep_dir else [1] * lmax  # :538-545
    nd = lmax if cfg.sep_dir else 1
    kt = [1 + nd + l for l in range(lmax)] if cfg.sep_tensor else [1 + nd] * lmax  # :548-555
    Xs = X[src]
    dX = torch.cat([Y[:, a0:a1, None] * o[:, kd[l], None, :] + Xs[:, a0:a1, :] * o[:, kt[l], None, :]
                    for l, (a0, a1) in enumerate(blocks)], dim=1)      # :558
    h = h + torch.zeros_like(h).index_add_(0, tgt, o[:, 0, :])         # :638, :426
    X = X + torch.zeros_like(X).index_add_(0, tgt, dX)                 # :639, :427

    if not last and cfg.edge_updates:                                  # :429-447
        EQ = F.linear(X, sd[p + "W_vq.weight"])
        if cfg.sep_htr:
            EK = torch.cat([F.linear(X[:, a0:a1], sd[p + f"W_vk.{l}.weight"])
                            for l, (a0, a1) in enumerate(blocks)], dim=1)
            groups = blocks
        else:
            EK = F.linear(X, sd[p + "W_vk.weight"])
            groups = [(0, cfg.L)]
        EQi, EKj = EQ[tgt], EK[src]                                    # _i = target, _j = source
        w = 0
        rej = not (isinstance(cfg.edge_updates, str) and "norej" in cfg.edge_updates.split("_"))   # :176-177
        for a0, a1 in groups:                                          # :580-609 (rejection :351-364)
            y = Y[:, a0:a1, None]
            Qr, Kr = EQi[:, a0:a1], EKj[:, a0:a1]
            if rej:
                Qr = Qr - (Qr * y).sum(1, keepdim=True) * y
                Kr = Kr - (Kr * y).sum(1, keepdim=True) * y
            w = w + (Qr * Kr).sum(1)
        parts = cfg.edge_updates.split("_") if isinstance(cfg.edge_updates, str) else []
        lin_w, lin_ln = edge_lin_flags(cfg)
        if lin_w:                                                      # gamma_w network on w [E, evec_dim], :270-282
            if lin_ln == 1:
                w = F.layer_norm(w, w.shape[-1:], sd[p + "gamma_w.0.weight"], sd[p + "gamma_w.0.bias"], 1e-5)
            if lin_w == 2:
                w = F.silu(w)
            w = _lin(sd, p + "W_edp", w)
            if lin_ln == 2:                                            # Dense(norm="layer"): after the linear map
                w = F.layer_norm(w, w.shape[-1:], sd[p + "W_edp.norm.weight"], sd[p + "W_edp.norm.bias"], 1e-5)
        if "act" in parts:                                             # gamma_w, :283-289 (later parts win, :168-173)
            w = F.silu(w)
        elif "gatedt" in parts:
            w = torch.tanh(w)
        elif "gated" in parts:
            w = torch.sigmoid(w)
        gt = _lin(sd, p + "gamma_t.dense_layers.0", t)
        if cfg.edge_ln and ("mlp" in parts or "mlpa" in parts):        # Dense: linear -> norm -> activation (layers.py:523-528)
            gt = F.layer_norm(gt, gt.shape[-1:], sd[p + "gamma_t.dense_layers.0.norm.weight"],
                              sd[p + "gamma_t.dense_layers.0.norm.bias"], 1e-5)
        gt = F.silu(gt)
        if "mlp" in parts or "mlpa" in parts:                          # MLP([C, emlp, C]), last activation None for "mlp"
            gt = _lin(sd, p + "gamma_t.dense_layers.1", gt)
            if "mlp" not in parts:
                gt = F.silu(gt)
        t = t + gt * w                                                 # :611, :445
    return h, X, t


def eqff_layer(sd, cfg: OracleConfig, i: int, h, X):
    """EQFF.forward (gotennet.py:728-748)."""
    p = f"eqff_list.{i}."
    C = cfg.n_atom_basis
    P = F.linear(X, sd[p + "W_vu.weight"])
    n = torch.sqrt((P * P).sum(dim=-2) + cfg.epsilon)
    m = _lin(sd, p + "gamma_m.1", F.silu(_lin(sd, p + "gamma_m.0", torch.cat([h, n], dim=-1))))
    return h + m[:, :C], X + m[:, None, C:] * P


def gotennet_forward(sd, cfg: OracleConfig, z, edge_index, edge_diff, edge_vec,
                     intermediates: Optional[dict] = None, drop_masks=None):
    """GotenNet.forward (gotennet.py:956-1010).  Unlike the reference, `edge_vec`
    is NOT mutated in place (quirk App. C.3)."""
    src, tgt = edge_index[0], edge_index[1]
    h = sd["A_na.weight"][z]                                           # :973
    phi = radial_basis(sd, cfg, edge_diff)                              # :974
    h = node_init(sd, cfg, z, h, edge_index, edge_diff, phi)           # :976
    t = edge_init(sd, edge_index, phi, h)                              # :977
    mask = (src != tgt).unsqueeze(-1)                                  # :978-980
    safe = torch.where(mask, edge_vec, torch.ones_like(edge_vec))
    u = torch.where(mask, edge_vec / safe.norm(dim=1, keepdim=True), edge_vec)
    Y = sph_harm(cfg.lmax, u)                                          # :982
    deg = torch.zeros(h.size(0), dtype=h.dtype).index_add_(0, src, torch.ones_like(edge_diff))  # :986-988
    n_edges = deg[src]                                                 # :989
    X = h.new_zeros(h.size(0), cfg.L, cfg.n_atom_basis)                # :992
    if intermediates is not None:
        intermediates.update(phi=phi, h0=h, t0=t, Y=Y, n_edges=n_edges)
    for i in range(cfg.n_interactions):                                # :995-1007
        h, X, t = gata_layer(sd, cfg, i, edge_index, h, X, Y, t, edge_diff, n_edges,
                             None if drop_masks is None else drop_masks[i])
        h, X = eqff_layer(sd, cfg, i, h, X)
        if intermediates is not None:
            intermediates[f"h{i + 1}"], intermediates[f"X{i + 1}"], intermediates[f"t{i + 1}"] = h, X, t
    return h, X


def wrapper_forward(sd, cfg: OracleConfig, z, pos, batch, intermediates: Optional[dict] = None, drop_masks=None):
    """GotenNetWrapper.forward (gotennet.py:1043-1045).  drop_masks: per-layer [E,H] attention-dropout factors."""
    ei = radius_graph(pos.detach(), batch, cfg.cutoff, cfg.max_num_neighbors, loop=True)
    w, vec = edge_geometry(pos, ei)
    if intermediates is not None:
        intermediates.update(edge_index=ei, edge_weight=w, edge_vec=vec)
    return gotennet_forward(sd, cfg, z, ei, w, vec, intermediates, drop_masks)


# --------------------------------------------------------------------------
# synthetic molecules (SURVEY.md §8d generator)
# --------------------------------------------------------------------------
# --------------------------------------------------------------------------
# read-out head (components/outputs.py:232-376 Atomwise)
# --------------------------------------------------------------------------
def head_neurons(n_in: int, n_out: int = 1, n_layers: int = 2, n_hidden=None) -> List[int]:
    """SchnetMLP layer widths (components/layers.py:241-257): pyramidal halving when n_hidden is None."""
    if n_hidden is None:
        c, out = n_in, []
        for _ in range(n_layers):
            out.append(c)
            c = c // 2
        return out + [n_out]
    if isinstance(n_hidden, int):
        n_hidden = [n_hidden] * (n_layers - 1)
    return [n_in] + list(n_hidden) + [n_out]


def make_head_state_dict(n_in: int, n_out: int = 1, n_layers: int = 2, n_hidden=None, seed: int = 0, max_z: int = 100,
                         atomref: bool = True, dtype=torch.float32) -> Dict[str, Tensor]:
    """Deterministic weights under the reference's Atomwise key names (probe: `atomref.weight`,
    `out_net.1.out_net.i.{weight,bias}`, `standardize.{mean,stddev}`)."""
    g = torch.Generator().manual_seed(10_000 + seed)
    nn_ = head_neurons(n_in, n_out, n_layers, n_hidden)
    sd: Dict[str, Tensor] = {}
    if atomref:
        sd["atomref.weight"] = torch.randn(max_z, n_out, generator=g, dtype=torch.float64).to(dtype)
    for i in range(len(nn_) - 1):
        bound = math.sqrt(6.0 / (nn_[i] + nn_[i + 1]))
        sd[f"out_net.1.out_net.{i}.weight"] = ((torch.rand(nn_[i + 1], nn_[i], generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
        sd[f"out_net.1.out_net.{i}.bias"] = ((torch.rand(nn_[i + 1], generator=g, dtype=torch.float64) * 2 - 1) * 0.1).to(dtype)
    sd["standardize.mean"] = torch.tensor([0.37], dtype=dtype)
    sd["standardize.stddev"] = torch.tensor([1.83], dtype=dtype)
    return sd


def atomwise_forward(sd: Dict[str, Tensor], h: Tensor, z: Tensor, batch: Tensor, n_mol: int, activation: str = "silu",
                     aggregation_mode: Optional[str] = "sum") -> Tuple[Tensor, Tensor]:
    """Atomwise.forward (outputs.py:323-363): SchnetMLP on the scalar representation (Dense = linear -> act,
    layers.py:523-528; last layer linear) -> ScaleShift (layers.py:200) -> + atomref[z] (outputs.py:349-351) ->
    scatter over molecules (outputs.py:354-355).  Returns (y [n_mol, n_out] or yi when mode is None, yi [N, n_out])."""
    x = h
    n_lin = len([k for k in sd if k.startswith("out_net.1.out_net.") and k.endswith(".weight")])
    for i in range(n_lin):
        x = F.linear(x, sd[f"out_net.1.out_net.{i}.weight"], sd[f"out_net.1.out_net.{i}.bias"])
        if i < n_lin - 1:
            x = F.silu(x) if activation == "silu" else F.softplus(x) - math.log(2.0)  # layers.py:81, :619
    yi = x * sd["standardize.stddev"] + sd["standardize.mean"]
    if "atomref.weight" in sd:
        yi = yi + sd["atomref.weight"][z]
    if aggregation_mode is None:
        return yi, yi
    y = torch.zeros(n_mol, yi.shape[1], dtype=yi.dtype).index_add_(0, batch, yi)
    if aggregation_mode == "mean":
        y = y / torch.bincount(batch, minlength=n_mol).clamp(min=1).to(yi.dtype).unsqueeze(1)
    return y, yi


def energy_and_forces(sd, sd_head, cfg: OracleConfig, z, pos, batch, n_mol: int, activation: str = "silu",
                      drop_masks=None):
    """Representation + Atomwise(derivative=..., negative_dr=True) (goten_model.py:289, outputs.py:365-375):
    E [n_mol,1] and F = -dE/dpos [N,3] (differentiable: create_graph=True)."""
    if not pos.requires_grad:
        pos = pos.clone().requires_grad_(True)
    h, X = wrapper_forward(sd, cfg, z, pos, batch, drop_masks=drop_masks)
    y, yi = atomwise_forward(sd_head, h, z, batch, n_mol, activation)
    (dy,) = torch.autograd.grad(y, [pos], grad_outputs=torch.ones_like(y), create_graph=True, retain_graph=True)
    return y, -dy, h


# --------------------------------------------------------------------------
# equivariant read-out heads (components/outputs.py:24-104, :379-542)
# --------------------------------------------------------------------------
def _head_act(name: Optional[str], x: Tensor) -> Tensor:
    if name is None:
        return x
    return F.silu(x) if name == "silu" else F.softplus(x) - math.log(2.0)


def gated_equivariant_block(sd, prefix: str, scalars: Tensor, vectors: Tensor, n_sout: int, n_vout: int,
                            activation: str = "silu", sactivation: Optional[str] = None):
    """GatedEquivariantBlock.forward (outputs.py:76-103): scalars [N,n_sin], vectors [N,3,n_vin]."""
    vmix = F.linear(vectors, sd[prefix + "mix_vectors.weight"])                     # :89 (no bias)
    V, W = torch.split(vmix, n_vout, dim=-1)                                        # :90
    Vn = torch.norm(V, dim=-2)                                                      # :91
    ctx = torch.cat([scalars, Vn], dim=-1)                                          # :93
    x = _head_act(activation, F.linear(ctx, sd[prefix + "scalar_net.0.weight"], sd[prefix + "scalar_net.0.bias"]))
    x = F.linear(x, sd[prefix + "scalar_net.1.weight"], sd[prefix + "scalar_net.1.bias"])   # :94
    s_out, g = torch.split(x, [n_sout, n_vout], dim=-1)                             # :95
    v_out = g.unsqueeze(-2) * W                                                     # :96
    return _head_act(sactivation, s_out), v_out                                     # :98-101


def make_dipole_state_dict(n_in: int, n_hidden: Optional[int] = None, seed: int = 0, dtype=torch.float32):
    """Deterministic weights under the reference's Dipole key names (`equivariant_layers.{0,1}.*`, outputs.py:420-427)."""
    g = torch.Generator().manual_seed(20_000 + seed)
    nh = n_in if n_hidden is None else n_hidden
    sd: Dict[str, Tensor] = {}

    def lin(key, n_out, n_inp, bias=True):
        bound = math.sqrt(6.0 / (n_out + n_inp))
        sd[key + ".weight"] = ((torch.rand(n_out, n_inp, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
        if bias:
            sd[key + ".bias"] = ((torch.rand(n_out, generator=g, dtype=torch.float64) * 2 - 1) * 0.1).to(dtype)

    for i, (sin, vin, sout, vout) in enumerate([(n_in, n_in, nh, nh), (nh, nh, 1, 1)]):
        p = f"equivariant_layers.{i}."
        lin(p + "mix_vectors", 2 * vout, vin, bias=False)
        lin(p + "scalar_net.0", nh, sin + vout)
        lin(p + "scalar_net.1", sout + vout, nh)
    return sd


def dipole_forward(sd, h: Tensor, X: Tensor, pos: Tensor, batch: Tensor, n_mol: int, n_hidden: Optional[int] = None,
                   mean=None, stddev=None, predict_magnitude: bool = False, activation: str = "silu"):
    """Dipole.forward (outputs.py:437-467).  Returns (y [n_mol,3] or its magnitude [n_mol,1], y_vector [n_mol,3,1])."""
    nh = h.shape[1] if n_hidden is None else n_hidden
    l0, l1 = h, X[:, :3, :]                                                         # :449-450
    l0, l1 = gated_equivariant_block(sd, "equivariant_layers.0.", l0, l1, nh, nh, activation, activation)
    l0, l1 = gated_equivariant_block(sd, "equivariant_layers.1.", l0, l1, 1, 1, activation, None)
    if stddev is not None:
        l0 = stddev * l0 + mean                                                     # :456-457
    y = torch.squeeze(l1, -1) + pos * l0                                            # :459-463
    y = torch.zeros(n_mol, 3, dtype=y.dtype).index_add_(0, batch, y)                # :465
    yv = torch.zeros(n_mol, 3, 1, dtype=l1.dtype).index_add_(0, batch, l1)          # :467
    if predict_magnitude:
        y = torch.norm(y, dim=1, keepdim=True)                                      # :470-471
    return y, yv


def spatial_extent_forward(sd_head, h: Tensor, z: Tensor, pos: Tensor, batch: Tensor, n_mol: int, masses: Tensor,
                           activation: str = "ssp"):
    """ElectronicSpatialExtentV2.forward (outputs.py:522-541): the Atomwise MLP alone (no ScaleShift / atomref),
    weighted by the squared distance to the centre of mass.  Returns (y [n_mol,1], x [N,1])."""
    sd_mlp = {k: v for k, v in sd_head.items() if k.startswith("out_net")}
    x = h
    n_lin = len([k for k in sd_mlp if k.endswith(".weight")])
    for i in range(n_lin):
        x = F.linear(x, sd_mlp[f"out_net.1.out_net.{i}.weight"], sd_mlp[f"out_net.1.out_net.{i}.bias"])
        if i < n_lin - 1:
            x = _head_act(activation, x)
    mass = masses.to(pos.dtype)[z].view(-1, 1)                                      # :525
    num = torch.zeros(n_mol, 3, dtype=pos.dtype).index_add_(0, batch, mass * pos)
    den = torch.zeros(n_mol, 1, dtype=pos.dtype).index_add_(0, batch, mass)
    c = num / den                                                                   # :526
    yi = torch.norm(pos - c[batch], dim=1, keepdim=True) ** 2 * x                   # :528-529
    y = torch.zeros(n_mol, 1, dtype=yi.dtype).index_add_(0, batch, yi)              # :531
    return y, x


def synth_batch(kind: str, n_mol: int, seed: int = 0):
    """Deterministic synthetic batches: returns z [N] int64, pos [N,3] f32, batch [N] int64."""
    g = torch.Generator().manual_seed(seed)
    if kind == "qm9":
        n = (18.0 + 2.94 * torch.randn(n_mol, generator=g)).round().clamp(3, 29).long()
        sigma0 = 1.45
    elif kind == "aspirin":
        n = torch.full((n_mol,), 21, dtype=torch.long)
        sigma0 = 1.6
    elif kind == "md22":
        n = torch.full((n_mol,), 370, dtype=torch.long)
        sigma0 = 1.5
    else:
        raise ValueError(kind)
    N = int(n.sum())
    batch = torch.repeat_interleave(torch.arange(n_mol), n)
    sigma = sigma0 * (n.double() / 18.0).pow(1.0 / 3.0).float()
    pos = torch.randn(N, 3, generator=g) * sigma[batch].unsqueeze(-1)
    species = torch.tensor([1, 6, 7, 8, 9])
    probs = torch.tensor([0.51, 0.35, 0.06, 0.078, 0.002])
    z = species[torch.multinomial(probs, N, replacement=True, generator=g)]
    return z, pos, batch
