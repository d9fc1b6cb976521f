"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header
declares (no compute calls), module/state_dict parity with the reference layout, constructor
error behaviour, and that the product fails loudly without a GPU."""
import ctypes
import os

import pytest
import torch

from oracle import gotennet_oracle as orc


@pytest.fixture(scope="module")
def g():
    import gotennet_b200
    return gotennet_b200


def test_library_exports_every_declared_symbol():
    from gotennet_b200 import _build, _lib
    path = _build.lib_path()
    if not os.path.exists(path):
        _build.build()
    protos = _lib.parse_header()
    assert len(protos) >= 30
    cdll = ctypes.CDLL(path)
    for name in protos:
        assert hasattr(cdll, name), name
    assert cdll.goten_abi_version() == 1
    L = _lib.lib()
    assert L.cdll.goten_gemm_workspace_bytes(1792, 256, 300000, 1, 0) > 0


def test_state_dict_layout_matches_reference_spec(g):
    cfg = orc.OracleConfig(n_atom_basis=256, n_interactions=4, lmax=2, sep_dir=True, sep_tensor=True, scale_edge=False)
    m = g.GotenNetWrapper(n_atom_basis=256, n_interactions=4, lmax=2, sep_dir=True, sep_tensor=True, scale_edge=False,
                          cutoff_fn=g.CosineCutoff(5.0), activation="swish")
    sd = orc.expand_aliases(orc.make_state_dict(cfg, 0))
    assert set(sd) == set(m.state_dict())
    assert len(sd) == 127 and sum(p.numel() for p in m.parameters()) == 7_630_080
    m.load_state_dict(sd, strict=True)
    assert m.hidden_dim == 256 and m.cutoff == 5.0 and m.n_interactions == 4 and m.sphere.l == 2
    # aliased MLP registration
    a = m.node_init.W_nrd_nru
    assert a.dense_layers[0].weight is a.layers[0].weight


def test_state_dict_matches_reference_module(g):
    ref_root = "/root/reference"
    if not os.path.isdir(ref_root):
        pytest.skip("reference tree not present on this box")
    from oracle.ref_standins import import_reference
    ref = import_reference(ref_root)
    from gotennet.models.components.layers import CosineCutoff
    for kw in (dict(n_atom_basis=64, n_interactions=2, lmax=1),
               dict(n_atom_basis=32, n_interactions=3, lmax=3, sep_dir=True, sep_tensor=True, sep_htr=False),
               # switches that are no-ops in the reference (no module, no key): "norm", edge_ln with a one-layer gamma_t
               dict(n_atom_basis=32, n_interactions=2, lmax=2, edge_updates="norm_gated", edge_ln="layer"),
               # gamma_w networks (W_edp aliased into the gamma_w Sequential), evec_dim, edge_ln inside a two-layer gamma_t
               dict(n_atom_basis=32, n_interactions=2, lmax=2, edge_updates="linw_ln", evec_dim=24),
               dict(n_atom_basis=32, n_interactions=2, lmax=1, edge_updates="linwa_postln_gated", activation="silu"),
               dict(n_atom_basis=32, n_interactions=3, lmax=2, edge_updates="linw_ln_postln_act_mlp", emlp_dim=40,
                    sep_htr=False, evec_dim=16),
               dict(n_atom_basis=32, n_interactions=2, lmax=2, edge_updates="mlpa", emlp_dim=48, edge_ln="layer")):
        a = ref.GotenNetWrapper(cutoff_fn=CosineCutoff(5.0), **kw).state_dict()
        b = g.GotenNetWrapper(cutoff_fn=g.CosineCutoff(5.0), **kw).state_dict()
        assert set(a) == set(b)
        assert all(a[k].shape == b[k].shape for k in a)
    # read-out heads: identical state_dict layouts and result keys
    from gotennet.models.components import outputs as ro
    for mk_ref, mk_ours in ((lambda: ro.Atomwise(n_in=32, atomref=torch.zeros(100, 1)),
                             lambda: g.Atomwise(n_in=32, atomref=torch.zeros(100, 1))),
                            (lambda: ro.Dipole(n_in=32, predict_magnitude=True), lambda: g.Dipole(n_in=32, predict_magnitude=True)),
                            (lambda: ro.Dipole(n_in=32, n_hidden=24), lambda: g.Dipole(n_in=32, n_hidden=24)),
                            (lambda: ro.ElectronicSpatialExtentV2(n_in=32), lambda: g.ElectronicSpatialExtentV2(n_in=32))):
        a, b = mk_ref().state_dict(), mk_ours().state_dict()
        assert set(a) == set(b), set(a) ^ set(b)
        assert all(a[k].shape == b[k].shape for k in a)


def test_oracle_alias_expansion_of_the_gamma_w_network():
    """W_edp is both an attribute and an element of the gamma_w Sequential (gotennet.py:277-292): the oracle's state dict
    holds it once, expand_aliases adds the position-dependent `gamma_w.k.*` duplicates and needs the config for that."""
    cfg = orc.OracleConfig(n_atom_basis=32, n_interactions=2, lmax=1, num_heads=4, edge_updates="linwa_ln_gated", evec_dim=16)
    sd = orc.make_state_dict(cfg, seed=1)
    assert sd["gata_list.0.W_edp.weight"].shape == (32, 16) and sd["gata_list.0.W_vq.weight"].shape == (16, 32)
    with pytest.raises(ValueError):
        orc.expand_aliases(sd)
    full = orc.expand_aliases(sd, cfg)
    # LayerNorm at 0, the activation at 1, W_edp at 2
    assert full["gata_list.0.gamma_w.2.weight"] is sd["gata_list.0.W_edp.weight"]
    assert "gata_list.0.gamma_w.0.weight" in full and "gata_list.1.W_edp.weight" not in full   # last layer has no HTR
    assert orc.edge_lin_flags(cfg) == (2, 1)


def test_constructor_errors(g):
    with pytest.raises(ValueError):
        g.GotenNet()  # cutoff_fn is mandatory (the reference dies with AttributeError, gotennet.py:839)
    with pytest.raises(ValueError):
        g.GATA(64, torch.nn.functional.silu, edge_updates="bogus")  # gotennet.py:164-167
    with pytest.raises(ValueError):   # gamma_t(t) [E,C] * w [E,evec_dim] cannot broadcast without W_edp (gotennet.py:611)
        g.GATA(64, torch.nn.functional.silu, evec_dim=32)
    with pytest.raises(NotImplementedError):
        g.GATA(64, torch.nn.functional.silu, edge_updates="mlp", edge_ln="batch")
    lw = g.GATA(64, torch.nn.functional.silu, edge_updates="linwa_ln_gatedt", evec_dim=32)
    assert [type(x).__name__ for x in lw.gamma_w] == ["LayerNorm", "SiLU", "Dense", "Tanh"] and lw.gamma_w[2] is lw.W_edp
    assert lw.W_vq.weight.shape == (32, 64) and lw.W_edp.weight.shape == (64, 32) and lw._composed
    assert isinstance(g.GATA(64, torch.nn.functional.silu, edge_updates="gated_gatedt").gamma_w[0], torch.nn.Tanh)
    m = g.GotenNet(n_atom_basis=32, n_interactions=1, cutoff_fn=g.CosineCutoff(5.0), radial_basis="BesselBasis", n_rbf=8)
    assert set(k for k in m.state_dict() if k.startswith("radial_basis.")) == {"radial_basis.freqs", "radial_basis.norm1"}
    m = g.GotenNet(n_atom_basis=32, n_interactions=1, cutoff_fn=g.CosineCutoff(5.0), radial_basis="GaussianRBF", n_rbf=8)
    assert set(k for k in m.state_dict() if k.startswith("radial_basis.")) == {"radial_basis.widths", "radial_basis.offsets"}
    with pytest.raises(ValueError):
        g.GotenNet(cutoff_fn=g.CosineCutoff(5.0), radial_basis="nope")
    with pytest.raises(ValueError):
        g.GotenNet(cutoff_fn=g.CosineCutoff(5.0), radial_basis="gaussianrbf")  # the reference matches this name case-sensitively
    with pytest.raises(ValueError):
        g.GotenNet(cutoff_fn=g.CosineCutoff(5.0), activation="nope")


def test_cpu_tensors_are_rejected(g):
    m = g.GotenNetWrapper(n_atom_basis=32, n_interactions=1, cutoff_fn=g.CosineCutoff(5.0))

    class D:
        pass

    d = D()
    d.z, d.pos, d.batch = torch.ones(3, dtype=torch.long), torch.randn(3, 3), torch.zeros(3, dtype=torch.long)
    with pytest.raises(g.GotenError):
        m(d)


def test_scalar_utilities_match_oracle(g):
    d = torch.linspace(0, 6, 50)
    assert torch.allclose(g.CosineCutoff(5.0)(d), orc.cosine_cutoff(d, 5.0))
    rb = g.ExpNormalSmearing(cutoff=5.0, n_rbf=32)
    means, betas = orc.rbf_buffers(orc.OracleConfig())
    assert torch.equal(rb.means, means) and torch.equal(rb.betas, betas)
    assert torch.allclose(rb(d), orc.expnorm_rbf(d, means, betas, 5.0))
    u = torch.nn.functional.normalize(torch.randn(20, 3), dim=1)
    for l in (1, 2, 3):
        assert torch.allclose(g.TensorInit(l)(u), orc.sph_harm(l, u))


def test_flat_grad_buffer_layout():
    """FlatGradBuffer: every tensor starts on a 16-byte boundary, views alias the flat storage, padding stays zero,
    pack() copies present gradients and zero-fills missing ones (host logic, CPU tensors)."""
    from gotennet_b200.parallel import FlatGradBuffer
    ps = [torch.nn.Parameter(torch.randn(*s)) for s in [(7, 5), (33,), (3,), (64, 64)]]
    frozen = torch.nn.Parameter(torch.randn(5), requires_grad=False)
    fb = FlatGradBuffer(ps + [frozen])
    assert len(fb.params) == 4 and all(o % 4 == 0 for o in fb.offsets) and fb.numel % 4 == 0
    assert fb.numel == 36 + 36 + 4 + 4096
    for p in ps[:3]:
        p.grad = torch.ones_like(p)
    flat = fb.pack()
    assert flat.data_ptr() == fb.flat.data_ptr()
    for o, n, v, p in zip(fb.offsets, fb.sizes, fb.views, ps):
        assert v.data_ptr() == fb.flat.data_ptr() + 4 * o and v.shape == p.shape
        expect = 1.0 if p.grad is not None else 0.0
        assert torch.all(flat[o:o + n] == expect)
    assert float(flat[35]) == 0.0 and float(flat[36 + 33]) == 0.0      # padding
    fb.unpack()
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(ps, fb.views))


def test_amax_scope_inert_without_fp16_arm(monkeypatch):
    """AmaxScope never touches the device when the split-fp16 arm is disabled (GOTEN_GEMM=simt / GOTEN_TC16=0)."""
    from gotennet_b200 import ops
    monkeypatch.setenv("GOTEN_GEMM", "simt")
    sc = ops.AmaxScope()
    assert not sc.enabled and sc.export([torch.zeros(3)]) == []
    monkeypatch.setenv("GOTEN_GEMM", "auto")
    monkeypatch.setenv("GOTEN_TC16", "0")
    assert not ops.AmaxScope().enabled
    monkeypatch.delenv("GOTEN_TC16")
    assert ops.AmaxScope().enabled


def test_collate_ragged_molecules(g):
    mols = [([1, 6, 8], [[0, 0, 0], [1, 0, 0], [0, 1, 0]]), ([6], [[5.0, 5.0, 5.0]]), ([], []), ([7, 7], [[0, 0, 1], [0, 0, 2]])]
    b = g.collate(mols)
    assert b.num_graphs == 4 and b.z.tolist() == [1, 6, 8, 6, 7, 7] and b.pos.shape == (6, 3)
    assert b.batch.tolist() == [0, 0, 0, 1, 3, 3] and b.ptr.tolist() == [0, 3, 4, 4, 6]
    assert b["pos"] is b.pos and b.pos.dtype == torch.float32 and b.z.dtype == torch.int64
    with pytest.raises(ValueError):
        g.collate([([1, 1], [[0, 0, 0]])])
    assert g.collate([]).num_graphs == 0


def test_force_matching_algorithm_on_a_torch_model(g):
    """The central-difference force-matching step (gotennet_b200/training.py) is model-agnostic host logic: on a small
    float64 PyTorch energy model its parameter gradient equals the exact double backward through the forces."""
    torch.manual_seed(0)

    class Rep(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(3, 8).double()

        def forward(self, d):
            r = d.pos - d.pos.mean(0, keepdim=True)
            return torch.tanh(self.lin(r)) * r.pow(2).sum(1, keepdim=True), None

    class Head(torch.nn.Module):
        property, derivative = "y", None

        def __init__(self):
            super().__init__()
            self.out = torch.nn.Linear(8, 1).double()

        def forward(self, d):
            yi = self.out(d.representation)
            return {"y": torch.zeros(2, 1, dtype=yi.dtype).index_add_(0, d.batch, yi)}

    rep, head = Rep(), Head()

    class D:
        pass

    d = D()
    d.z, d.batch = torch.ones(7, dtype=torch.long), torch.tensor([0, 0, 0, 1, 1, 1, 1])
    d.pos = torch.randn(7, 3, dtype=torch.float64)
    E_t, F_t = torch.randn(2, 1, dtype=torch.float64), torch.randn(7, 3, dtype=torch.float64)
    loss_fn = lambda E, F: (E - E_t).pow(2).mean() + 3.0 * (F - F_t).pow(2).mean()  # noqa: E731
    # exact
    pos = d.pos.clone().requires_grad_(True)
    dd = D()
    dd.z, dd.batch, dd.pos = d.z, d.batch, pos
    dd.representation, _ = rep(dd)
    E = head(dd)["y"]
    (gp,) = torch.autograd.grad(E.sum(), pos, create_graph=True)
    loss_fn(E, -gp).backward()
    params = list(rep.parameters()) + list(head.parameters())
    exact = [p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    for order, tol in ((4, 1e-7), (2, 1e-4)):
        loss, E2, F2 = g.force_matching_backward(rep, head, d, loss_fn, h=1e-3, order=order)
        assert torch.allclose(E2, E.detach()) and torch.allclose(F2, -gp.detach())
        for p, e in zip(params, exact):
            assert (p.grad - e).abs().max() <= tol * e.abs().max(), (order, (p.grad - e).abs().max() / e.abs().max())
            p.grad = None
    with pytest.raises(ValueError):
        g.force_matching_backward(rep, head, d, loss_fn, order=3)


YAML_REPRESENTATION = {  # reference configs/model/gotennet.yaml:18-40 as Lightning stores it in hyper_parameters
    "__target__": "gotennet.models.representation.gotennet.GotenNetWrapper",
    "n_atom_basis": 64, "n_interactions": 3, "n_rbf": 32,
    "cutoff_fn": {"__target__": "gotennet.models.components.layers.CosineCutoff", "cutoff": 5.0},
    "radial_basis": "expnorm", "activation": "swish", "max_z": 100, "weight_init": "xavier_uniform",
    "bias_init": "zeros", "num_heads": 8, "attn_dropout": 0.1, "edge_updates": True, "lmax": 2, "aggr": "add",
    "scale_edge": False, "evec_dim": None, "emlp_dim": None, "sep_htr": True, "sep_dir": True, "sep_tensor": True,
    "edge_ln": "",
}


def _lightning_ckpt(path, rep_state, extra_hp=None):
    sd = {"representation." + k: v.clone() for k, v in rep_state.items()}
    sd["output_modules.0.out_net.1.out_net.0.weight"] = torch.zeros(32, 64)   # task head entries are skipped
    sd["output_modules.0.standardize.mean"] = torch.zeros(1)
    hp = {"representation": dict(YAML_REPRESENTATION), "lr": 1e-4, "task": "QM9", "label": "U0"}
    hp.update(extra_hp or {})
    torch.save({"epoch": 3, "global_step": 1234, "pytorch-lightning_version": "2.4.0", "state_dict": sd,
                "hyper_parameters": hp, "optimizer_states": [], "lr_schedulers": []}, path)


def test_load_from_lightning_checkpoint(g, tmp_path):
    """GotenNet.load_from_checkpoint (reference gotennet.py:904-946) on a Lightning-shaped .ckpt: `representation.`
    prefix, `output_modules.*` present, Hydra node with `__target__` and a nested `cutoff_fn` target."""
    cfg = orc.OracleConfig(n_atom_basis=64, n_interactions=3, lmax=2, sep_dir=True, sep_tensor=True, scale_edge=False)
    src = orc.expand_aliases(orc.make_state_dict(cfg, seed=4))
    path = str(tmp_path / "model.ckpt")
    _lightning_ckpt(path, src)
    m = g.GotenNetWrapper.load_from_checkpoint(path)
    assert isinstance(m, g.GotenNetWrapper) and isinstance(m.cutoff_fn, g.CosineCutoff) and m.cutoff == 5.0
    assert m.gata_list[0].dropout == 0.1 and m.sphere.l == 2 and m.hidden_dim == 64
    got = m.state_dict()
    assert set(got) == set(src) and all(torch.equal(got[k], src[k]) for k in src)
    # nested form {"representation": {...}} (gotennet.py:921-922) and the `_target_` spelling
    inner = torch.load(path, weights_only=False)
    inner["hyper_parameters"]["representation"]["_target_"] = inner["hyper_parameters"]["representation"].pop("__target__")
    torch.save({"representation": inner}, path)
    m2 = g.GotenNetWrapper.load_from_checkpoint(path)
    assert all(torch.equal(m2.state_dict()[k], src[k]) for k in src)
    # error behaviour: missing file, missing keys, unknown nested target, stray state entries (strict)
    with pytest.raises(FileNotFoundError):
        g.GotenNet.load_from_checkpoint(str(tmp_path / "absent.ckpt"))
    torch.save({"state_dict": {}}, path)
    with pytest.raises(AssertionError):
        g.GotenNet.load_from_checkpoint(path)
    _lightning_ckpt(path, src)
    bad = torch.load(path, weights_only=False)
    bad["hyper_parameters"]["representation"]["cutoff_fn"]["__target__"] = "os.system"
    torch.save(bad, path)
    with pytest.raises(ValueError):
        g.GotenNetWrapper.load_from_checkpoint(path)
    _lightning_ckpt(path, {**src, "gata_list.0.not_a_key": torch.zeros(1)})
    with pytest.raises(RuntimeError):
        g.GotenNetWrapper.load_from_checkpoint(path)


def test_checkpoint_round_trip_with_reference_module(g, tmp_path):
    """A checkpoint written from the VERBATIM reference module's state_dict loads into ours bit-for-bit, and ours
    loads back into the reference (build container only)."""
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference tree not present on this box")
    from oracle.ref_standins import import_reference
    ref = import_reference("/root/reference")
    from gotennet.models.components.layers import CosineCutoff
    kw = {k: v for k, v in YAML_REPRESENTATION.items() if k not in ("__target__", "cutoff_fn")}
    torch.manual_seed(3)
    rm = ref.GotenNetWrapper(cutoff_fn=CosineCutoff(5.0), **kw)
    path = str(tmp_path / "ref.ckpt")
    _lightning_ckpt(path, rm.state_dict())
    m = g.GotenNetWrapper.load_from_checkpoint(path)
    a, b = rm.state_dict(), m.state_dict()
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
    rm2 = ref.GotenNetWrapper(cutoff_fn=CosineCutoff(5.0), **kw)
    rm2.load_state_dict(m.state_dict(), strict=True)
