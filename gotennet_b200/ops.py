"""Host-side glue: torch tensors -> raw pointers -> C ABI (include/gotennet_b200.h).

PyTorch is used for device memory, streams and autograd bookkeeping only; every
arithmetic step of the path runs in the hand-written sm_100a kernels of
libgotennet_b200.so.  There is no fallback: a missing library or a non-CUDA tensor
raises.

The three autograd.Functions mirror the reference's blocks
(reference representation/gotennet.py): `InitBlockFn` = NodeInit + EdgeInit
(:976-977), `GataBlockFn` = GATA.forward incl. HTR edge update (:366-450),
`EqffBlockFn` = EQFF.forward (:716-748).  Their backward passes are explicit kernel
sequences (no autograd graph inside), so saved activations and workspaces are under
our control.
"""
from __future__ import annotations

import contextlib
import ctypes
import os
from typing import Optional

import torch

from ._lib import GotenError, lib

_WS = {}
# GOTEN_GEMM: auto = split-fp16 -> 3xTF32 -> SIMT (first arm that accepts the shape); tc / tc16 = prefer that tensor-core
# arm and fall back to SIMT (negative impl code); simt = exact-fp32 SIMT only.  Explicit positive codes are strict.
_IMPL = {"simt": 1, "tc": -2, "tc16": -3, "auto": 0}


def gemm_impl() -> int:
    return _IMPL[os.environ.get("GOTEN_GEMM", "auto")]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def device_of(t: torch.Tensor):
    """Context manager making t's CUDA device current (the C ABI launches on the current device's stream); a no-op
    for CPU tensors, which the kernels reject further down with GotenError."""
    return torch.cuda.device(t.device) if t.is_cuda else contextlib.nullcontext()


def _ptr(t: Optional[torch.Tensor], off: int = 0):
    if t is None:
        return None
    return t.data_ptr() + off * t.element_size()


def _chk(*ts):
    cur = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise GotenError("gotennet_b200 kernels need CUDA tensors (there is no CPU path)")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            # the C ABI launches on the current device's stream; a foreign pointer there would fault
            raise GotenError(f"tensor on cuda:{t.device.index} but the current device is cuda:{cur}: call the module "
                             "entry points (they select the device) or wrap the call in torch.cuda.device(...)")
        if not t.is_contiguous():
            raise GotenError("non-contiguous tensor passed to a kernel")


def _f32(t):
    if t.dtype != torch.float32:
        raise GotenError(f"float32 expected, got {t.dtype}")
    return t


def workspace(nbytes: int, device) -> torch.Tensor:
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


# ---------------------------------------------------------------------------
# thin kernel wrappers
# ---------------------------------------------------------------------------
def gemm(A, lda, ta, B, ldb, tb, C, ldc, M, N, K, *, a_off=0, b_off=0, c_off=0, bias=None, add_src=None,
         ld_add=0, add_off=0, act_out=None, ld_act=0, act_off=0, act_lo=0, act_hi=0, colsum=None, impl=None,
         a_amax=None, b_amax=None, am=None):
    """C[M,N] = op(A) op(B) (+bias) (+add_src); see goten_gemm / goten_gemm_scaled in the header.
    a_amax / b_amax: optional 1-element device tensors bounding max|A| / max|B| (split-fp16 arm);
    am: an AmaxScope that supplies (and caches) them for the whole tensors A and B."""
    L = lib()
    if am is not None and am.enabled:
        a_amax = am.of(A) if a_amax is None else a_amax
        b_amax = am.of(B) if b_amax is None else b_amax
    nbytes = L.cdll.goten_gemm_workspace_bytes(M, N, K, ta, tb)
    ws = workspace(nbytes, C.device) if nbytes > 0 else None
    L.call("goten_gemm_scaled", _ptr(A, a_off), lda, ta, _ptr(B, b_off), ldb, tb, _ptr(C, c_off), ldc, M, N, K,
           _ptr(bias), _ptr(add_src, add_off), ld_add, _ptr(act_out, act_off), ld_act, act_lo, act_hi,
           _ptr(colsum), _ptr(a_amax), _ptr(b_amax), _ptr(ws), nbytes, gemm_impl() if impl is None else impl,
           _stream())


def _own(g) -> bool:
    """True when an incoming gradient may be overwritten: a contiguous tensor that is not a view.  Inside the model
    every such tensor is a temporary produced by the next block's backward (or by autograd's accumulation), so the
    residual GEMMs  out = A B + g  accumulate straight into it (C == add_src: TMA reduction-store epilogue) instead
    of reading it through the SMs.  Views (stand-alone module calls return views) take the out-of-place form."""
    return g is not None and g._base is None and g.is_contiguous()


def absmax(A, lda, M, N, *, a_off=0, out=None):
    """1-element device tensor holding max |A[m][n]| (no host sync); accumulates into `out` when given."""
    if out is None:
        out = torch.zeros(1, device=A.device, dtype=torch.float32)
    if M > 0 and N > 0:
        lib().call("goten_absmax", _ptr(A, a_off), lda, M, N, _ptr(out), _stream())
    return out


class AmaxScope:
    """max|T| device scalars of the GEMM operands of one forward / backward call, each measured once
    (goten_absmax, or written by the producing kernel) and shared by every GEMM that reads the tensor;
    `export` / `load` carry them from a block's forward to its backward.  Only the split-fp16 GEMM arm
    uses them (operand scaling), so the scope is inert under GOTEN_GEMM=simt|tc or GOTEN_TC16=0."""

    def __init__(self):
        self.enabled = gemm_impl() in (0, -3) and os.environ.get("GOTEN_TC16", "1") != "0"
        self._d = {}
        self._pool, self._used = None, 0

    def slot(self, device):
        """A zeroed 1-element device tensor (one fill kernel per 32 slots, not per tensor)."""
        if self._pool is None or self._used == self._pool.numel() or self._pool.device != device:
            self._pool, self._used = torch.zeros(32, device=device, dtype=torch.float32), 0
        out = self._pool[self._used:self._used + 1]
        self._used += 1
        return out

    def put(self, T, a):
        if T is not None and a is not None:
            self._d[T.data_ptr()] = (T, a)

    def of(self, T):
        ent = self._d.get(T.data_ptr())
        if ent is None:
            cols = T.shape[-1]
            ent = (T, absmax(T, cols, T.numel() // max(cols, 1), cols, out=self.slot(T.device)))
            self._d[T.data_ptr()] = ent
        return ent[1]

    def prefetch(self, tensors):
        """max |T| of several contiguous tensors (the weight matrices of a block) in ONE launch instead of one pass
        each; tensors already known are skipped."""
        if not self.enabled:
            return
        todo = [T for T in tensors if T is not None and T.data_ptr() not in self._d and T.is_contiguous()][:16]
        if len(todo) < 2:
            return
        dev = todo[0].device
        if self._pool is None or self._used + len(todo) > self._pool.numel() or self._pool.device != dev:
            self._pool, self._used = torch.zeros(32, device=dev, dtype=torch.float32), 0
        out = self._pool[self._used:self._used + len(todo)]
        self._used += len(todo)
        import ctypes
        ptrs = (ctypes.c_void_p * len(todo))(*[T.data_ptr() for T in todo])
        nums = (ctypes.c_int64 * len(todo))(*[T.numel() for T in todo])
        lib().call("goten_absmax_multi", ctypes.addressof(ptrs), ctypes.addressof(nums), len(todo), _ptr(out), _stream())
        for i, T in enumerate(todo):
            self._d[T.data_ptr()] = (T, out[i:i + 1])

    def export(self, tensors):
        """amax tensors (or an empty list when inert) of `tensors`, in order; None entries are skipped."""
        if not self.enabled:
            return []
        return [self.of(T) for T in tensors if T is not None]

    def load(self, tensors, amaxes):
        if not self.enabled or not amaxes:
            return
        for T, a in zip([T for T in tensors if T is not None], amaxes):
            self.put(T, a)


def linear_fwd(a, w, b=None, *, act=False, am=None):
    """z = a w^T + b ; returns (z, silu(z)) if act else z.   a [M,K], w [N,K]."""
    _chk(a, w, b)
    M, K = a.shape
    N = w.shape[0]
    z = torch.empty(M, N, device=a.device, dtype=torch.float32)
    y = torch.empty_like(z) if act else None
    gemm(a, K, 0, w, K, 1, z, N, M, N, K, bias=b, act_out=y, ld_act=N, act_lo=0, act_hi=N if act else 0, am=am)
    return (z, y) if act else z


def linear_bwd(g, a, w, *, need_da=True, need_bias=True, add_src=None, am=None):
    """g [M,N] gradient of z = a w^T + b.  Returns (da [M,K] (+add_src), dw [N,K], db [N])."""
    _chk(g, a, w)
    M, N = g.shape
    K = a.shape[1]
    da = None
    if need_da:
        da = torch.empty(M, K, device=g.device, dtype=torch.float32)
        gemm(g, N, 0, w, K, 0, da, K, M, K, N, add_src=add_src, ld_add=K, am=am)
    dw = torch.empty(N, K, device=g.device, dtype=torch.float32)
    db = torch.empty(N, device=g.device, dtype=torch.float32) if need_bias else None
    gemm(g, N, 1, a, K, 0, dw, K, N, K, M, colsum=db, am=am)
    return da, dw, db


def dsilu_mul(g, ldg, g_off, pre, ldp, p_off, out, ldo, o_off, M, N, out_amax=None):
    """out = g * silu'(pre); out_amax (optional, zeroed 1-element device tensor): running max |out| for a consuming GEMM."""
    lib().call("goten_dsilu_mul", _ptr(g, g_off), ldg, _ptr(pre, p_off), ldp, _ptr(out, o_off), ldo, M, N, _ptr(out_amax),
               _stream())


def permute_nlc(x, to_degree_major: bool):
    """[N,L,C] -> [L,N,C] (to_degree_major) or back."""
    _chk(x)
    if to_degree_major:
        n, l, c = x.shape
        out = torch.empty(l, n, c, device=x.device, dtype=torch.float32)
    else:
        l, n, c = x.shape
        out = torch.empty(n, l, c, device=x.device, dtype=torch.float32)
    lib().call("goten_permute_nlc", _ptr(x), _ptr(out), n, l, c, 1 if to_degree_major else 0, _stream())
    return out


class PermuteFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, to_dm):
        ctx.to_dm = to_dm
        return permute_nlc(x.contiguous(), to_dm)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        return permute_nlc(g.contiguous(), not ctx.to_dm), None


class EmbeddingFn(torch.autograd.Function):
    """table[idx] (gotennet.py:973 A_na, layers.py:1665 A_nbr) with a deterministic backward."""

    @staticmethod
    def forward(ctx, table, idx):
        _chk(table, idx)
        n, C = idx.numel(), table.shape[1]
        out = torch.empty(n, C, device=table.device, dtype=torch.float32)
        lib().call("goten_embedding_fwd", _ptr(table), _ptr(idx), n, C, _ptr(out), _stream())
        ctx.save_for_backward(idx)
        ctx.rows = table.shape[0]
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        g = g.contiguous()
        n, C = g.shape
        nparts = max(1, (n + 63) // 64)
        nbytes = nparts * ctx.rows * C * 4
        ws = torch.empty(nbytes, dtype=torch.uint8, device=g.device)
        out = torch.empty(ctx.rows, C, device=g.device, dtype=torch.float32)
        lib().call("goten_embedding_bwd", _ptr(g), _ptr(idx), n, C, ctx.rows, _ptr(out), _ptr(ws), nbytes, _stream())
        return out, None


# ---------------------------------------------------------------------------
# geometry
# ---------------------------------------------------------------------------
class EdgeGeometryFn(torch.autograd.Function):
    """pos (or caller-supplied edge vectors) -> r, Y, fc, phi, u, kappa.
    components/layers.py:1591-1604, :744-746, :149-152, :805-869; gotennet.py:978-989."""

    @staticmethod
    def forward(ctx, pos, edge_vec, r_in, means, betas, plan, lmax, cutoff, scale_edge, C, basis=0):
        E, L, R = plan.E, (lmax + 1) ** 2 - 1, means.numel()
        dev = plan.src.device
        _chk(pos, edge_vec, r_in, means, betas)
        r = torch.empty(E, device=dev)
        u = torch.empty(E, 3, device=dev)
        Y = torch.empty(E, L, device=dev)
        fc = torch.empty(E, device=dev)
        kappa = torch.empty(E, device=dev)
        phi = torch.empty(E, R, device=dev)
        lib().call("goten_edge_geometry_fwd", _ptr(pos), _ptr(edge_vec), _ptr(r_in), _ptr(plan.src), _ptr(plan.tgt),
                   _ptr(plan.deg_out), E, lmax, float(cutoff), R, int(basis), _ptr(means), _ptr(betas), int(scale_edge), C,
                   _ptr(r), _ptr(u), _ptr(Y), _ptr(fc), _ptr(kappa), _ptr(phi), _stream())
        ctx.plan, ctx.lmax, ctx.cutoff, ctx.basis = plan, lmax, float(cutoff), int(basis)
        ctx.from_pos = pos is not None
        ctx.n_pos = pos.shape[0] if pos is not None else 0
        ctx.save_for_backward(r, u, means, betas)
        ctx.mark_non_differentiable(u, kappa)
        return r, Y, fc, phi, u, kappa

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_r, g_Y, g_fc, g_phi, _gu, _gk):
        r, u, means, betas = ctx.saved_tensors
        plan = ctx.plan
        E = plan.E
        g_vec = torch.empty(E, 3, device=r.device)
        g_Y = g_Y.contiguous() if g_Y is not None else None
        g_fc = g_fc.contiguous() if g_fc is not None else None
        g_phi = g_phi.contiguous() if g_phi is not None else None
        g_r = g_r.contiguous() if g_r is not None else None   # d|v|/dv = u is applied inside the kernel (loops: 0)
        lib().call("goten_edge_geometry_bwd", _ptr(r), _ptr(u), _ptr(plan.src), _ptr(plan.tgt), E, ctx.lmax,
                   ctx.cutoff, means.numel(), ctx.basis, _ptr(means), _ptr(betas), _ptr(g_phi), _ptr(g_fc), _ptr(g_Y),
                   _ptr(g_r), _ptr(g_vec), _stream())
        if ctx.from_pos:
            g_pos = torch.empty(ctx.n_pos, 3, device=r.device)
            lib().call("goten_edge_vec_to_pos_bwd", _ptr(g_vec), _ptr(plan.tgt_ptr), _ptr(plan.src_ptr),
                       _ptr(plan.src_perm), ctx.n_pos, _ptr(g_pos), _stream())
            return g_pos, None, None, None, None, None, None, None, None, None, None
        return None, g_vec, None, None, None, None, None, None, None, None, None


# ---------------------------------------------------------------------------
# init block: NodeInit + EdgeInit  (layers.py:1658-1675, :1704-1714)
# ---------------------------------------------------------------------------
class InitBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h0, hnbr, phi, fc, Wphi, bphi, W1, b1, ln_g, ln_b, W2, b2, plan, ln_eps):
        _chk(h0, hnbr, phi, fc, Wphi, bphi, W1, b1, ln_g, ln_b, W2, b2)
        L_ = lib()
        N, C = h0.shape
        E, R = phi.shape
        dev = h0.device
        st = _stream()
        F = torch.empty(E, 2 * C, device=dev)
        if E > 0:
            gemm(phi, R, 0, Wphi, R, 1, F, 2 * C, E, 2 * C, R, bias=bphi)
        m = torch.empty(N, C, device=dev)
        L_.call("goten_node_init_agg_fwd", _ptr(F), 2 * C, _ptr(hnbr), _ptr(fc), _ptr(plan.tgt_ptr), _ptr(plan.src),
                N, C, _ptr(m), st)
        ctx0 = torch.cat([h0, m], dim=1)
        y1 = linear_fwd(ctx0, W1, b1)
        y2 = torch.empty_like(y1)
        mean = torch.empty(N, device=dev)
        rstd = torch.empty(N, device=dev)
        L_.call("goten_ln_silu_fwd", _ptr(y1), _ptr(ln_g), _ptr(ln_b), N, C, float(ln_eps), _ptr(y2), _ptr(mean),
                _ptr(rstd), st)
        h = linear_fwd(y2, W2, b2)
        t = torch.empty(E, C, device=dev)
        t_amax = torch.zeros(1, device=dev)  # max |t|, written by the kernel (operand scale of the first edge GEMM)
        L_.call("goten_edge_init_fwd", _ptr(h), _ptr(F), 2 * C, C, _ptr(plan.src), _ptr(plan.tgt), E, C, _ptr(t),
                _ptr(t_amax), st)
        ctx.plan = plan
        ctx.save_for_backward(hnbr, phi, fc, Wphi, W1, ln_g, ln_b, W2, F, ctx0, y1, y2, mean, rstd, h)
        ctx.mark_non_differentiable(t_amax)
        return h, t, t_amax

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_h, g_t, _g_amax=None):
        hnbr, phi, fc, Wphi, W1, ln_g, ln_b, W2, F, ctx0, y1, y2, mean, rstd, h = ctx.saved_tensors
        plan = ctx.plan
        L_ = lib()
        N, C = h.shape
        E, R = phi.shape
        dev = h.device
        st = _stream()
        need_geom = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        gF = torch.empty(E, 2 * C, device=dev)
        g_h = g_h.contiguous() if g_h is not None else torch.zeros(N, C, device=dev)
        g_t = g_t.contiguous() if g_t is not None else torch.zeros(E, C, device=dev)
        g_he = torch.empty(N, C, device=dev)
        L_.call("goten_edge_init_bwd", _ptr(g_t), _ptr(h), _ptr(F), 2 * C, C, _ptr(plan.tgt_ptr), _ptr(plan.src),
                _ptr(plan.tgt), _ptr(plan.src_ptr), _ptr(plan.src_perm), N, E, C, _ptr(gF), 2 * C, _ptr(g_he), st)
        g_hh = torch.empty(N, C, device=dev)
        L_.call("goten_add", _ptr(g_h), _ptr(g_he), _ptr(g_hh), N * C, st)
        g_y2, dW2, db2 = linear_bwd(g_hh, y2, W2)
        g_y1 = torch.empty(N, C, device=dev)
        n_part = 296
        gpart = torch.empty(2, n_part, C, device=dev)
        L_.call("goten_ln_silu_bwd", _ptr(g_y2), _ptr(y1), _ptr(ln_g), _ptr(ln_b), _ptr(mean), _ptr(rstd), N, C,
                _ptr(g_y1), _ptr(gpart, 0), _ptr(gpart, n_part * C), n_part, st)
        dln = torch.empty(2, C, device=dev)
        ws = workspace(4 * n_part * C * 4, dev)
        for k in range(2):
            L_.call("goten_colsum", _ptr(gpart, k * n_part * C), C, n_part, C, _ptr(dln, k * C), _ptr(ws),
                    ws.numel(), st)
        g_ctx0, dW1, db1 = linear_bwd(g_y1, ctx0, W1)
        g_h0 = g_ctx0[:, :C].contiguous()
        g_m = g_ctx0[:, C:].contiguous()
        g_fc = torch.zeros(E, device=dev) if need_geom else None
        L_.call("goten_node_init_agg_bwd_tgt", _ptr(g_m), _ptr(F), 2 * C, _ptr(hnbr), _ptr(fc), _ptr(plan.tgt_ptr),
                _ptr(plan.src), N, C, _ptr(gF), 2 * C, _ptr(g_fc), st)
        g_hnbr = torch.empty(N, C, device=dev)
        L_.call("goten_node_init_agg_bwd_src", _ptr(g_m), _ptr(F), 2 * C, _ptr(fc), _ptr(plan.src_ptr),
                _ptr(plan.src_perm), _ptr(plan.tgt), N, C, _ptr(g_hnbr), st)
        if E > 0:
            g_phi, dWphi, dbphi = linear_bwd(gF, phi, Wphi, need_da=need_geom)
        else:
            g_phi = torch.zeros(E, R, device=dev) if need_geom else None
            dWphi, dbphi = torch.zeros_like(Wphi), torch.zeros(2 * C, device=dev)
        return (g_h0, g_hnbr, g_phi, g_fc, dWphi, dbphi, dW1, db1, dln[0], dln[1], dW2, db2, None, None)


# ---------------------------------------------------------------------------
# optional pre-norms of the GATA block (gotennet.py:306-315, :397-398)
# ---------------------------------------------------------------------------
class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm(C) on h [N,C] (layer_norm != "")."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        _chk(x, gamma, beta)
        x = _f32(x.contiguous())
        N, C = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(N, device=x.device)
        rstd = torch.empty(N, device=x.device)
        lib().call("goten_layernorm_fwd", _ptr(x), _ptr(gamma), _ptr(beta), N, C, float(eps), _ptr(y), _ptr(mean),
                   _ptr(rstd), _stream())
        ctx.save_for_backward(x, gamma, beta, mean, rstd)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, gamma, beta, mean, rstd = ctx.saved_tensors
        N, C = x.shape
        dev = x.device
        g = g.contiguous()
        g_x = torch.empty_like(x)
        n_part = 296
        gpart = torch.empty(2, n_part, C, device=dev)
        L_ = lib()
        st = _stream()
        L_.call("goten_layernorm_bwd", _ptr(g), _ptr(x), _ptr(gamma), _ptr(beta), _ptr(mean), _ptr(rstd), N, C,
                _ptr(g_x), _ptr(gpart, 0), _ptr(gpart, n_part * C), n_part, st)
        dln = torch.empty(2, C, device=dev)
        ws = workspace(4 * n_part * C * 4, dev)
        for k in range(2):
            L_.call("goten_colsum", _ptr(gpart, k * n_part * C), C, n_part, C, _ptr(dln, k * C), _ptr(ws), ws.numel(), st)
        return g_x, dln[0], dln[1], None


class TensorLayerNormFn(torch.autograd.Function):
    """TensorLayerNorm (layers.py:1497-1563) on degree-major Xd [L,N,C] (steerable_norm != ""); weight is a buffer."""

    @staticmethod
    def forward(ctx, Xd, weight, lmax):
        _chk(Xd, weight)
        Xd = _f32(Xd.contiguous())
        L, N, C = Xd.shape
        out = torch.empty_like(Xd)
        lib().call("goten_tensor_layernorm_fwd", _ptr(Xd), _ptr(weight), N, C, int(lmax), _ptr(out), _stream())
        ctx.save_for_backward(Xd, weight)
        ctx.lmax = int(lmax)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        Xd, weight = ctx.saved_tensors
        L, N, C = Xd.shape
        g_X = torch.empty_like(Xd)
        lib().call("goten_tensor_layernorm_bwd", _ptr(g.contiguous()), _ptr(Xd), _ptr(weight), N, C, ctx.lmax, _ptr(g_X),
                   _stream())
        return g_X, None, None


# ---------------------------------------------------------------------------
# GATA block: projections + message/softmax/aggregate + HTR edge update
# ---------------------------------------------------------------------------
class GataBlockFn(torch.autograd.Function):
    """Inputs: h [N,C], Xd [L,N,C], t [E,C], geometry (Y [E,L], fc [E], kappa [E]) and the
    layer parameters in concatenated form:
      Wn1 [4C,C]/bn1 = [W_q; W_k; gamma_s.0; gamma_v.0]       (gotennet.py:400-405)
      Ws2 [S*C,C]/bs2 = gamma_s.1,  Wv2/bv2 = gamma_v.1
      We [(S+1|S+2)C, C]/be = [W_re; W_rs; (gamma_t)]          (:406-407, :611)
      Wvq [C,C], Wvk [G,C,C] (G = lmax if sep_htr else 1)     (:432-441), absent on the last layer
    """

    @staticmethod
    def forward(ctx, h, Xd, t, Y, fc, kappa, Wn1, bn1, Ws2, bs2, Wv2, bv2, We, be, Wvq, Wvk, plan, cfg, t_amax=None,
                drop=None, Wt2=None, bt2=None):
        """drop [E,H] (optional): attention dropout factors mask / (1 - p) in plan edge order (gotennet.py:513).
        Wt2 [C,Em] / bt2 (optional): second layer of a two-layer gamma_t ("mlp" / "mlpa" edge updates, gotennet.py:239-250);
        the last Em rows of We are then its first layer."""
        _chk(h, Xd, t, Y, fc, kappa, Wn1, bn1, Ws2, bs2, Wv2, bv2, We, be, Wvq, Wvk, drop, Wt2, bt2)
        L_ = lib()
        st = _stream()
        N, C = h.shape
        L = Xd.shape[0]
        E = t.shape[0]
        H, lmax, S = cfg["H"], cfg["lmax"], cfg["S"]
        htr = Wvq is not None
        ldz = We.shape[0]
        dev = h.device
        SC = S * C
        am = AmaxScope()
        am.put(t, t_amax)  # max |t| from the kernel that produced t (None: measured on first use)
        Wqk = None
        if htr:
            # [W_vq; W_vk,l] stacked per degree group (one GEMM per group below)
            Wqk = torch.cat([Wvq.unsqueeze(0).expand(len(cfg["vk_groups"]), C, C), Wvk], dim=1).contiguous()  # [G][2C][C]
        am.prefetch([Wn1, Ws2, Wv2, We, Wqk, Wt2])  # all weight maxima of the block in one launch
        # node projections: Z1 = [q | k | pre_s | pre_v], A1 = silu(Z1[:, 2C:])
        Z1 = torch.empty(N, 4 * C, device=dev)
        A1 = torch.empty(N, 2 * C, device=dev)
        gemm(h, C, 0, Wn1, C, 1, Z1, 4 * C, N, 4 * C, C, bias=bn1, act_out=A1, ld_act=2 * C, act_lo=2 * C,
             act_hi=4 * C, am=am)
        x = torch.empty(N, SC, device=dev)
        v = torch.empty(N, SC, device=dev)
        gemm(A1, 2 * C, 0, Ws2, C, 1, x, SC, N, SC, C, bias=bs2, am=am)
        gemm(A1, 2 * C, 0, Wv2, C, 1, v, SC, N, SC, C, bias=bv2, a_off=C, am=am)
        # edge projections (pre-activations; consumers apply SiLU)
        Ze = torch.empty(E, ldz, device=dev)
        two = Wt2 is not None                       # two-layer gamma_t
        Em = ldz - (S + 1) * C                      # width of gamma_t's first layer (0 on the last block)
        A_t = torch.empty(E, Em, device=dev) if two else None   # SiLU of its pre-activation, input of the second layer
        h1 = torch.empty_like(h)
        Xd1 = torch.empty_like(Xd)
        alpha = torch.empty(E, H, device=dev)
        hints = torch.zeros(2, device=dev)  # [max |Xd1|, max |t1|], written by the producing kernels
        xd_amax, t1_amax = hints[0:1], hints[1:2]
        am.put(Xd1, xd_amax)
        fused = ctypes.c_int(0)
        if E > 0 and not two and am.enabled:
            # ONE kernel for projections + attention + messages (edge_fused.cu): the projection tile is consumed from
            # tensor memory; Ze is written in full only when a backward will read it, else just gamma_t's columns
            training = any(ctx.needs_input_grad)
            nb = L_.cdll.goten_gata_fused_workspace_bytes(C, ldz)
            ws = workspace(nb, dev)
            L_.call("goten_gata_fused_fwd", _ptr(t), _ptr(We), _ptr(be), _ptr(am.of(t)), _ptr(am.of(We)), _ptr(h),
                    _ptr(Xd), _ptr(Z1), 4 * C, _ptr(x), _ptr(v), _ptr(Y), _ptr(fc), _ptr(kappa), _ptr(drop),
                    _ptr(plan.tgt_ptr), _ptr(plan.src), N, E, C, H, lmax, cfg["gata_flags"], plan.max_deg_in, ldz,
                    0 if training else (S + 1) * C, _ptr(Ze), _ptr(alpha), _ptr(h1), _ptr(Xd1), _ptr(xd_amax), _ptr(ws),
                    ws.numel(), ctypes.addressof(fused), st)
        if not fused.value:
            if E > 0:
                if two:
                    gemm(t, C, 0, We, C, 1, Ze, ldz, E, ldz, C, bias=be, act_out=A_t, ld_act=Em, act_lo=(S + 1) * C,
                         act_hi=ldz, am=am)
                else:
                    gemm(t, C, 0, We, C, 1, Ze, ldz, E, ldz, C, bias=be, am=am)
            L_.call("goten_gata_fwd", _ptr(h), _ptr(Xd), _ptr(Z1), 4 * C, _ptr(x), _ptr(v), _ptr(Ze), ldz, _ptr(Y),
                    _ptr(fc), _ptr(kappa), _ptr(drop), _ptr(plan.tgt_ptr), _ptr(plan.src), N, C, H, lmax,
                    cfg["gata_flags"], plan.max_deg_in, _ptr(h1), _ptr(Xd1), _ptr(alpha), _ptr(xd_amax), st)
        EQK = None
        t1 = t
        if htr:
            # EQ = W_vq X and EK^l = W_vk,l X^l (gotennet.py:432-441) as ONE GEMM per degree group with the stacked
            # weight [W_vq; W_vk,l]: rows of EQK are [EQ | EK] (pitch 2C), X is read once
            G = len(cfg["vk_groups"])
            EQK = torch.empty(L, N, 2 * C, device=dev)
            for g, (lo, hi) in enumerate(cfg["vk_groups"]):
                rows = (hi - lo) * N
                gemm(Xd1, C, 0, Wqk, C, 1, EQK, 2 * C, rows, 2 * C, C, a_off=lo * N * C, b_off=g * 2 * C * C,
                     c_off=lo * N * 2 * C, am=am)
            t1 = torch.empty_like(t)
            Zt, ldt, zt0 = Ze, ldz, (S + 1) * C     # where the kernels read gamma_t's last pre-activation
            if two:
                Zt, ldt, zt0 = torch.empty(E, C, device=dev), C, 0
                if E > 0:
                    gemm(A_t, Em, 0, Wt2, Em, 1, Zt, C, E, C, Em, bias=bt2, am=am)
            L_.call("goten_htr_fwd", _ptr(EQK), _ptr(EQK, C), 2 * C, _ptr(Y), _ptr(Zt), ldt, zt0, _ptr(t),
                    _ptr(plan.tgt_ptr), _ptr(plan.src), N, C, lmax, cfg["htr_flags"], _ptr(t1), _ptr(t1_amax), st)
        ctx.plan, ctx.cfg, ctx.htr = plan, cfg, htr
        amx = am.export([h, t, Wn1, Ws2, Wv2, We, Wqk, A1, Xd1 if htr else None])
        ctx.n_amax = len(amx)
        ctx.has_drop, ctx.two = drop is not None, two
        ctx.save_for_backward(h, Xd, t, Y, fc, kappa, Wn1, Ws2, Wv2, We, Wqk, Z1, A1, x, v, Ze, alpha, Xd1, EQK, *amx,
                              *([drop] if drop is not None else []), *([A_t, Zt, Wt2] if two else []))
        ctx.mark_non_differentiable(hints)
        if htr:
            return h1, Xd1, t1, hints
        return h1, Xd1, hints  # last layer: t_ij passes through unchanged (gotennet.py:449-450)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_h1, g_Xd1, g_t1=None, _g_hints=None):
        saved = ctx.saved_tensors
        if not ctx.htr:
            g_t1 = None  # (third output of the last layer is the hint tensor)
        (h, Xd, t, Y, fc, kappa, Wn1, Ws2, Wv2, We, Wqk, Z1, A1, x, v, Ze, alpha, Xd1, EQK) = saved[:19]
        plan, cfg, htr = ctx.plan, ctx.cfg, ctx.htr
        am = AmaxScope()
        am.load([h, t, Wn1, Ws2, Wv2, We, Wqk, A1, Xd1 if htr else None], saved[19:19 + ctx.n_amax])
        drop = saved[19 + ctx.n_amax] if ctx.has_drop else None
        A_t, Zt2, Wt2 = saved[-3:] if ctx.two else (None, None, None)
        L_ = lib()
        st = _stream()
        N, C = h.shape
        L = Xd.shape[0]
        E = t.shape[0]
        H, lmax, S = cfg["H"], cfg["lmax"], cfg["S"]
        ldz = We.shape[0]
        SC = S * C
        dev = h.device
        need_gY, need_gfc = ctx.needs_input_grad[3], ctx.needs_input_grad[4]
        g_h1 = g_h1.contiguous() if g_h1 is not None else torch.zeros(N, C, device=dev)
        g_Xd1 = g_Xd1.contiguous() if g_Xd1 is not None else torch.zeros(L, N, C, device=dev)
        if g_t1 is not None:
            g_t1 = g_t1.contiguous()
        gZe = torch.empty(E, ldz, device=dev)
        gze_amax = am.slot(dev) if am.enabled else None  # written by the two kernels that fill gZe
        am.put(gZe, gze_amax)
        g_Y = torch.zeros(E, L, device=dev) if need_gY else None
        g_fc = torch.zeros(E, device=dev) if need_gfc else None
        dWvq = dWvk = dWt2 = dbt2 = None
        g_Xm = g_Xd1  # gradient reaching the post-message X
        if htr:
            if g_t1 is None:
                g_t1 = torch.zeros(E, C, device=dev)
            g_EQK = torch.empty_like(EQK)  # rows [g_EQ | g_EK], pitch 2C
            geqk_amax = am.slot(dev) if am.enabled else None
            am.put(g_EQK, geqk_amax)
            zt0 = (S + 1) * C
            Zt, ldt, zc0, gZt, ldgt, gzt_amax = Ze, ldz, zt0, gZe, ldz, gze_amax
            if ctx.two:  # gamma_t's last pre-activation and its gradient live in their own [E, C] buffers
                Zt, ldt, zc0 = Zt2, C, 0
                gZt, ldgt = torch.empty(E, C, device=dev), C
                gzt_amax = am.slot(dev) if am.enabled else None
                am.put(gZt, gzt_amax)
            L_.call("goten_htr_bwd_tgt", _ptr(g_t1), _ptr(EQK), _ptr(EQK, C), 2 * C, _ptr(Y), _ptr(Zt), ldt, zc0,
                    _ptr(plan.tgt_ptr), _ptr(plan.src), N, C, lmax, cfg["htr_flags"], _ptr(g_EQK), _ptr(gZt), ldgt,
                    _ptr(g_Y), _ptr(gzt_amax), _ptr(geqk_amax), st)
            L_.call("goten_htr_bwd_src", _ptr(g_t1), _ptr(EQK), _ptr(EQK, C), 2 * C, _ptr(Y), _ptr(Zt), ldt, zc0,
                    _ptr(plan.src_ptr), _ptr(plan.src_perm), _ptr(plan.tgt), N, C, lmax, cfg["htr_flags"],
                    _ptr(g_EQK, C), _ptr(geqk_amax), st)
            if ctx.two:  # back through gamma_t's second layer into the first layer's columns of gZe
                Em = ldz - zt0
                g_At = torch.empty(E, Em, device=dev)
                dWt2 = torch.empty_like(Wt2)
                dbt2 = torch.empty(C, device=dev)
                if E > 0:
                    gemm(gZt, C, 0, Wt2, Em, 0, g_At, Em, E, Em, C, am=am)
                    gemm(gZt, C, 1, A_t, Em, 0, dWt2, Em, C, Em, E, colsum=dbt2, am=am)
                    dsilu_mul(g_At, Em, 0, Ze, ldz, zt0, gZe, ldz, zt0, E, Em)
                    if am.enabled:
                        absmax(gZe, ldz, E, Em, a_off=zt0, out=gze_amax)
                else:
                    dWt2.zero_()
                    dbt2.zero_()
            # X gradient: g_Xm^l = g_Xd1^l + [g_EQ | g_EK]^l [W_vq; W_vk,l] (one K = 2C GEMM per degree group, the
            # residual added once); weight gradients of the stacked weight, un-stacked below
            G = len(cfg["vk_groups"])
            g_Xm = g_Xd1 if _own(g_Xd1) else torch.empty_like(Xd1)   # in place: g_Xd1 is not read again
            dWqk = torch.empty_like(Wqk)
            for g, (lo, hi) in enumerate(cfg["vk_groups"]):
                rows, off = (hi - lo) * N, lo * N * C
                gemm(g_EQK, 2 * C, 0, Wqk, C, 0, g_Xm, C, rows, C, 2 * C, a_off=2 * off, b_off=g * 2 * C * C, c_off=off,
                     add_src=g_Xd1, ld_add=C, add_off=off, am=am)
                gemm(g_EQK, 2 * C, 1, Xd1, C, 0, dWqk, C, 2 * C, C, rows, a_off=2 * off, b_off=off,
                     c_off=g * 2 * C * C, am=am)
            # W_vq is shared by the degree groups: its gradient is the sum of the groups' [C, C] blocks (goten_add)
            dWvq = dWqk[0, :C].contiguous()
            for g in range(1, G):
                L_.call("goten_add", _ptr(dWvq), _ptr(dWqk, g * 2 * C * C), _ptr(dWvq), C * C, st)
            dWvk = dWqk[:, C:].contiguous()
        # message block
        g_Z1 = torch.empty(N, 4 * C, device=dev)
        da = torch.empty(E, H, device=dev)
        L_.call("goten_gata_bwd_tgt", _ptr(g_h1), _ptr(g_Xm), _ptr(Xd), _ptr(Z1), 4 * C, _ptr(x), _ptr(v), _ptr(Ze),
                ldz, _ptr(Y), _ptr(fc), _ptr(kappa), _ptr(drop), _ptr(alpha), _ptr(plan.tgt_ptr), _ptr(plan.src), N, C, H,
                lmax,
                cfg["gata_flags"], plan.max_deg_in, _ptr(g_Z1), 4 * C, _ptr(gZe), ldz, _ptr(da), _ptr(g_fc),
                _ptr(g_Y), _ptr(gze_amax), st)
        g_x = torch.empty(N, SC, device=dev)
        g_v = torch.empty(N, SC, device=dev)
        gx_amax = am.slot(dev) if am.enabled else None   # written by the kernel that fills g_x / g_v
        gv_amax = am.slot(dev) if am.enabled else None
        am.put(g_x, gx_amax)
        am.put(g_v, gv_amax)
        g_Xd = torch.empty_like(Xd)
        L_.call("goten_gata_bwd_src", _ptr(g_h1), _ptr(g_Xm), _ptr(Xd), _ptr(Z1), 4 * C, _ptr(x), _ptr(v), _ptr(Ze),
                ldz, _ptr(Y), _ptr(fc), _ptr(kappa), _ptr(drop), _ptr(alpha), _ptr(da), _ptr(plan.src_ptr),
                _ptr(plan.src_perm),
                _ptr(plan.tgt), N, C, H, lmax, cfg["gata_flags"], _ptr(g_Z1), 4 * C, _ptr(g_x), _ptr(g_v),
                _ptr(g_Xd), _ptr(gx_amax), _ptr(gv_amax), st)
        # gamma_s.1 / gamma_v.1
        g_A1 = torch.empty(N, 2 * C, device=dev)
        gemm(g_x, SC, 0, Ws2, C, 0, g_A1, 2 * C, N, C, SC, am=am)
        gemm(g_v, SC, 0, Wv2, C, 0, g_A1, 2 * C, N, C, SC, c_off=C, am=am)
        dWs2 = torch.empty_like(Ws2)
        dbs2 = torch.empty(SC, device=dev)
        gemm(g_x, SC, 1, A1, 2 * C, 0, dWs2, C, SC, C, N, colsum=dbs2, am=am)
        dWv2 = torch.empty_like(Wv2)
        dbv2 = torch.empty(SC, device=dev)
        gemm(g_v, SC, 1, A1, 2 * C, 0, dWv2, C, SC, C, N, b_off=C, colsum=dbv2, am=am)
        # through the SiLU of gamma_s.0 / gamma_v.0 into g_Z1[:, 2C:4C]
        dsilu_mul(g_A1, 2 * C, 0, Z1, 4 * C, 2 * C, g_Z1, 4 * C, 2 * C, N, 2 * C)
        g_h = g_h1 if _own(g_h1) else torch.empty_like(h)           # (the graph kernels above were its last readers)
        gemm(g_Z1, 4 * C, 0, Wn1, C, 0, g_h, C, N, C, 4 * C, add_src=g_h1, ld_add=C, am=am)
        dWn1 = torch.empty_like(Wn1)
        dbn1 = torch.empty(4 * C, device=dev)
        gemm(g_Z1, 4 * C, 1, h, C, 0, dWn1, C, 4 * C, C, N, colsum=dbn1, am=am)
        # edge projections
        g_t = g_t1 if (_own(g_t1) and E > 0) else torch.empty_like(t)
        dWe = torch.empty_like(We)
        dbe = torch.empty(ldz, device=dev)
        if E > 0:
            gemm(gZe, ldz, 0, We, C, 0, g_t, C, E, C, ldz, add_src=g_t1, ld_add=C, am=am)
            gemm(gZe, ldz, 1, t, C, 0, dWe, C, ldz, C, E, colsum=dbe, am=am)
        else:
            dWe.zero_()
            dbe.zero_()
        return (g_h, g_Xd, g_t, g_Y, g_fc, None, dWn1, dbn1, dWs2, dbs2, dWv2, dbv2, dWe, dbe, dWvq, dWvk, None, None,
                None, None, dWt2 if ctx.two else None, dbt2 if ctx.two else None)


# ---------------------------------------------------------------------------
# EQFF block (gotennet.py:728-748)
# ---------------------------------------------------------------------------
class EqffBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, Xd, Wvu, Wm1, bm1, Wm2, bm2, eps, xd_amax=None):
        _chk(h, Xd, Wvu, Wm1, bm1, Wm2, bm2)
        L_ = lib()
        st = _stream()
        N, C = h.shape
        L = Xd.shape[0]
        dev = h.device
        am = AmaxScope()
        am.put(Xd, xd_amax)  # max |Xd| from the GATA kernel that produced it (None: measured on first use)
        am.prefetch([Wvu, Wm1, Wm2])
        P = torch.empty_like(Xd)
        gemm(Xd, C, 0, Wvu, C, 1, P, C, L * N, C, C, am=am)
        cx = torch.empty(N, 2 * C, device=dev)
        cx_amax = am.slot(dev) if am.enabled else None   # written by the kernel that fills cx
        am.put(cx, cx_amax)
        L_.call("goten_eqff_ctx_fwd", _ptr(h), _ptr(P), N, C, L, float(eps), _ptr(cx), _ptr(cx_amax), st)
        Zm, Am = linear_fwd(cx, Wm1, bm1, act=True, am=am)
        M = linear_fwd(Am, Wm2, bm2, am=am)
        h2 = torch.empty_like(h)
        Xd2 = torch.empty_like(Xd)
        L_.call("goten_eqff_update_fwd", _ptr(h), _ptr(Xd), _ptr(P), _ptr(M), N, C, L, _ptr(h2), _ptr(Xd2), st)
        amx = am.export([Xd, Wvu, Wm1, Wm2, cx, Am])
        ctx.n_amax = len(amx)
        ctx.save_for_backward(Xd, Wvu, Wm1, Wm2, P, cx, Zm, Am, M, *amx)
        return h2, Xd2

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_h2, g_Xd2):
        saved = ctx.saved_tensors
        Xd, Wvu, Wm1, Wm2, P, cx, Zm, Am, M = saved[:9]
        am = AmaxScope()
        am.load([Xd, Wvu, Wm1, Wm2, cx, Am], saved[9:9 + ctx.n_amax])
        L_ = lib()
        st = _stream()
        L, N, C = Xd.shape
        dev = Xd.device
        g_h2 = g_h2.contiguous() if g_h2 is not None else torch.zeros(N, C, device=dev)
        g_Xd2 = g_Xd2.contiguous() if g_Xd2 is not None else torch.zeros(L, N, C, device=dev)
        g_M = torch.empty(N, 2 * C, device=dev)
        gm_amax = am.slot(dev) if am.enabled else None
        am.put(g_M, gm_amax)
        L_.call("goten_eqff_update_bwd", _ptr(g_h2), _ptr(g_Xd2), _ptr(P), N, C, L, _ptr(g_M), _ptr(gm_amax), st)
        g_Am, dWm2, dbm2 = linear_bwd(g_M, Am, Wm2, am=am)
        g_Zm = torch.empty_like(g_Am)
        gzm_amax = am.slot(dev) if am.enabled else None
        am.put(g_Zm, gzm_amax)
        dsilu_mul(g_Am, C, 0, Zm, C, 0, g_Zm, C, 0, N, C, out_amax=gzm_amax)
        g_cx, dWm1, dbm1 = linear_bwd(g_Zm, cx, Wm1, am=am)
        g_P = torch.empty_like(P)
        g_h = torch.empty(N, C, device=dev)
        gp_amax = am.slot(dev) if am.enabled else None
        am.put(g_P, gp_amax)
        L_.call("goten_eqff_ctx_bwd", _ptr(g_h2), _ptr(g_Xd2), _ptr(g_cx), _ptr(P), _ptr(M), _ptr(cx), N, C, L,
                _ptr(g_P), _ptr(g_h), _ptr(gp_amax), st)
        g_Xd = g_Xd2 if _own(g_Xd2) else torch.empty_like(Xd)
        gemm(g_P, C, 0, Wvu, C, 0, g_Xd, C, L * N, C, C, add_src=g_Xd2, ld_add=C, am=am)
        dWvu = torch.empty_like(Wvu)
        gemm(g_P, C, 1, Xd, C, 0, dWvu, C, C, C, L * N, am=am)
        return g_h, g_Xd, dWvu, dWm1, dbm1, dWm2, dbm2, None, None


# ---------------------------------------------------------------------------
# read-out head (reference models/components/outputs.py:232-376)
# ---------------------------------------------------------------------------
ACT_NONE, ACT_SILU, ACT_SSP, ACT_SIGMOID, ACT_TANH = 0, 1, 2, 3, 4


class DenseActFn(torch.autograd.Function):
    """y = act(x W^T + b) for the head MLP (components/layers.py:225-273 SchnetMLP of Dense layers).
    SiLU rides in the GEMM epilogue; shifted softplus (layers.py:69-81) is one element-wise kernel."""

    @staticmethod
    def forward(ctx, x, w, b, kind):
        _chk(x, w, b)
        x, w = _f32(x.contiguous()), w.contiguous()
        ctx.kind, ctx.has_bias = kind, b is not None
        if kind == ACT_SILU:
            z, y = linear_fwd(x, w, b, act=True)
        else:
            z = linear_fwd(x, w, b)
            y = z
            if kind == ACT_SSP:
                y = torch.empty_like(z)
                lib().call("goten_act_fwd", kind, _ptr(z), z.numel(), _ptr(y), _stream())
        ctx.save_for_backward(x, w, z)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, w, z = ctx.saved_tensors
        g = g.contiguous()
        if ctx.kind != ACT_NONE:
            gz = torch.empty_like(g)
            lib().call("goten_act_bwd", ctx.kind, _ptr(g), _ptr(z), g.numel(), _ptr(gz), _stream())
            g = gz
        da, dw, db = linear_bwd(g, x, w, need_da=ctx.needs_input_grad[0], need_bias=ctx.has_bias)
        return da, dw, db, None


class ActFn(torch.autograd.Function):
    """Element-wise activation (goten_act_*): SiLU / shifted softplus / sigmoid / tanh."""

    @staticmethod
    def forward(ctx, x, kind):
        _chk(x)
        x = _f32(x.contiguous())
        ctx.kind = kind
        ctx.save_for_backward(x)
        return _act_apply(kind, x)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = g.contiguous()
        out = torch.empty_like(g)
        lib().call("goten_act_bwd", ctx.kind, _ptr(g), _ptr(x), g.numel(), _ptr(out), _stream())
        return out, None


class MulAddFn(torch.autograd.Function):
    """a * b + c, all [E, C]: the residual edge update t + gamma_t(t) * gamma_w(w) (gotennet.py:611, :445) of the
    host-composed HTR variants."""

    @staticmethod
    def forward(ctx, a, b, c):
        _chk(a, b, c)
        a, b, c = a.contiguous(), b.contiguous(), c.contiguous()
        out = torch.empty_like(a)
        lib().call("goten_mul_add_fwd", _ptr(a), _ptr(b), _ptr(c), a.numel(), _ptr(out), _stream())
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = g.contiguous()
        g_a, g_b = torch.empty_like(a), torch.empty_like(b)
        lib().call("goten_mul_add_bwd", _ptr(g), _ptr(a), _ptr(b), g.numel(), _ptr(g_a), _ptr(g_b), _stream())
        return g_a, g_b, g


class HtrWeightFn(torch.autograd.Function):
    """w_ij [E, Ev] = sum over degree groups of rej(EQ_i) . rej(EK_j) (GATA.edge_update up to the gamma_w call,
    gotennet.py:580-609) from degree-major EQ / EK [L, N, Ev]: the weight-only mode of the HTR kernels (flags bit 5).
    Gradients: g_EQ, g_EK and, when the geometry needs it (forces), g_Y."""

    @staticmethod
    def forward(ctx, EQ, EK, Y, plan, lmax, flags):
        _chk(EQ, EK, Y)
        EQ, EK, Y = EQ.contiguous(), EK.contiguous(), Y.contiguous()
        L, N, Ev = EQ.shape
        w = torch.empty(plan.E, Ev, device=EQ.device)
        ctx.plan, ctx.lmax, ctx.flags = plan, lmax, flags | 32
        lib().call("goten_htr_fwd", _ptr(EQ), _ptr(EK), Ev, _ptr(Y), None, 0, 0, None, _ptr(plan.tgt_ptr),
                   _ptr(plan.src), N, Ev, lmax, ctx.flags, _ptr(w), None, _stream())
        ctx.save_for_backward(EQ, EK, Y)
        return w

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_w):
        EQ, EK, Y = ctx.saved_tensors
        plan = ctx.plan
        L, N, Ev = EQ.shape
        g_w = g_w.contiguous()
        g_EQ, g_EK = torch.empty_like(EQ), torch.empty_like(EK)
        g_Y = torch.zeros_like(Y) if ctx.needs_input_grad[2] else None
        L_, st = lib(), _stream()
        L_.call("goten_htr_bwd_tgt", _ptr(g_w), _ptr(EQ), _ptr(EK), Ev, _ptr(Y), None, 0, 0, _ptr(plan.tgt_ptr),
                _ptr(plan.src), N, Ev, ctx.lmax, ctx.flags, _ptr(g_EQ), None, 0, _ptr(g_Y), None, None, st)
        L_.call("goten_htr_bwd_src", _ptr(g_w), _ptr(EQ), _ptr(EK), Ev, _ptr(Y), None, 0, 0, _ptr(plan.src_ptr),
                _ptr(plan.src_perm), _ptr(plan.tgt), N, Ev, ctx.lmax, ctx.flags, _ptr(g_EK), None, st)
        return g_EQ, g_EK, g_Y, None, None, None


def mol_ptr_from_batch(batch: torch.Tensor, n_mol: int, check_sorted: bool = True) -> torch.Tensor:
    """[n_mol+1] int32 atom offsets of the molecules of a (sorted) PyG batch vector."""
    _chk(batch)
    batch = batch.contiguous().long()
    mol_ptr = torch.empty(n_mol + 1, dtype=torch.int32, device=batch.device)
    flag = torch.empty(1, dtype=torch.int32, device=batch.device) if check_sorted else None
    lib().call("goten_mol_ptr", _ptr(batch), batch.numel(), n_mol, _ptr(mol_ptr), _ptr(flag), _stream())
    if check_sorted and int(flag.item()) != 0:
        raise GotenError("the batch vector must be sorted by molecule id (PyG batches are)")
    return mol_ptr


class AtomwiseReduceFn(torch.autograd.Function):
    """raw [N,n_out] -> yi = raw*stddev + mean (+ atomref[z]);  y = segment sum / mean over molecules
    (outputs.py:347-355).  mode: 0 none, 1 sum, 2 mean."""

    @staticmethod
    def forward(ctx, raw, z, atomref, mean, stddev, mol_ptr, n_mol, mode):
        _chk(raw, z, atomref, mean, stddev, mol_ptr)
        raw = _f32(raw.contiguous())
        N, n_out = raw.shape
        dev = raw.device
        n_stat = stddev.numel() if stddev is not None else (mean.numel() if mean is not None else 1)
        yi = torch.empty_like(raw)
        y = torch.zeros(n_mol, n_out, device=dev) if mode != 0 else None
        lib().call("goten_atomwise_reduce_fwd", _ptr(raw), _ptr(z), _ptr(atomref),
                   atomref.shape[0] if atomref is not None else 0, _ptr(mean), _ptr(stddev), n_stat, _ptr(mol_ptr), N,
                   n_mol, n_out, mode, _ptr(yi), _ptr(y), _stream())
        ctx.mode, ctx.n_mol, ctx.n_stat, ctx.N = mode, n_mol, n_stat, N
        ctx.save_for_backward(stddev, mol_ptr)
        if mode == 0:
            return yi, yi
        return yi, y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_yi, g_y):
        stddev, mol_ptr = ctx.saved_tensors
        ref = g_yi if g_yi is not None else g_y
        N = ctx.N
        n_out = ref.shape[1]
        g_raw = torch.zeros(N, n_out, device=ref.device)
        lib().call("goten_atomwise_reduce_bwd", _ptr(g_y.contiguous() if g_y is not None else None),
                   _ptr(g_yi.contiguous() if g_yi is not None else None), _ptr(stddev), ctx.n_stat, _ptr(mol_ptr), N,
                   ctx.n_mol, n_out, ctx.mode, _ptr(g_raw), _stream())
        return g_raw, None, None, None, None, None, None, None


# ---------------------------------------------------------------------------
# equivariant read-out heads (reference models/components/outputs.py:24-104, :379-542)
# ---------------------------------------------------------------------------
def _act_apply(kind, z):
    if kind == ACT_NONE:
        return z
    y = torch.empty_like(z)
    lib().call("goten_act_fwd", kind, _ptr(z), z.numel(), _ptr(y), _stream())
    return y


class GatedEquivariantFn(torch.autograd.Function):
    """GatedEquivariantBlock.forward (outputs.py:76-103): scalars [N,ns], vectors [N,3,nvin] ->
    (s_out [N,nso], v_out [N,3,nv]).  Three GEMMs (mix_vectors, scalar_net.0 with its activation, scalar_net.1) around
    two element-wise kernels (vector norms into the context, gating)."""

    @staticmethod
    def forward(ctx, scalars, vectors, Wmix, W1, b1, W2, b2, nso, nv, act, sact):
        scalars, vectors = _f32(scalars.contiguous()), _f32(vectors.contiguous())  # X[:, :3] arrives as a strided view
        _chk(scalars, vectors, Wmix, W1, b1, W2, b2)
        N, ns = scalars.shape
        nvin = vectors.shape[2]
        if vectors.shape[:2] != (N, 3) or Wmix.shape != (2 * nv, nvin) or W2.shape[0] != nso + nv:
            raise GotenError("GatedEquivariantBlock: inconsistent shapes")
        dev = scalars.device
        L_, st = lib(), _stream()
        vmix = torch.empty(N, 3, 2 * nv, device=dev)
        if N > 0:
            gemm(vectors, nvin, 0, Wmix, nvin, 1, vmix, 2 * nv, 3 * N, 2 * nv, nvin)
        cx = torch.empty(N, ns + nv, device=dev)
        L_.call("goten_geb_ctx_fwd", _ptr(scalars), _ptr(vmix), N, ns, nv, _ptr(cx), st)
        if act == ACT_SILU:
            Z1, A1 = linear_fwd(cx, W1, b1, act=True)
        else:
            Z1 = linear_fwd(cx, W1, b1)
            A1 = _act_apply(act, Z1)
        x = linear_fwd(A1, W2, b2)
        s_out = torch.empty(N, nso, device=dev)
        v_out = torch.empty(N, 3, nv, device=dev)
        L_.call("goten_geb_gate_fwd", _ptr(x), _ptr(vmix), N, nso, nv, sact, _ptr(s_out), _ptr(v_out), st)
        ctx.dims = (N, ns, nvin, nso, nv, act, sact)
        ctx.save_for_backward(vectors, Wmix, W1, W2, vmix, cx, Z1, A1, x)
        return s_out, v_out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_s, g_v):
        vectors, Wmix, W1, W2, vmix, cx, Z1, A1, x = ctx.saved_tensors
        N, ns, nvin, nso, nv, act, sact = ctx.dims
        dev = x.device
        L_, st = lib(), _stream()
        g_s = g_s.contiguous() if g_s is not None else None
        g_v = g_v.contiguous() if g_v is not None else None
        g_x = torch.empty_like(x)
        g_vmix = torch.empty_like(vmix)
        L_.call("goten_geb_gate_bwd", _ptr(g_s), _ptr(g_v), _ptr(x), _ptr(vmix), N, nso, nv, sact, _ptr(g_x), _ptr(g_vmix),
                st)
        g_A1, dW2, db2 = linear_bwd(g_x, A1, W2)
        if act != ACT_NONE:
            g_Z1 = torch.empty_like(g_A1)
            L_.call("goten_act_bwd", act, _ptr(g_A1), _ptr(Z1), g_A1.numel(), _ptr(g_Z1), st)
        else:
            g_Z1 = g_A1
        g_cx, dW1, db1 = linear_bwd(g_Z1, cx, W1)
        g_sc = torch.empty(N, ns, device=dev)
        L_.call("goten_geb_ctx_bwd", _ptr(g_cx), _ptr(vmix), N, ns, nv, _ptr(g_sc), _ptr(g_vmix), st)
        g_vec = torch.empty_like(vectors)
        dWmix = torch.empty_like(Wmix)
        if N > 0:
            gemm(g_vmix, 2 * nv, 0, Wmix, nvin, 0, g_vec, nvin, 3 * N, nvin, 2 * nv)
            gemm(g_vmix, 2 * nv, 1, vectors, nvin, 0, dWmix, nvin, 2 * nv, nvin, 3 * N)
        else:
            dWmix.zero_()
        return g_sc, g_vec, dWmix, dW1, db1, dW2, db2, None, None, None, None


class DipoleAtomFn(torch.autograd.Function):
    """yi [N,3] = atomic dipoles + pos * charges, charges = stddev * q + mean when standardised (outputs.py:446-453)."""

    @staticmethod
    def forward(ctx, l1, l0, pos, sd, mu):
        l1, l0, pos = _f32(l1.contiguous()), _f32(l0.contiguous()), _f32(pos.contiguous())
        _chk(l1, l0, pos)
        N = l0.numel()
        yi = torch.empty(N, 3, device=l1.device)
        lib().call("goten_dipole_atom_fwd", _ptr(l1), _ptr(l0), _ptr(pos), float(sd), float(mu), N, _ptr(yi), _stream())
        ctx.stats = (float(sd), float(mu))
        ctx.save_for_backward(l0, pos)
        return yi

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        l0, pos = ctx.saved_tensors
        sd, mu = ctx.stats
        N = l0.numel()
        g = g.contiguous()
        g_l1 = torch.empty(N, 3, device=g.device)
        g_l0 = torch.empty_like(l0)
        g_pos = torch.empty(N, 3, device=g.device) if ctx.needs_input_grad[2] else None
        lib().call("goten_dipole_atom_bwd", _ptr(g), _ptr(l0), _ptr(pos), sd, mu, N, _ptr(g_l1), _ptr(g_l0), _ptr(g_pos),
                   _stream())
        return g_l1, g_l0, g_pos, None, None


class RowNormFn(torch.autograd.Function):
    """||v[m,:]||_2 with keepdim (predict_magnitude, outputs.py:460-461)."""

    @staticmethod
    def forward(ctx, v):
        v = _f32(v.contiguous())
        _chk(v)
        M, D = v.shape
        y = torch.empty(M, 1, device=v.device)
        lib().call("goten_rownorm_fwd", _ptr(v), M, D, _ptr(y), _stream())
        ctx.save_for_backward(v)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (v,) = ctx.saved_tensors
        g_v = torch.empty_like(v)
        lib().call("goten_rownorm_bwd", _ptr(g.contiguous()), _ptr(v), v.shape[0], v.shape[1], _ptr(g_v), _stream())
        return g_v


class SpatialExtentFn(torch.autograd.Function):
    """ElectronicSpatialExtentV2.forward (outputs.py:522-541): x [N,1], pos, z, mass table -> y [n_mol,1]."""

    @staticmethod
    def forward(ctx, x, pos, z, mass, mol_ptr, n_mol):
        x, pos = _f32(x.contiguous()), _f32(pos.contiguous())
        _chk(x, pos, z, mass, mol_ptr)
        N = x.shape[0]
        dev = x.device
        yi = torch.empty(N, 1, device=dev)
        y = torch.zeros(n_mol, 1, device=dev)
        cen = torch.zeros(max(n_mol, 1), 4, device=dev)
        lib().call("goten_ese_fwd", _ptr(x), _ptr(pos), _ptr(z), _ptr(mass), mass.numel(), _ptr(mol_ptr), n_mol, _ptr(yi),
                   _ptr(y), _ptr(cen), _stream())
        ctx.n_mol = n_mol
        ctx.save_for_backward(x, pos, z, mass, mol_ptr, cen)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, pos, z, mass, mol_ptr, cen = ctx.saved_tensors
        g_x = torch.empty_like(x)
        g_pos = torch.empty_like(pos) if ctx.needs_input_grad[1] else None
        lib().call("goten_ese_bwd", _ptr(g.contiguous()), _ptr(x), _ptr(pos), _ptr(z), _ptr(mass), mass.numel(),
                   _ptr(mol_ptr), ctx.n_mol, _ptr(cen), _ptr(g_x), _ptr(g_pos), _stream())
        return g_x, g_pos, None, None, None, None
