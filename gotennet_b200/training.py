"""Force-matching training step without second-order kernels (remainder).

The reference trains energy + force models by back-propagating THROUGH the forces
(`Atomwise(derivative="forces", create_graph=True)`, components/outputs.py:365-375, loss assembled in
models/goten_model.py:448-519): the parameter gradient of a force loss is a mixed second derivative of the
energy.  The kernels of this package are first-order (their backward passes are explicit kernel sequences
and are marked once-differentiable; differentiating through them raises), so the mixed term is evaluated as a
directional derivative instead:

    L = loss(E, F),  F = -dE/dpos,   u = dL/dF (per atom),   gE = dL/dE (per molecule)
    dL/dtheta = sum_mol gE dE/dtheta  -  d/dtheta [ (dE/dpos) . u ]
              = sum_mol gE dE/dtheta  -  d/deps  grad_theta E_total(pos + eps u) |_{eps=0}

and the last term is a central difference of ordinary first-order parameter gradients at displaced positions
(4th-order stencil by default: four extra forward+backward passes, positions moved by at most `h` and `2h`).
Every pass runs the same CUDA kernels as the energy-only step.  This is an APPROXIMATION of the reference's
exact double backward: in fp32 the force-loss part of the gradient carries a relative error of about 1e-3 or better
(tests/test_gpu_parity.py::test_force_matching_step_vs_oracle pins it against the oracle's exact second-order
gradient); the energy-loss part and E, F themselves are exact to the usual 1e-5.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional, Tuple

import torch


class _Batch:
    pass


def _with_pos(batch, pos):
    d = _Batch()
    for k in ("z", "batch", "num_graphs", "ptr"):
        if hasattr(batch, k):
            setattr(d, k, getattr(batch, k))
    d.pos = pos
    return d


def _energy(model, head, batch, pos, plan=None, masks=None):
    d = _with_pos(batch, pos)
    if plan is not None or masks is not None:
        d.representation, d.vector_representation = model(d, plan=plan, attn_drop_masks=masks)
    else:
        d.representation, d.vector_representation = model(d)
    return head(d)[head.property]


def _frozen_network(model, batch, attn_drop_masks=None):
    """(plan, masks) that pin the stochastic / discrete parts of one training step: the radius graph built at the base
    positions (the reference's exact double backward differentiates at FIXED topology: edge set, softmax support and
    out-degrees do not move with the positions) and ONE draw of the per-layer attention-dropout factors shared by every
    pass of the step.  Models without the GotenNetWrapper extensions get (None, None) and are called plainly."""
    if not (hasattr(model, "distance") and hasattr(model, "draw_attn_drop_masks")):
        if attn_drop_masks is not None:
            raise ValueError("attn_drop_masks need a GotenNetWrapper")
        return None, None
    plan = model.distance.plan(batch.pos.detach(), batch.batch)
    masks = attn_drop_masks if attn_drop_masks is not None else model.draw_attn_drop_masks(plan)
    return plan, masks


def energy_and_forces(model, head, batch, attn_drop_masks=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """E [n_mol, n_out] and F = -dE_total/dpos [N, 3] (first order, detached)."""
    pos = batch.pos.detach().clone().requires_grad_(True)
    E = _energy(model, head, batch, pos, None, attn_drop_masks)
    (g,) = torch.autograd.grad(E.sum(), pos)
    return E.detach(), -g


def force_matching_backward(model, head, batch, loss_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
                            params: Optional[Iterable[torch.nn.Parameter]] = None, h: float = 2.5e-3, order: int = 4,
                            attn_drop_masks=None):
    """Accumulates d loss_fn(E, F) / d theta into `.grad` of the parameters (model + head unless `params` is given)
    and returns (loss, E, F) detached.  `head` must not compute derivatives itself (derivative=None).
    `h` is the largest atomic displacement (in the units of pos) of the inner stencil points; the default balances
    truncation (the energy surface has large higher derivatives: h = 1e-2 is already 10 % off) against fp32 round-off
    (measured with the fp32 oracle: worst parameter tensor 3e-4, median 4e-5 at h = 2.5e-3).
    All passes of one call run the SAME network: the radius graph is built once at the base positions and reused by the
    displaced passes, and in training mode with attn_dropout > 0 the per-layer dropout factors are drawn once
    (`attn_drop_masks`: supply them to reproduce a step; the ones used are left in `model.last_attn_drop_masks`)."""
    if order not in (2, 4):
        raise ValueError("order must be 2 or 4")
    if getattr(head, "derivative", None):
        raise ValueError("use a head without derivative=...: the forces are formed here")
    params = [p for p in (params if params is not None else list(model.parameters()) + list(head.parameters()))
              if p.requires_grad]
    plan, masks = _frozen_network(model, batch, attn_drop_masks)
    if plan is not None:
        model.last_attn_drop_masks = masks
    pos0 = batch.pos.detach().clone().requires_grad_(True)
    E = _energy(model, head, batch, pos0, plan, masks)
    (gpos,) = torch.autograd.grad(E.sum(), pos0, retain_graph=True)
    F = -gpos
    E_leaf, F_leaf = E.detach().requires_grad_(True), F.detach().requires_grad_(True)
    loss = loss_fn(E_leaf, F_leaf)
    gE, u = torch.autograd.grad(loss, [E_leaf, F_leaf], allow_unused=True)

    def grads_of(scalar_fn):
        """first-order parameter gradients of one forward+backward pass, as a list (None -> zeros)"""
        saved = [p.grad for p in params]
        for p in params:
            p.grad = None
        scalar_fn()
        out = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
        for p, s in zip(params, saved):
            p.grad = s
        return out

    total = [torch.zeros_like(p) for p in params]
    if gE is not None:   # energy part: exact, on the graph of the first pass
        torch._foreach_add_(total, grads_of(lambda: E.backward(gradient=gE)))
    scale = float(u.abs().max()) if u is not None else 0.0
    if scale > 0.0:      # force part: -d/deps grad_theta E_total(pos + eps u) at eps = 0
        uhat = (u / scale).detach()
        base = batch.pos.detach()
        # central-difference weights of d/deps for the stencil pairs (+s h, -s h): each pair is differenced first (the
        # subtraction of two nearly equal gradients is exact) and folded into `total` with one multi-tensor axpy, so at
        # most two gradient sets are alive at a time
        weights = {1.0: 1.0 / (2.0 * h)} if order == 2 else {1.0: 8.0 / (12.0 * h), 2.0: -1.0 / (12.0 * h)}

        def G(s_):
            return grads_of(lambda: _energy(model, head, batch, base + s_ * h * uhat, plan, masks).sum().backward())

        for s_, w_ in weights.items():
            gp, gm = G(s_), G(-s_)
            torch._foreach_sub_(gp, gm)
            torch._foreach_add_(total, gp, alpha=-scale * w_)
            del gp, gm
    for p, t in zip(params, total):
        p.grad = t if p.grad is None else p.grad + t
    return loss.detach(), E.detach(), F.detach()
