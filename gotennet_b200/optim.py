"""Fused optimiser step of the reference training loop.

The reference trains with `torch.optim.AdamW(params, lr, weight_decay, eps=1e-7)`
(models/goten_model.py:528-534), Lightning norm clipping at 5.0
(configs/trainer/default.yaml:10) and a linear lr warm-up that rewrites
`param_groups[*]["lr"]` before every step (goten_model.py:557-572).

`FusedAdamW` keeps that surface (`param_groups`, `step`, `zero_grad`, `state_dict`) but
stores parameters, gradients and both moments in FLAT fp32 buffers: the data-parallel
all-reduce (parallel.FlatGradBuffer) already produces the flat gradient, and the whole
step is two kernels behind it — a deterministic sum of squares and one clip+AdamW update
(csrc/optim.cu) — with the clip coefficient formed on the device (no host read).
There is no CPU or eager-PyTorch fallback: a CPU parameter raises.
"""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import torch

from ._lib import GotenError, lib
from .parallel import FlatGradBuffer


class FusedAdamW:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-4, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-7, weight_decay: float = 1e-2, max_grad_norm: Optional[float] = None,
                 grad_buffer: Optional[FlatGradBuffer] = None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        for p in self.params:
            if not p.is_cuda:
                raise GotenError("FusedAdamW needs CUDA parameters (there is no CPU path)")
            if p.dtype != torch.float32:
                raise GotenError(f"float32 parameters expected, got {p.dtype}")
        dev = self.params[0].device
        self.grads = grad_buffer if grad_buffer is not None else FlatGradBuffer(self.params)
        if [id(p) for p in self.grads.params] != [id(p) for p in self.params]:
            raise ValueError("grad_buffer must be built over the same parameters in the same order")
        n = self.grads.numel
        # flat parameter storage with the gradient buffer's layout (16-byte aligned slots, zero padding):
        # every Parameter becomes a view of one buffer, values preserved
        self.flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        for o, sz, p in zip(self.grads.offsets, self.grads.sizes, self.params):
            view = self.flat_p[o:o + sz]
            view.copy_(p.data.reshape(-1))
            p.data = view.view_as(p)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self._partial = torch.empty(lib().cdll.goten_sumsq_workspace_floats(), dtype=torch.float32, device=dev)
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.step_count = 0
        self.max_grad_norm = max_grad_norm
        self.param_groups = [{"params": self.params, "lr": lr, "betas": betas, "eps": eps, "weight_decay": weight_decay}]

    # -- torch.optim surface -------------------------------------------------
    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @property
    def grad_norm(self) -> torch.Tensor:
        """Global L2 norm of the (un-clipped, un-scaled) gradient of the last step, as a device scalar."""
        return self._sumsq.sqrt()

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0, packed: bool = False) -> None:
        """One AdamW step.  `packed=True`: the flat gradient buffer is already filled (e.g. by
        FlatGradBuffer.all_reduce); `grad_scale` multiplies the gradient first (1/world_size for a mean)."""
        g = self.grads.flat if packed else self.grads.pack()
        pg = self.param_groups[0]
        b1, b2 = pg["betas"]
        self.step_count += 1
        st = torch.cuda.current_stream(self.flat_p.device).cuda_stream
        L = lib()
        n = self.flat_p.numel()
        clip = self.max_grad_norm is not None and self.max_grad_norm > 0
        L.call("goten_sumsq", g.data_ptr(), n, self._partial.data_ptr(), self._sumsq.data_ptr(), st)
        L.call("goten_adamw_step", self.flat_p.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(),
               self.exp_avg_sq.data_ptr(), n, float(pg["lr"]), float(b1), float(b2), float(pg["eps"]),
               float(pg["weight_decay"]), 1.0 - b1 ** self.step_count, 1.0 - b2 ** self.step_count,
               float(self.max_grad_norm) if clip else 0.0, self._sumsq.data_ptr(), float(grad_scale), st)

    def state_dict(self) -> dict:
        return {"step": self.step_count, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "param_groups": [{k: v for k, v in self.param_groups[0].items() if k != "params"}]}

    def load_state_dict(self, sd: dict) -> None:
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.param_groups[0].update(sd["param_groups"][0])
