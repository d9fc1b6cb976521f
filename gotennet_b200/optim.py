"""Fused optimiser step of the reference training loop.

The reference trains with `torch.optim.AdamW(params, lr, weight_decay, eps=1e-7)`
(models/goten_model.py:528-534), Lightning norm clipping at 5.0
(configs/trainer/default.yaml:10) and a linear lr warm-up that rewrites
`param_groups[*]["lr"]` before every step (goten_model.py:557-572).

`FusedAdamW` is a `torch.optim.Optimizer` (closure-taking `step`, `defaults`, `state`, `param_groups`,
torch-format `state_dict`: it drops into `GotenModel.configure_optimizers` and under the reference's
ReduceLROnPlateau / CosineAnnealingLR schedulers) but stores parameters, gradients and both moments in FLAT fp32 buffers: the data-parallel
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


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.Optimizer subclass (Lightning / lr schedulers accept it: `step(closure)`, `defaults`, `state`,
    `param_groups`, `add_param_group`, torch-format `state_dict` / `load_state_dict` with per-parameter `step`,
    `exp_avg`, `exp_avg_sq`) whose storage is flat: the per-parameter state tensors are views of two flat buffers."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-4, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-7, weight_decay: float = 1e-2, max_grad_norm: Optional[float] = None,
                 grad_buffer: Optional[FlatGradBuffer] = None):
        params = list(params)
        if params and isinstance(params[0], dict):
            if len(params) != 1:
                raise NotImplementedError("FusedAdamW keeps one flat buffer: a single parameter group (the reference "
                                          "passes self.parameters(), goten_model.py:528)")
            params = list(params[0]["params"])
        plist = [p for p in params if p.requires_grad]
        if not plist:
            raise ValueError("optimizer got an empty parameter list")
        for p in plist:
            if not p.is_cuda:
                raise GotenError("FusedAdamW needs CUDA parameters (there is no CPU path)")
            if p.dtype != torch.float32:
                raise GotenError(f"float32 parameters expected, got {p.dtype}")
        super().__init__(plist, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.params = plist
        dev = plist[0].device
        self.grads = grad_buffer if grad_buffer is not None else FlatGradBuffer(plist)
        if [id(p) for p in self.grads.params] != [id(p) for p in plist]:
            raise ValueError("grad_buffer must be built over the same parameters in the same order")
        n = self.grads.numel
        # flat parameter storage with the gradient buffer's layout (16-byte aligned slots, zero padding):
        # every Parameter becomes a view of one buffer, values preserved
        self.flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        for o, sz, p in zip(self.grads.offsets, self.grads.sizes, plist):
            view = self.flat_p[o:o + sz]
            view.copy_(p.data.reshape(-1))
            p.data = view.view_as(p)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self._partial = torch.empty(lib().cdll.goten_sumsq_workspace_floats(), dtype=torch.float32, device=dev)
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.step_count = 0
        self.max_grad_norm = max_grad_norm
        self._alias_state()

    def _alias_state(self) -> None:
        """self.state[p] in torch.optim.AdamW's format, the moment tensors aliasing the flat buffers."""
        for o, sz, p in zip(self.grads.offsets, self.grads.sizes, self.params):
            self.state[p] = {"step": torch.tensor(float(self.step_count)),
                             "exp_avg": self.exp_avg[o:o + sz].view_as(p),
                             "exp_avg_sq": self.exp_avg_sq[o:o + sz].view_as(p)}

    def add_param_group(self, param_group) -> None:
        if getattr(self, "flat_p", None) is not None:
            raise NotImplementedError("FusedAdamW's flat buffers are sized at construction: one parameter group")
        super().add_param_group(param_group)

    @property
    def grad_norm(self) -> torch.Tensor:
        """Global L2 norm of the (un-clipped, un-scaled) gradient of the last step, as a device scalar."""
        return self._sumsq.sqrt()

    def step(self, closure=None, grad_scale: float = 1.0, packed: bool = False):
        """One AdamW step; returns the closure's loss (torch.optim contract: Lightning passes the closure that runs
        training_step + backward).  `packed=True`: the flat gradient buffer is already filled (e.g. by
        FlatGradBuffer.all_reduce); `grad_scale` multiplies the gradient first (1/world_size for a mean)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        with torch.no_grad():
            g = self.grads.flat if packed else self.grads.pack()
            pg = self.param_groups[0]
            b1, b2 = pg["betas"]
            self.step_count += 1
            for st_ in self.state.values():
                st_["step"].fill_(float(self.step_count))
            st = torch.cuda.current_stream(self.flat_p.device).cuda_stream
            L = lib()
            n = self.flat_p.numel()
            clip = self.max_grad_norm is not None and self.max_grad_norm > 0
            with torch.cuda.device(self.flat_p.device):
                L.call("goten_sumsq", g.data_ptr(), n, self._partial.data_ptr(), self._sumsq.data_ptr(), st)
                L.call("goten_adamw_step", self.flat_p.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(),
                       self.exp_avg_sq.data_ptr(), n, float(pg["lr"]), float(b1), float(b2), float(pg["eps"]),
                       float(pg["weight_decay"]), 1.0 - b1 ** self.step_count, 1.0 - b2 ** self.step_count,
                       float(self.max_grad_norm) if clip else 0.0, self._sumsq.data_ptr(), float(grad_scale), st)
        return loss

    def load_state_dict(self, state_dict: dict) -> None:
        """torch-format state (also what torch.optim.AdamW.state_dict() produces for the same parameter list): the
        loaded moments are copied into the flat buffers and the per-parameter entries re-aliased."""
        super().load_state_dict(state_dict)
        steps = set()
        for o, sz, p in zip(self.grads.offsets, self.grads.sizes, self.params):
            st_ = self.state.get(p, {})
            if "exp_avg" in st_:
                self.exp_avg[o:o + sz].copy_(st_["exp_avg"].reshape(-1))
                self.exp_avg_sq[o:o + sz].copy_(st_["exp_avg_sq"].reshape(-1))
                steps.add(int(float(st_["step"])))
        if len(steps) > 1:
            raise ValueError(f"per-parameter step counts differ ({sorted(steps)}): not a single-group AdamW state")
        self.step_count = steps.pop() if steps else 0
        self._alias_state()
