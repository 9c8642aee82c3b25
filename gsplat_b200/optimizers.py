"""SelectiveAdam: Adam whose step only touches the Gaussians that were visible this iteration, one fused
kernel per parameter tensor (mirror of /root/reference/gsplat/optimizers/selective_adam.py; the "sparse Adam"
of the Taming-3DGS paper).  The kernel is ``gsplat_b200.ops.adam`` (csrc/pergauss.cu adam_kernel)."""
from __future__ import annotations

import torch

from .ops import adam


class SelectiveAdam(torch.optim.Adam):
    """``step(visibility)``: visibility is a bool / 0-1 tensor with one entry per Gaussian (= row of every
    parameter).  Every param group must hold exactly one tensor, as in the reference."""

    def __init__(self, params, eps, betas):
        super().__init__(params=params, eps=eps, betas=betas)

    @torch.no_grad()
    def step(self, visibility):
        for group in self.param_groups:
            if len(group["params"]) != 1:
                raise AssertionError("more than one tensor in group")
            param = group["params"][0]
            if param.grad is None:
                continue
            state = self.state[param]
            if len(state) == 0:
                state["step"] = torch.tensor(0.0, dtype=torch.float32)
                state["exp_avg"] = torch.zeros_like(param, memory_format=torch.preserve_format)
                state["exp_avg_sq"] = torch.zeros_like(param, memory_format=torch.preserve_format)
            beta1, beta2 = group["betas"]
            adam(param, param.grad, state["exp_avg"], state["exp_avg_sq"], visibility, group["lr"], beta1, beta2, group["eps"])
