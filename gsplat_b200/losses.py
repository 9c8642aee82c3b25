"""Fused photometric loss of the train step (caller-side fusion).

``l1_loss(render, target)`` == ``(render - target).abs().mean()`` -- the L1 term of the reference trainer
(/root/reference/examples/simple_trainer.py: ``F.l1_loss(colors, pixels)``) -- in two launches forward and
one backward instead of seven ATen launches over the 25 MB image.  The sum is a fixed-order two-level
reduction: bit-reproducible run to run."""
from __future__ import annotations

import torch
from torch import Tensor

from ._cabi import check, lib, ptr, require_cuda


class _L1Loss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a: Tensor, b: Tensor):
        dev = require_cuda(a, b)
        if a.shape != b.shape:
            raise ValueError(f"l1_loss: shapes differ, {tuple(a.shape)} vs {tuple(b.shape)}")
        if a.dtype != torch.float32 or b.dtype != torch.float32:
            raise TypeError("l1_loss: float32 tensors expected")
        a, b = a.contiguous(), b.contiguous()
        # the kernels read float4: a contiguous VIEW with a storage offset may be only 4-byte aligned
        if a.data_ptr() % 16:
            a = a.clone()
        if b.data_ptr() % 16:
            b = b.clone()
        if a.numel() == 0:
            raise ValueError("l1_loss of an empty tensor")
        L = lib()
        loss = torch.empty((), device=dev, dtype=torch.float32)
        ws = torch.empty(L.gsb200_l1_loss_workspace_bytes(), device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream().cuda_stream
            check(L.gsb200_l1_loss_fwd(a.numel(), ptr(a), ptr(b), ptr(loss), ptr(ws), st), "l1_loss")
        ctx.save_for_backward(a, b)
        return loss

    @staticmethod
    def backward(ctx, v_loss: Tensor):
        a, b = ctx.saved_tensors
        v_loss = v_loss.to(dtype=torch.float32).contiguous()
        v_a = torch.empty_like(a)
        with torch.cuda.device(a.device):
            st = torch.cuda.current_stream().cuda_stream
            check(lib().gsb200_l1_loss_bwd(a.numel(), ptr(a), ptr(b), ptr(v_loss), ptr(v_a), st), "l1_loss_bwd")
        return (v_a if ctx.needs_input_grad[0] else None), (-v_a if ctx.needs_input_grad[1] else None)


def l1_loss(input: Tensor, target: Tensor) -> Tensor:
    """Mean absolute error over all elements (``torch.nn.functional.l1_loss`` with the default reduction)."""
    return _L1Loss.apply(input, target)
