"""Fused photometric loss of the train step (caller-side fusion).

``l1_loss(render, target)`` == ``(render - target).abs().mean()`` -- the L1 term of the reference trainer
(/root/reference/examples/simple_trainer.py: ``F.l1_loss(colors, pixels)``) -- in two launches forward and
one backward instead of seven ATen launches over the 25 MB image.  The sum is a fixed-order two-level
reduction: bit-reproducible run to run."""
from __future__ import annotations

import ctypes

import torch

from ._cabi import stream as _raw_stream
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
            st = _raw_stream()
            check(L.gsb200_l1_loss_fwd(a.numel(), ptr(a), ptr(b), ptr(loss), ptr(ws), st), "l1_loss")
        ctx.save_for_backward(a, b)
        return loss

    @staticmethod
    def backward(ctx, v_loss: Tensor):
        a, b = ctx.saved_tensors
        v_loss = v_loss.to(dtype=torch.float32).contiguous()
        v_a = torch.empty_like(a)
        with torch.cuda.device(a.device):
            st = _raw_stream()
            check(lib().gsb200_l1_loss_bwd(a.numel(), ptr(a), ptr(b), ptr(v_loss), ptr(v_a), st), "l1_loss_bwd")
        return (v_a if ctx.needs_input_grad[0] else None), (-v_a if ctx.needs_input_grad[1] else None)


def l1_loss(input: Tensor, target: Tensor) -> Tensor:
    """Mean absolute error over all elements (``torch.nn.functional.l1_loss`` with the default reduction)."""
    return _L1Loss.apply(input, target)


class _SsimLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img1: Tensor, img2: Tensor):
        dev = require_cuda(img1, img2)
        B, C, H, W = img1.shape
        L = lib()
        need_grad = ctx.needs_input_grad[0]
        loss = torch.empty((), device=dev, dtype=torch.float32)
        maps = torch.empty((3, B, C, H, W), device=dev, dtype=torch.float32) if need_grad else None
        ws = torch.empty(max(L.gsb200_ssim_workspace_bytes(B, C, H, W), 4), device=dev, dtype=torch.uint8)
        xs = (ctypes.c_int64 * 4)(*img1.stride())
        ys = (ctypes.c_int64 * 4)(*img2.stride())
        with torch.cuda.device(dev):
            st = _raw_stream()
            check(L.gsb200_ssim_fwd(B, C, H, W, ptr(img1), xs, ptr(img2), ys, ptr(maps), ptr(ws), ptr(loss), st), "ssim_loss")
        ctx.save_for_backward(img1, img2, maps)
        return loss

    @staticmethod
    def backward(ctx, v_loss: Tensor):
        img1, img2, maps = ctx.saved_tensors
        if maps is None:
            return None, None
        B, C, H, W = img1.shape
        v_loss = v_loss.to(dtype=torch.float32).contiguous()
        v_x = torch.empty_strided(img1.shape, img1.stride(), device=img1.device, dtype=torch.float32)
        xs = (ctypes.c_int64 * 4)(*img1.stride())
        ys = (ctypes.c_int64 * 4)(*img2.stride())
        vs = (ctypes.c_int64 * 4)(*v_x.stride())
        with torch.cuda.device(img1.device):
            st = _raw_stream()
            check(
                lib().gsb200_ssim_bwd(B, C, H, W, ptr(img1), xs, ptr(img2), ys, ptr(maps), ptr(v_loss), ptr(v_x), vs, st),
                "ssim_loss_bwd",
            )
        return v_x, None


def ssim_loss(img1: Tensor, img2: Tensor, window_size: int = 11) -> Tensor:
    """``1 - SSIM`` of two image batches ``(B, C, H, W)`` -- the reference's ``gsplat.losses.ssim_loss``
    (/root/reference/gsplat/losses.py:154-201) as it evaluates without the third-party ``fused_ssim`` package:
    11x11 Gaussian window (sigma 1.5), zero padding, mean over batch, channels and pixels.  One fused pass forward,
    one backward; the images are read through their strides (the trainer's ``render.permute(0, 3, 1, 2)`` view is
    not copied).  Gradient flows to ``img1`` only (the trainer's target image needs none)."""
    if window_size != 11:
        raise NotImplementedError("gsplat_b200.ssim_loss is built for the reference's default window_size=11")
    if img1.dim() != 4 or img1.shape != img2.shape:
        raise ValueError(f"ssim_loss expects two (B, C, H, W) tensors of equal shape, got {tuple(img1.shape)} and {tuple(img2.shape)}")
    if img1.dtype != torch.float32 or img2.dtype != torch.float32:
        raise TypeError("ssim_loss: float32 tensors expected")
    if img2.requires_grad:
        raise NotImplementedError("ssim_loss: gradient with respect to the second image is not implemented")
    if any(s < 0 for s in img1.stride() + img2.stride()):
        img1, img2 = img1.contiguous(), img2.contiguous()
    return _SsimLoss.apply(img1, img2)
