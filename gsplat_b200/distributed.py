"""View-axis data parallelism for the rasterization() hot path.

Every rank holds the full (replicated) set of Gaussian parameters and renders its own camera views; the
only coupling is the sum of the parameter gradients, done here with ONE bucketed all-reduce of the
59 floats / Gaussian (means 3 | quats 4 | scales 3 | opacities 1 | SH 48 -> 236 B) over NCCL
(NVLink 5 / NVSwitch; gloo in the CPU tests).  This is not the reference's ``distributed=True`` mode
(Gaussian-sharded, /root/reference/gsplat/distributed.py:117-272), which is a "next" row.

Launch: one process per GPU (torchrun / ``cli`` below, the reference's spawner restated from
/root/reference/gsplat/distributed.py:319-375).
"""
from __future__ import annotations

import os
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor


def shard_views(n_views: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> List[int]:
    """Indices of the camera views rank ``rank`` renders: ``rank::world_size`` (round robin)."""
    rank = dist.get_rank() if rank is None else rank
    world_size = dist.get_world_size() if world_size is None else world_size
    return list(range(rank, n_views, world_size))


class GradBucket:
    """Flat fp32 buffer the gradients of all parameter tensors are packed into for a single collective."""

    def __init__(self, params: Sequence[Tensor]):
        self.params = list(params)
        self.sizes = [p.numel() for p in self.params]
        dev = self.params[0].device
        self.flat = torch.zeros(sum(self.sizes), device=dev, dtype=torch.float32)

    def pack(self) -> Tensor:
        o = 0
        for p, n in zip(self.params, self.sizes):
            if p.grad is None:
                self.flat[o : o + n].zero_()
            else:
                self.flat[o : o + n].copy_(p.grad.reshape(-1))
            o += n
        return self.flat

    def unpack(self, scale: float = 1.0) -> None:
        o = 0
        for p, n in zip(self.params, self.sizes):
            g = self.flat[o : o + n].view_as(p)
            if scale != 1.0:
                g = g * scale
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            o += n


def all_reduce_gaussian_grads(
    params: Sequence[Tensor], bucket: Optional[GradBucket] = None, average: bool = False, group=None,
    async_op: bool = False, coalesced: bool = False,
):
    """Sum (or average) the ``.grad`` of the replicated Gaussian parameters over all ranks with one
    all-reduce.  Returns the bucket (and the work handle when ``async_op``); call ``bucket.unpack()``
    after ``work.wait()`` in the async case.

    ``coalesced=True``: no staging buffer -- the gradient tensors themselves are reduced in place inside
    one coalesced launch (``torch.distributed._coalescing_manager``: a single NCCL group call on NVLink);
    parameters without a gradient get a zero gradient first so that every rank issues the same calls."""
    if coalesced:
        grads = []
        for p in params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            elif not p.grad.is_contiguous():
                p.grad = p.grad.contiguous()
            grads.append(p.grad)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if str(dist.get_backend(group)).lower() == "nccl":
                with dist._coalescing_manager(group=group, device=grads[0].device, async_ops=False):
                    for g in grads:
                        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
            else:  # gloo (CPU tests) has no coalescing: same calls, one by one
                for g in grads:
                    dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
            if average:
                inv = 1.0 / dist.get_world_size(group)
                for g in grads:
                    g.mul_(inv)
        return None
    if bucket is None:
        bucket = GradBucket(params)
    flat = bucket.pack()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        bucket.unpack()
        return (bucket, None) if async_op else bucket
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    scale = 1.0 / dist.get_world_size(group) if average else 1.0
    if async_op:
        bucket._scale = scale
        return bucket, work
    bucket.unpack(scale)
    return bucket


def render_views_dp(
    rasterize: Callable[..., tuple], gaussians: Dict[str, Tensor], viewmats: Tensor, Ks: Tensor, width: int, height: int,
    **kwargs,
):
    """Renders this rank's share (``rank::world``) of the C views in ``viewmats`` [C,4,4] / ``Ks`` [C,3,3]
    with the replicated ``gaussians`` (keys: means, quats, scales, opacities, colors).  Returns
    (render_colors, render_alphas, meta, view_ids)."""
    ids = shard_views(viewmats.shape[0]) if (dist.is_available() and dist.is_initialized()) else list(range(viewmats.shape[0]))
    idx = torch.as_tensor(ids, device=viewmats.device, dtype=torch.long)
    rc, ra, meta = rasterize(
        gaussians["means"], gaussians["quats"], gaussians["scales"], gaussians["opacities"], gaussians["colors"],
        viewmats.index_select(0, idx), Ks.index_select(0, idx), width, height, **kwargs,
    )
    return rc, ra, meta, ids


def cli(fn: Callable, args, verbose: bool = False) -> None:
    """Spawn one process per visible GPU and run ``fn(local_rank, world_rank, world_size, args)``
    (single node).  Under torchrun (RANK set) the process group is created from the environment."""
    if "RANK" in os.environ:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl")
        try:
            fn(local, dist.get_rank(), dist.get_world_size(), args)
        finally:
            dist.barrier()
            dist.destroy_process_group()
        return
    world = torch.cuda.device_count()
    if world <= 1:
        fn(0, 0, 1, args)
        return

    def _worker(local_rank: int):
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", rank=local_rank, world_size=world)
        try:
            fn(local_rank, local_rank, world, args)
        finally:
            dist.barrier()
            dist.destroy_process_group()

    torch.multiprocessing.spawn(_worker, nprocs=world, join=True)
