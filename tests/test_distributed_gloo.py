"""World-size-2 gloo tests (CPU) of the view-axis DP host logic in gsplat_b200.distributed: view
sharding and the bucketed gradient all-reduce.  The render kernels need a GPU, so a linear stand-in
"renderer" is used whose gradients are known in closed form."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gsplat_b200 import distributed as D

    torch.manual_seed(0)  # replicated parameters
    params = [torch.randn(7, 3, requires_grad=True), torch.randn(7, 4, requires_grad=True), torch.randn(7, requires_grad=True)]
    views = torch.arange(5, dtype=torch.float32)  # 5 "cameras"
    ids = D.shard_views(5)
    assert ids == list(range(rank, 5, world))
    loss = sum((views[i] + 1.0) * sum((p * p).sum() for p in params) for i in ids)
    loss.backward()
    D.all_reduce_gaussian_grads(params)
    total = sum(float(v) + 1.0 for v in views)
    for p in params:
        assert torch.allclose(p.grad, 2.0 * total * p.detach(), rtol=1e-5, atol=1e-6)
    # a rank whose parameter got no gradient still participates (zeros are reduced)
    q = [torch.ones(3, requires_grad=True)]
    if rank == 0:
        (q[0] * 2).sum().backward()
    D.all_reduce_gaussian_grads(q, average=True)
    assert torch.allclose(q[0].grad, torch.full((3,), 1.0))
    # async variant
    r = [torch.ones(4, requires_grad=True)]
    (r[0] * float(rank + 1)).sum().backward()
    bucket, work = D.all_reduce_gaussian_grads(r, async_op=True)
    work.wait()
    bucket.unpack(bucket._scale)
    assert torch.allclose(r[0].grad, torch.full((4,), 3.0))
    # coalesced in-place variant (gloo runs the calls one by one; NCCL fuses them into one launch)
    c = [torch.ones(5, requires_grad=True), torch.ones(2, 3, requires_grad=True)]
    (c[0] * float(rank + 1)).sum().backward()  # c[1] has no gradient on any rank
    D.all_reduce_gaussian_grads(c, coalesced=True)
    assert torch.allclose(c[0].grad, torch.full((5,), 3.0)) and torch.equal(c[1].grad, torch.zeros(2, 3))
    # ---- reference-semantics seams: variable-count, differentiable row collectives
    n_mine = 3 + rank  # rank 0 owns 3 rows, rank 1 owns 4
    counts = D.all_gather_ints(n_mine, "cpu")
    assert counts == [3, 4]
    a = (torch.arange(n_mine * 2, dtype=torch.float32).reshape(n_mine, 2) + 100 * rank).requires_grad_(True)
    b = torch.arange(n_mine, dtype=torch.float32) + 10 * rank
    ga, gb = D.all_gather_rows([a, b], counts)
    assert ga.shape == (7, 2) and gb.shape == (7,)
    assert torch.equal(gb, torch.tensor([0.0, 1, 2, 10, 11, 12, 13]))
    (ga * (rank + 1)).sum().backward()  # every rank's gradient flows back to the owner: 1 + 2 = 3
    assert torch.allclose(a.grad, torch.full_like(a, 3.0))
    # all-to-all: rank r sends (r + 1 + j) rows to rank j
    send = [rank + 1 + j for j in range(world)]
    recv = [j + 1 + rank for j in range(world)]
    x = (torch.arange(sum(send), dtype=torch.float32)[:, None] + 1000 * rank).requires_grad_(True)
    ids = torch.arange(sum(send), dtype=torch.int32) + 1000 * rank
    rx, rid = D.all_to_all_rows([x, ids], send, recv)
    assert rx.shape == (sum(recv), 1) and rid.dtype == torch.int32 and torch.equal(rx[:, 0].long(), rid.long())
    rx.sum().backward()
    assert torch.allclose(x.grad, torch.ones_like(x))
    # the reference's helper names, on the examples of its docstrings (gsplat/distributed.py:133-141, 210-221)
    cpu = torch.device("cpu")
    assert D.all_gather_int32(world, 5 + rank, cpu) == [5, 6]
    assert D.all_to_all_int32(world, [10 * rank + 1, 10 * rank + 2], cpu) == [1 + rank, 11 + rank]
    tl = [torch.tensor([1, 2, 3]), torch.tensor([4, 5, 6])] if rank == 0 else [torch.tensor([7, 8, 9]), torch.tensor([10, 11, 12])]
    g0, g1 = D.all_gather_tensor_list(world, tl)
    assert g0.tolist() == [1, 2, 3, 7, 8, 9] and g1.tolist() == [4, 5, 6, 10, 11, 12]
    if rank == 0:
        tl, sp = [torch.tensor([1.0, 2.0, 3.0], requires_grad=True), torch.tensor([4, 5, 6])], [2, 1]
    else:
        tl, sp = [torch.tensor([7.0, 8.0], requires_grad=True), torch.tensor([9, 10])], [1, 1]
    e0, e1 = D.all_to_all_tensor_list(world, tl, sp)
    assert (e0.tolist(), e1.tolist()) == (([1.0, 2.0, 7.0], [4, 5, 9]) if rank == 0 else ([3.0, 8.0], [6, 10]))
    (e0 * (rank + 1)).sum().backward()  # row r of the input went to rank dst: its gradient is dst + 1
    assert tl[0].grad.tolist() == ([1.0, 1.0, 2.0] if rank == 0 else [1.0, 2.0])
    lists = D.all_gather_int_lists([rank, 10 + rank, 7], torch.device("cpu"))
    assert lists == [[0, 10, 7], [1, 11, 7]]
    # camera-major -> local layout
    t = torch.arange(2 * 3 + 2 * 4).float()
    loc = D.camera_major_to_local(t, 2, [3, 4])
    assert loc.shape == (2, 7) and loc[1, 3:].tolist() == [10.0, 11.0, 12.0, 13.0]
    dist.barrier()
    dist.destroy_process_group()
    out.put(rank)


def test_dp_allreduce_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(out.get(timeout=5) for _ in range(2)) == [0, 1]


def test_single_process_is_identity():
    from gsplat_b200 import distributed as D

    p = [torch.ones(3, requires_grad=True)]
    (p[0] * 5).sum().backward()
    D.all_reduce_gaussian_grads(p)
    assert torch.allclose(p[0].grad, torch.full((3,), 5.0))
    assert D.shard_views(5, rank=1, world_size=2) == [1, 3]


def test_grad_arena_allocator_logic():
    """NvlsGradArena.allocator (host logic only; the arena memory itself needs GPUs): hands out FRESH views of its
    segments -- which autograd then adopts as .grad without a copy -- except while a segment still holds an
    accumulated gradient."""
    from gsplat_b200 import distributed as D

    p = torch.nn.Parameter(torch.ones(4, 3))
    flat = torch.zeros(64)
    arena = D.NvlsGradArena.__new__(D.NvlsGradArena)
    arena.params = {"means": p}
    arena.views = {"means": flat[:12].view(4, 3)}
    arena._handed_out = set()
    a = arena.allocator("means", p)
    arena.reset()
    assert a is not arena.views["means"] and a.data_ptr() == flat.data_ptr()
    assert arena.allocator("quats", p) is None and arena.allocator("means", torch.ones(5, 3)) is None

    class F(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            return x * 2

        @staticmethod
        def backward(ctx, g):
            out = arena.allocator("means", p)
            out = torch.empty_like(p) if out is None else out
            out.copy_(g * 2)
            return out

    F.apply(p).sum().backward()
    assert p.grad.data_ptr() == flat.data_ptr() and torch.equal(p.grad, torch.full((4, 3), 2.0))
    F.apply(p).sum().backward()  # accumulation: the segment is occupied, the second gradient is added into it
    assert p.grad.data_ptr() == flat.data_ptr() and torch.equal(p.grad, torch.full((4, 3), 4.0))


def test_grad_arena_two_renders_in_one_graph():
    """Two fused backward nodes feeding the same leaf within ONE backward graph (several renders summed into one
    loss): the kernels write (not accumulate) their outputs, so only the first may get the arena segment; the leaf
    must end up with g1 + g2."""
    from gsplat_b200 import distributed as D

    p = torch.nn.Parameter(torch.ones(4, 3))
    flat = torch.zeros(64)
    arena = D.NvlsGradArena.__new__(D.NvlsGradArena)
    arena.params, arena.views, arena._handed_out = {"means": p}, {"means": flat[:12].view(4, 3)}, set()

    class F(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, k):
            ctx.k = k
            return x * k

        @staticmethod
        def backward(ctx, g):
            out = arena.allocator("means", p)
            out = torch.empty_like(p) if out is None else out
            out.copy_(g * ctx.k)  # WRITES, like project_sh_bwd
            return out, None

    (F.apply(p, 2.0).sum() + F.apply(p, 3.0).sum()).backward()
    assert torch.equal(p.grad, torch.full((4, 3), 5.0)), p.grad
    assert arena._handed_out == {"means"}


def test_cli_spawn_target_is_picklable():
    """cli() hands its worker to torch.multiprocessing.spawn, which pickles the target."""
    import pickle

    from gsplat_b200 import distributed as D

    assert pickle.loads(pickle.dumps(D._distributed_worker)) is D._distributed_worker


def test_async_average_scale_is_applied_by_unpack():
    from gsplat_b200 import distributed as D

    p = [torch.ones(3, requires_grad=True)]
    (p[0] * 4).sum().backward()
    bucket = D.GradBucket(p)
    bucket.pack()
    bucket._scale = 0.5  # what all_reduce_gaussian_grads(average=True, async_op=True) stores for world_size 2
    bucket.unpack()
    assert torch.allclose(p[0].grad, torch.full((3,), 2.0))
    bucket.pack()
    bucket.unpack()  # the stored factor is consumed once
    assert torch.allclose(p[0].grad, torch.full((3,), 2.0))
