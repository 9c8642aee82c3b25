"""N-rank check + timing of the NVLS gradient all-reduce kernel against NCCL:
    torchrun --nproc-per-node 2 tests/dist_nvls_check.py"""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from gsplat_b200 import distributed as D  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    N = 1006065
    shapes = {"means": (N, 3), "quats": (N, 4), "scales": (N, 3), "opacities": (N,), "sh": (N, 16, 3)}
    params = {k: torch.zeros(s, device=dev).requires_grad_(True) for k, s in shapes.items()}
    sweep = (("nvls", 32), ("nvls", 64), ("nvls", 128), ("p2p", 32), ("p2p", 64), ("p2p", 128))
    if os.environ.get("GSB200_CHECK_FAST", "0") == "1":  # 8-GPU runs are charged eightfold: one configuration per kernel
        sweep = (("nvls", 128), ("p2p", 128))
    for algo, blocks in sweep:
        arena = D.NvlsGradArena(params, blocks=blocks, algo=algo)
        g = torch.Generator(device=dev).manual_seed(17 + rank)
        for it in range(3):
            want = {}
            for k, p in params.items():
                src = torch.randn(p.shape, device=dev, generator=g)
                if k in ("means", "sh"):  # written in place in the arena (what the fused backward does)
                    arena.views[k].copy_(src)
                    p.grad = arena.views[k]
                else:  # produced elsewhere -> staged by all_reduce()
                    p.grad = src.clone()
                w = src.clone()
                dist.all_reduce(w)
                want[k] = w
            arena.all_reduce()
            torch.cuda.synchronize()
            for k, p in params.items():
                assert p.grad.data_ptr() == arena.views[k].data_ptr()
                if world == 2:
                    assert torch.equal(p.grad, want[k]), f"{k}: {algo} sum differs from NCCL (blocks={blocks}, it={it})"
                else:
                    torch.testing.assert_close(p.grad, want[k], rtol=1e-5, atol=1e-5)
        # timing: ours vs one coalesced NCCL all-reduce of the same payload
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        res = {}
        for name, fn in (("nvls", arena.all_reduce), ("nccl", lambda: D.all_reduce_gaussian_grads(list(params.values()), coalesced=True))):
            for _ in range(3):
                fn()
            dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[name] = float(t)
        if rank == 0:
            mb = arena.numel * 4 / 1e6
            print(f"nvls check ok: world={world} algo={algo} blocks={blocks} payload={mb:.1f} MB  nvls {res['nvls']:.3f} ms  nccl {res['nccl']:.3f} ms", flush=True)
        del arena

    # row-sparse variants (rows-nvls / rows-p2p): every rank marks a different random 30 % of the rows as touched, its
    # other rows are zero; the result must equal the dense NCCL sum, and only the union's rows may have moved
    for algo in ("nvls", "p2p"):
        arena = D.NvlsGradArena(params, blocks=64, algo=algo, row_sparse=True)
        assert arena.row_sparse and arena.bits is not None
        g = torch.Generator(device=dev).manual_seed(99 + rank)
        for it in range(2):
            seen = torch.rand(N, device=dev, generator=g) < 0.3
            words = torch.zeros(((N + 31) // 32) * 32, dtype=torch.int64, device=dev)
            words[:N] = seen.long()
            packed = (words.view(-1, 32) << torch.arange(32, device=dev)).sum(1)
            arena.bits.copy_(torch.where(packed >= 2**31, packed - 2**32, packed).to(torch.int32))
            want = {}
            for k, p in params.items():
                src = torch.randn(p.shape, device=dev, generator=g) * seen.view((N,) + (1,) * (p.dim() - 1))
                arena.views[k].copy_(src)
                p.grad = arena.views[k]
                w = src.clone()
                dist.all_reduce(w)
                want[k] = w
            arena.stats.zero_()
            arena._bits_requests = 1
            arena.all_reduce()
            torch.cuda.synchronize()
            assert arena.last_kind == "rows-" + algo
            for k, p in params.items():
                if world == 2:
                    assert torch.equal(p.grad, want[k]), f"{k}: rows-{algo} differs from NCCL"
                else:
                    torch.testing.assert_close(p.grad, want[k], rtol=1e-5, atol=1e-5)
            moved = torch.tensor([int(arena.stats.item())], device=dev)
            dist.all_reduce(moved)
            dense_vec = sum(p.numel() for p in params.values()) // 4
            frac = float(moved) / dense_vec
            # whole-vector rows (quats 4 + SH 48 of the 59 floats) move only when touched by some rank (union of `world`
            # random 30 % sets); the narrow segments (means, scales, opacities: 7 floats) move with their 32-row group
            bound = (1.0 - 0.7**world) * 52 / 59 + 7 / 59 + 0.03
            assert frac < bound, f"rows-{algo} moved {frac:.3f} of the payload (bound {bound:.3f})"
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            arena._bits_requests = 1
            arena.all_reduce()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"nvls check ok: world={world} rows-{algo} moved {frac * 100:.1f} % of the dense payload, {float(t):.3f} ms", flush=True)
        del arena

    # end to end: the fused backward writes its gradients INTO the arena (no staging copy), and the reduced
    # gradients equal those of the NCCL path
    import numpy as np

    import gsplat_b200
    from gsplat_b200 import ops
    from tests import scene

    sc = scene.make_scene(n_max=60000, sh_degree=3)
    W, H = 640, 360
    Ks = scene.rescale_K(sc["Ks"], sc["width"], sc["height"], W, H)
    P = {k: torch.from_numpy(sc[k]).to(dev).requires_grad_(True) for k in ("means", "quats", "scales", "opacities", "sh")}
    vm, K = torch.from_numpy(sc["viewmats"][rank % 3][None]).to(dev), torch.from_numpy(Ks[rank % 3][None]).to(dev)
    tgt = torch.rand((1, H, W, 3), device=dev, generator=torch.Generator(device=dev).manual_seed(rank))

    def backward():
        for p in P.values():
            p.grad = None
        rc, _, _ = gsplat_b200.rasterization(P["means"], P["quats"], P["scales"], P["opacities"], P["sh"], vm, K, W, H, sh_degree=3, packed=False)
        (rc - tgt).abs().mean().backward()

    backward()
    D.all_reduce_gaussian_grads(list(P.values()), coalesced=True)
    want = {k: p.grad.clone() for k, p in P.items()}
    arena = D.NvlsGradArena(P)
    ops.set_gradient_allocator(arena.allocator)
    for _ in range(2):
        backward()
        for k in ("means", "quats", "scales", "sh"):
            assert P[k].grad.data_ptr() == arena.views[k].data_ptr(), f"{k}: gradient was not produced in the arena"
        arena.all_reduce()
        torch.cuda.synchronize()
        for k, p in P.items():  # the backward's float atomics are not order-deterministic: compare in norm
            rel = float((p.grad - want[k]).norm() / want[k].norm().clamp_min(1e-30))
            assert rel < 1e-5, f"{k}: rel error {rel:.3e} vs the NCCL path"
    ops.set_gradient_allocator(None)
    assert arena.last_kind.startswith("rows-"), arena.last_kind  # the fused backward published its row bitmap
    if rank == 0:
        print(f"nvls check ok: end-to-end gradients in the arena ({arena.last_kind})", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
