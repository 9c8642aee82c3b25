"""2-rank NCCL check of rasterization(distributed=True) (gaussian-sharded, reference semantics):
run with   torchrun --nproc-per-node 2 tests/dist_sharded_check.py
Every rank owns a contiguous shard of the Gaussians and ONE camera; its render and the gradients of its
shard must equal those of a single-process render of ALL Gaussians from BOTH cameras.
(Reference test of the same property: /root/reference/tests/_test_distributed.py, test_rasterization.py:816-870.)"""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gsplat_b200  # noqa: E402
from tests import scene  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    for sh_degree, packed in ((3, False), (None, False), (2, True), (None, True)):
        sc = scene.make_scene(n_max=40001, sh_degree=3)
        W, H = 480, 270
        Ks = scene.rescale_K(sc["Ks"], sc["width"], sc["height"], W, H)
        N = sc["means"].shape[0]
        bounds = [0, N // 3, N] if world == 2 else list(np.linspace(0, N, world + 1).astype(int))  # uneven shards
        lo, hi = bounds[rank], bounds[rank + 1]
        K = None if sh_degree is None else (sh_degree + 1) ** 2
        col_all = sc["colors"] if sh_degree is None else np.ascontiguousarray(sc["sh"][:, :K])
        full = {k: torch.from_numpy(sc[k]).to(dev).requires_grad_(True) for k in ("means", "quats", "scales", "opacities")}
        full["colors"] = torch.from_numpy(col_all).to(dev).requires_grad_(True)
        cams = list(range(world))
        vm_all = torch.from_numpy(sc["viewmats"][[c % 3 for c in cams]]).to(dev)
        K_all = torch.from_numpy(Ks[[c % 3 for c in cams]]).to(dev)
        g = torch.Generator(device=dev).manual_seed(5)
        v_all = torch.randn((world, H, W, 3), device=dev, generator=g)
        rc_f, ra_f, _ = gsplat_b200.rasterization(
            full["means"], full["quats"], full["scales"], full["opacities"], full["colors"], vm_all, K_all, W, H,
            sh_degree=sh_degree, packed=False,
        )
        (rc_f * v_all).sum().backward()
        loc = {k: full[k].detach()[lo:hi].clone().requires_grad_(True) for k in full}
        rc, ra, meta = gsplat_b200.rasterization(
            loc["means"], loc["quats"], loc["scales"], loc["opacities"], loc["colors"], vm_all[rank : rank + 1],
            K_all[rank : rank + 1], W, H, sh_degree=sh_degree, packed=packed, distributed=True,
        )
        meta["means2d"].retain_grad()
        (rc * v_all[rank : rank + 1]).sum().backward()
        if packed and sh_degree is not None:  # packed SH forms the view direction per row (torch), not in the fused kernel
            torch.testing.assert_close(rc, rc_f[rank : rank + 1], rtol=1e-5, atol=2e-6)
        else:
            assert torch.equal(rc, rc_f[rank : rank + 1]), f"rank {rank}: render differs (max {(rc - rc_f[rank:rank+1]).abs().max():.3e})"
        assert torch.equal(ra, ra_f[rank : rank + 1])
        for k in loc:
            a, b = loc[k].grad, full[k].grad[lo:hi]
            rel = float((a - b).norm() / b.norm().clamp_min(1e-20))
            assert rel < 2e-5, f"rank {rank} {k}: rel grad error {rel:.3e}"
        assert meta["means2d"].grad is not None and meta["n_cameras"] == 1
        if not packed:
            assert tuple(meta["means2d"].shape) == (1, N, 2)
        else:  # the visible rows of ALL gaussians for this rank's camera, ascending global gaussian index
            gids = meta["gaussian_ids"]
            assert meta["means2d"].dim() == 2 and int(gids.max()) >= hi - 1 - (hi - lo) and bool((gids[1:] > gids[:-1]).all())
            assert int(gids.max()) < N and int(meta["camera_ids"].max()) == 0
        dist.barrier()
        if rank == 0:
            print(f"sharded check ok: sh_degree={sh_degree} packed={packed} shards={[bounds[i+1]-bounds[i] for i in range(world)]}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
