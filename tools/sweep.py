#!/usr/bin/env python
"""Ours vs the real reference CUDA kernels (oracle/_ref/gsplat_ref.so) over a few workloads:
fwd and fwd+bwd milliseconds of the rasterization path (CUDA events, 10 reps after 3 warm-ups)."""
import math
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import gsplat_b200  # noqa: E402
from tests import scene  # noqa: E402

dev = "cuda:0"
so = os.path.join(ROOT, "oracle", "_ref", "gsplat_ref.so")
R = None
if os.path.exists(so):
    torch.ops.load_library(so)
    R = torch.ops.gsplat


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(name, grid, n_max, C, W, H, scale_mult=1.0, sh_degree=3):
    sc = scene.make_scene(scene_grid=grid, n_max=n_max, sh_degree=sh_degree)
    sc["scales"] = sc["scales"] * scale_mult
    Ks = scene.rescale_K(sc["Ks"], sc["width"], sc["height"], W, H)
    cams = [i % 3 for i in range(C)]
    vm = torch.from_numpy(sc["viewmats"][cams]).to(dev)
    K = torch.from_numpy(Ks[cams]).to(dev)
    P = {k: torch.from_numpy(sc[k]).to(dev).requires_grad_(True) for k in ("means", "quats", "scales", "opacities", "sh")}
    v_rc = torch.randn((C, H, W, 3), device=dev) / (C * H * W)

    def ours(bwd):
        rc, ra, meta = gsplat_b200.rasterization(P["means"], P["quats"], P["scales"], P["opacities"], P["sh"], vm, K, W, H, sh_degree=sh_degree, packed=False)
        if bwd:
            for p in P.values():
                p.grad = None
            (rc * v_rc).sum().backward()
        return meta

    meta = ours(False)
    S = meta["flatten_ids"].numel()
    o_f, o_fb = timeit(lambda: ours(False)), timeit(lambda: ours(True))
    r_f = r_fb = float("nan")
    if R is not None:
        means, quats, scales, opac, sh = (P[k].detach() for k in ("means", "quats", "scales", "opacities", "sh"))
        tw, th = math.ceil(W / 16), math.ceil(H / 16)
        op_cn = opac[None].expand(C, -1).contiguous()

        def ref(bwd):
            radii, m2, dep, con, _ = R.projection_ewa_3dgs_fused(means, None, quats, scales, opac, vm, K, W, H, 0.3, 0.01, 1e10, 0.0, False, 0)
            valid = (radii > 0).all(-1)
            raw = R.spherical_harmonics(sh_degree, means, vm, sh, valid, None, None, None, None)
            col = torch.clamp_min(raw + 0.5, 0.0)
            tpg, ids, fl = R.intersect_tile(m2, radii, dep, con, op_cn, None, None, C, 16, tw, th, True, False)
            off = R.intersect_offset(ids, C, tw, th)
            rc, ra, _, last = R.rasterize_to_pixels_3dgs(m2, con, col, op_cn, None, None, W, H, 16, off, fl, False, False)
            if bwd:
                rb = R.rasterize_to_pixels_3dgs_bwd(m2, con, col, op_cn, None, None, off, fl, ra, last, W, H, 16, False, v_rc, torch.zeros_like(ra), False)
                v_col = rb[3] * (col > 0)
                R.spherical_harmonics_bwd(sh_degree, means, vm, sh, valid, None, None, None, None, v_col, True, False, False)
                R.projection_ewa_3dgs_fused_bwd(means, None, quats, scales, vm, K, W, H, 0.3, 0, radii, con, None, rb[1], torch.zeros_like(dep), rb[2], None, False)

        r_f, r_fb = timeit(lambda: ref(False)), timeit(lambda: ref(True))
    print(f"{name:34s} N={P['means'].shape[0]:8d} C={C} {W}x{H} S={S:9d} | ours fwd {o_f:7.3f} fwd+bwd {o_fb:7.3f} | ref fwd {r_f:7.3f} fwd+bwd {r_fb:7.3f} | speedup {r_f / o_f:4.2f}x / {r_fb / o_fb:4.2f}x", flush=True)


if __name__ == "__main__":
    run("cfg2 garden x1, 1 view 1080p", 1, None, 1, 1920, 1080)
    run("cfg3 garden x3x3, 1 view 1080p", 3, None, 1, 1920, 1080)
    run("cfg3 4 views in one call", 3, None, 4, 1920, 1080)
    run("garden x5x5 (2.8M), 1 view 1080p", 5, None, 1, 1920, 1080)
    run("cfg3 big gaussians (scales x4)", 3, None, 1, 1920, 1080, scale_mult=4.0)
    run("cfg3 SH0, 720p", 3, None, 1, 1280, 720, sh_degree=0)
