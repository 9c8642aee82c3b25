"""Executes INTEGRATION.md section B in a process of its own (run by tests/test_reference_suite.py):
the reference package's op-level ``rasterize_to_pixels`` -- its own Python wrapper and its own registered autograd --
is evaluated on the reference's kernels, then gsplat_b200.registry.register() re-binds the two schemas and the same
call is repeated.  Forward must agree to rtol 1e-4 / atol 1e-5 on all but decision-flip pixels, gradients to the
reference's run-to-run spread."""
import math
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import refcuda  # noqa: E402
from tests import scene  # noqa: E402


def main():
    gsplat = refcuda.import_package()
    from gsplat.cuda._wrapper import fully_fused_projection, isect_offset_encode, isect_tiles, rasterize_to_pixels

    dev = "cuda:0"
    sc = scene.make_scene(n_max=60000)
    W, H, C = 640, 360, 2
    Ks = scene.rescale_K(sc["Ks"], sc["width"], sc["height"], W, H)[:C]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    means, quats, scales, opac = t(sc["means"]), t(sc["quats"]), t(sc["scales"]), t(sc["opacities"])
    vm, K = t(sc["viewmats"][:C]), t(Ks)
    radii, m2, dep, con, _ = fully_fused_projection(means, None, quats, scales, vm, K, W, H)
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    _, ids, fl = isect_tiles(m2, radii, dep, 16, tw, th)
    off = isect_offset_encode(ids, C, tw, th)
    g = torch.Generator(device=dev).manual_seed(1)
    col = torch.rand((C, means.shape[0], 3), device=dev, generator=g)
    op = opac[None].expand(C, -1).contiguous()
    bg = torch.rand((C, 3), device=dev, generator=g)
    v_rc, v_ra = torch.randn((C, H, W, 3), device=dev, generator=g), torch.randn((C, H, W, 1), device=dev, generator=g)

    def run():
        ins = [x.detach().clone().requires_grad_(True) for x in (m2, con, col, op, bg)]
        rc, ra = rasterize_to_pixels(ins[0], ins[1], ins[2], ins[3], W, H, 16, off, fl, backgrounds=ins[4])
        grads = torch.autograd.grad((rc * v_rc).sum() + (ra * v_ra).sum(), ins)
        return rc.detach(), ra.detach(), grads

    r1, r2 = run(), run()
    import gsplat_b200.registry as reg

    reg.register()
    o = run()
    bad = float((((o[0] - r1[0]).abs() - (1e-4 * r1[0].abs() + 1e-5)).amax(-1) > 0).float().mean())
    assert bad < 5e-4, f"{bad * 100:.4f}% pixels out of tolerance"
    rel = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-30))  # noqa: E731
    for name, a, b, b2 in zip(("v_means2d", "v_conics", "v_colors", "v_opacities", "v_backgrounds"), o[2], r1[2], r2[2]):
        spread, ours = rel(b2, b), rel(a, b)
        assert ours < max(2e-4, 20 * spread), f"{name}: rel L2 {ours:.3e} (reference run-to-run {spread:.3e})"
    print("registry binding ok: torch.ops.gsplat.rasterize_to_pixels_3dgs[_bwd] -> libgsplat_b200.so under the reference's autograd")


if __name__ == "__main__":
    main()
