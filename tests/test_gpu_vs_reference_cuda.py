"""GPU parity against the REAL reference CUDA operators (nerfstudio-project/gsplat v1.6.0), compiled for
sm_100a from /root/reference by oracle/build_ref.py into oracle/_ref/gsplat_ref.so and called through
torch.ops.gsplat.* (schemas: /root/reference/gsplat/cuda/ext.cpp:984-1089).  The reference Python package
does not exist on the GPU box; only the prebuilt .so travels.  Skipped when it is absent.

The reference is built with -use_fast_math (its default) and accumulates gradients with float atomics,
so it is itself only reproducible to ~1e-6 relative; its own tests compare at rtol 1e-5..2.5e-4 /
atol 1e-3..2e-3 for gradients (tests/test_basic.py:2664-2692).  Here: north_star's rtol 1e-4 / atol 1e-5
on rendered colours / alphas (with a bounded count of threshold-flip pixels), relative-L2 bounds on
gradients, exact equality on sort keys when fed identical projections.
"""
import math
import os

import numpy as np
import pytest
import torch

from tests import scene

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
@pytest.fixture(scope="module")
def ref():
    from oracle import refcuda  # one policy for which physical copy of the reference build a process loads

    if not refcuda.available():
        pytest.skip("reference CUDA build missing (python baseline/install_ref.py in the build container)")
    return refcuda.load_ops()


@pytest.fixture(scope="module")
def gs():
    import gsplat_b200

    return gsplat_b200


def _t(a, rg=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t.requires_grad_(True) if rg else t


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _scene(n_max, W, H, C, sh_degree=3, grid=1):
    sc = scene.make_scene(scene_grid=grid, n_max=n_max, sh_degree=sh_degree)
    Ks = scene.rescale_K(sc["Ks"], sc["width"], sc["height"], W, H)[:C]
    return sc, _t(sc["viewmats"][:C]), _t(Ks)


def test_projection_and_sh_vs_reference(ref, gs):
    W, H, C = 960, 540, 2
    sc, vm, Ks = _scene(80000, W, H, C)
    means, quats, scales, opac, sh = (_t(sc[k]) for k in ("means", "quats", "scales", "opacities", "sh"))
    r_radii, r_m2, r_dep, r_con, _ = ref.projection_ewa_3dgs_fused(
        means, None, quats, scales, opac, vm, Ks, W, H, 0.3, 0.01, 1e10, 0.0, False, 0
    )
    radii, m2, dep, con, col, _ = gs.fused_project_sh(means, quats, scales, opac, sh, vm, Ks, W, H, 3)
    rv, ov = (r_radii > 0).all(-1), (radii > 0).all(-1)
    assert (rv != ov).float().mean() < 1e-4, "visibility sets differ"
    both = rv & ov
    assert both.sum() > 10000
    assert ((r_radii[both] - radii[both]).abs() <= 1).all() and (r_radii[both] != radii[both]).float().mean() < 1e-3
    torch.testing.assert_close(m2[both], r_m2[both], rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(dep[both], r_dep[both], rtol=1e-5, atol=1e-6)
    assert _rel(con[both], r_con[both]) < 1e-4
    torch.testing.assert_close(con[both], r_con[both], rtol=5e-3, atol=1e-5)  # fast-math conic of tiny gaussians
    r_col = ref.spherical_harmonics(3, means, vm, sh, rv, None, None, None, None)
    r_col = torch.clamp_min(r_col + 0.5, 0.0)
    torch.testing.assert_close(col[both], r_col[both], rtol=1e-4, atol=1e-5)
    # backward of both ops with common cotangents
    g = torch.Generator(device=DEV).manual_seed(0)
    v_m2, v_dep, v_con = (torch.randn(s, device=DEV, generator=g) for s in ((C, len(means), 2), (C, len(means)), (C, len(means), 3)))
    v_col = torch.randn((C, len(means), 3), device=DEV, generator=g)
    rb = ref.projection_ewa_3dgs_fused_bwd(
        means, None, quats, scales, vm, Ks, W, H, 0.3, 0, r_radii, r_con, None, v_m2, v_dep, v_con, None, False
    )
    r_vcoef, r_vmeans_sh, _, _ = ref.spherical_harmonics_bwd(
        3, means, vm, sh, rv, None, None, None, None, (v_col * (r_col > 0)).contiguous(), True, False, False
    )
    L = gs._cabi.lib()
    from gsplat_b200._cabi import ptr, stream

    v_means, v_quats, v_scales, v_sh = (torch.empty_like(x) for x in (means, quats, scales, sh))
    rc = L.gsb200_project_sh_bwd(
        C, len(means), 16, 3, ptr(means), ptr(quats), ptr(scales), ptr(sh), ptr(vm), ptr(Ks), W, H, 0.3, ptr(r_radii),
        ptr(r_con), None, ptr(r_col.contiguous()), ptr(v_m2), 2, ptr(v_dep), 1, ptr(v_con), 3, ptr(v_col), 3, None,
        ptr(v_means), ptr(v_quats), ptr(v_scales), ptr(v_sh), None, stream(),
    )
    assert rc == 0
    assert _rel(v_means, rb[0] + r_vmeans_sh) < 1e-4
    assert _rel(v_quats, rb[2]) < 1e-3 and _rel(v_scales, rb[3]) < 1e-3
    assert _rel(v_sh, r_vcoef) < 1e-5


def _ref_isect(ref, m2, radii, dep, con, op, C, tw, th):
    tpg, ids, fl = ref.intersect_tile(m2, radii, dep, con, op, None, None, C, 16, tw, th, True, False)
    off = ref.intersect_offset(ids, C, tw, th)
    return tpg, ids, fl, off


@pytest.mark.parametrize("model,cam_id", [("ortho", 1), ("fisheye", 2)])
def test_projection_camera_models_vs_reference(ref, gs, model, cam_id):
    """Ortho / fisheye projection fwd + bwd against the reference CUDA kernels (CameraModelType 1 / 2)."""
    W, H, C = 960, 540, 2
    sc, vm, Ks = _scene(60000, W, H, C)
    if model == "ortho":
        Ks = Ks.clone()
        Ks[:, 0, 0], Ks[:, 1, 1] = 180.0, 170.0
    means, quats, scales, opac = (_t(sc[k]) for k in ("means", "quats", "scales", "opacities"))
    r = ref.projection_ewa_3dgs_fused(means, None, quats, scales, opac, vm, Ks, W, H, 0.3, 0.01, 1e10, 0.0, False, cam_id)
    o = gs.fully_fused_projection(means, None, quats, scales, vm, Ks, W, H, opacities=opac, camera_model=model)
    rv, ov = (r[0] > 0).all(-1), (o[0] > 0).all(-1)
    assert (rv != ov).float().mean() < 1e-4
    both = rv & ov
    assert both.sum() > 5000
    assert ((r[0][both] - o[0][both]).abs() <= 1).all()
    torch.testing.assert_close(o[1][both], r[1][both], rtol=1e-4, atol=2e-3)
    torch.testing.assert_close(o[2][both], r[2][both], rtol=1e-5, atol=1e-6)
    assert _rel(o[3][both], r[3][both]) < 1e-4
    g = torch.Generator(device=DEV).manual_seed(1)
    N = len(means)
    v_m2, v_dep, v_con = (torch.randn(s, device=DEV, generator=g) for s in ((C, N, 2), (C, N), (C, N, 3)))
    rb = ref.projection_ewa_3dgs_fused_bwd(
        means, None, quats, scales, vm, Ks, W, H, 0.3, cam_id, r[0], r[3], None, v_m2, v_dep, v_con, None, True
    )
    L = gs._cabi.lib()
    from gsplat_b200._cabi import ptr, stream

    v_means, v_quats, v_scales, v_vm = (torch.empty_like(x) for x in (means, quats, scales, vm))
    rc = L.gsb200_projection_bwd(
        1, C, N, ptr(means), None, ptr(quats), ptr(scales), ptr(vm), ptr(Ks), W, H, 0.3, cam_id, ptr(r[0]), ptr(r[3]), None,
        ptr(v_m2), 2, ptr(v_dep), 1, ptr(v_con), 3, None, ptr(v_means), None, ptr(v_quats), ptr(v_scales), ptr(v_vm), stream(),
    )
    assert rc == 0
    assert _rel(v_means, rb[0]) < 2e-4 and _rel(v_quats, rb[2]) < 2e-4 and _rel(v_scales, rb[3]) < 2e-4
    assert _rel(v_vm, rb[4]) < 2e-3


def test_isect_vs_reference_on_identical_projection(ref, gs):
    W, H, C = 1280, 720, 2
    sc, vm, Ks = _scene(100000, W, H, C)
    means, quats, scales, opac = (_t(sc[k]) for k in ("means", "quats", "scales", "opacities"))
    radii, m2, dep, con, _ = gs.fully_fused_projection(means, None, quats, scales, vm, Ks, W, H, opacities=opac)
    op = opac[None].expand(C, -1).contiguous()
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    r = _ref_isect(ref, m2, radii, dep, con, op, C, tw, th)
    tpg, ids, fl = gs.isect_tiles(m2, radii, dep, 16, tw, th, conics=con, opacities=op)
    off = gs.isect_offset_encode(ids, C, tw, th)
    # the reference evaluates the ellipse bounds with fast-math (approximate div / sqrt / log): a gaussian
    # whose bound falls within an ulp of a tile edge may gain or lose a tile.  Such tiles hold no pixel with
    # alpha >= 1/255, so images are unaffected; here the symmetric difference must stay below 1e-4.
    diff = (tpg != r[0]).float().mean()
    assert diff < 1e-4, f"tiles_per_gauss differs on {diff * 100:.4f}% of gaussians"
    if ids.numel() == r[1].numel():
        same = (ids == r[1]).float().mean()
        assert same > 0.999
    assert abs(ids.numel() - r[1].numel()) <= 1e-4 * ids.numel() + 2


@pytest.mark.parametrize("D,with_bg", [(3, False), (3, True), (4, False), (1, False)])
def test_raster_vs_reference(ref, gs, D, with_bg):
    W, H, C = 1280, 720, 1
    sc, vm, Ks = _scene(150000, W, H, C)
    means, quats, scales, opac = (_t(sc[k]) for k in ("means", "quats", "scales", "opacities"))
    radii, m2, dep, con, _ = gs.fully_fused_projection(means, None, quats, scales, vm, Ks, W, H, opacities=opac)
    op = opac[None].expand(C, -1).contiguous()
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    tpg, ids, fl, off = _ref_isect(ref, m2, radii, dep, con, op, C, tw, th)  # the reference's own lists for both
    g = torch.Generator(device=DEV).manual_seed(D)
    col = torch.rand((C, len(means), D), device=DEV, generator=g)
    bg = torch.rand((C, D), device=DEV, generator=g) if with_bg else None
    r_rc, r_ra, _, r_last = ref.rasterize_to_pixels_3dgs(m2, con, col, op, bg, None, W, H, 16, off, fl, False, False)
    ins = [x.detach().clone().requires_grad_(True) for x in (m2, con, col, op)]
    rc, ra = gs.rasterize_to_pixels(*ins, W, H, 16, off, fl, backgrounds=bg)
    err = (rc - r_rc).abs() - (1e-4 * r_rc.abs() + 1e-5)
    bad = (err.amax(-1) > 0).float().mean()
    # __expf and the alpha >= 1/255 / T <= 1e-4 decisions can flip on a few pixels between two correct
    # implementations; a flip changes a pixel by at most ~0.4 % of a colour
    assert bad < 2e-4, f"{bad * 100:.4f}% of pixels exceed rtol 1e-4 / atol 1e-5"
    assert (rc - r_rc).abs().max() < 2e-2 and (ra - r_ra).abs().max() < 2e-2
    v_rc = torch.randn(rc.shape, device=DEV, generator=g)
    v_ra = torch.randn(ra.shape, device=DEV, generator=g)
    rb = ref.rasterize_to_pixels_3dgs_bwd(m2, con, col, op, bg, None, off, fl, r_ra, r_last, W, H, 16, False, v_rc, v_ra, False)
    grads = torch.autograd.grad((rc * v_rc).sum() + (ra * v_ra).sum(), ins)
    # two runs of the reference itself differ by its atomic order; measure that spread and require ours
    # to be within a small multiple of it (and within absolute bounds)
    rb2 = ref.rasterize_to_pixels_3dgs_bwd(m2, con, col, op, bg, None, off, fl, r_ra, r_last, W, H, 16, False, v_rc, v_ra, False)
    for name, a, b, b2, lim in (
        ("v_means2d", grads[0], rb[1], rb2[1], 2e-4), ("v_conics", grads[1], rb[2], rb2[2], 2e-4),
        ("v_colors", grads[2], rb[3], rb2[3], 5e-5), ("v_opacities", grads[3], rb[4], rb2[4], 2e-4),
    ):
        spread = _rel(b2, b)
        rel = _rel(a, b)
        assert rel < max(lim, 20 * spread), f"{name}: rel L2 {rel:.3e} (reference run-to-run {spread:.3e})"


def test_full_path_vs_reference_1080p(ref, gs):
    """BASELINE configs[1]: ~100k gaussians, 1 camera 1080p, SH3, fwd + bwd vs the reference CUDA ops chained
    exactly as its orchestrator does (csrc/Rendering.cpp:976-1447)."""
    W, H, C = 1920, 1080, 1
    sc, vm, Ks = _scene(None, W, H, C)
    P = {k: _t(sc[k], True) for k in ("means", "quats", "scales", "opacities", "sh")}
    rc, ra, meta = gs.rasterization(P["means"], P["quats"], P["scales"], P["opacities"], P["sh"], vm, Ks, W, H, sh_degree=3, packed=False)
    means, quats, scales, opac, sh = (P[k].detach() for k in ("means", "quats", "scales", "opacities", "sh"))
    r_radii, r_m2, r_dep, r_con, _ = ref.projection_ewa_3dgs_fused(means, None, quats, scales, opac, vm, Ks, W, H, 0.3, 0.01, 1e10, 0.0, False, 0)
    rv = (r_radii > 0).all(-1)
    r_raw = ref.spherical_harmonics(3, means, vm, sh, rv, None, None, None, None)
    r_col = torch.clamp_min(r_raw + 0.5, 0.0) * rv[..., None]
    op = opac[None].expand(C, -1).contiguous()
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    _, ids, fl, off = _ref_isect(ref, r_m2, r_radii, r_dep, r_con, op, C, tw, th)
    r_rc, r_ra, _, r_last = ref.rasterize_to_pixels_3dgs(r_m2, r_con, r_col.contiguous(), op, None, None, W, H, 16, off, fl, False, False)
    err = (rc - r_rc).abs() - (1e-4 * r_rc.abs() + 1e-5)
    bad = (err.amax(-1) > 0).float().mean()
    assert bad < 1e-3, f"{bad * 100:.4f}% of pixels exceed rtol 1e-4 / atol 1e-5 vs the reference CUDA rasterizer"
    assert (rc - r_rc).abs().max() < 5e-2
    g = torch.Generator(device=DEV).manual_seed(3)
    v_rc, v_ra = torch.randn(rc.shape, device=DEV, generator=g), torch.randn(ra.shape, device=DEV, generator=g)
    ((rc * v_rc).sum() + (ra * v_ra).sum()).backward()
    rb = ref.rasterize_to_pixels_3dgs_bwd(r_m2, r_con, r_col.contiguous(), op, None, None, off, fl, r_ra, r_last, W, H, 16, False, v_rc, v_ra, False)
    v_m2, v_con, v_col, v_op = rb[1], rb[2], rb[3], rb[4]
    r_vcoef, r_vmeans_sh, _, _ = ref.spherical_harmonics_bwd(3, means, vm, sh, rv, None, None, None, None, (v_col * (r_col > 0)).contiguous(), True, False, False)
    pb = ref.projection_ewa_3dgs_fused_bwd(means, None, quats, scales, vm, Ks, W, H, 0.3, 0, r_radii, r_con, None, v_m2, torch.zeros_like(r_dep), v_con, None, False)
    assert _rel(P["sh"].grad, r_vcoef) < 1e-3
    assert _rel(P["opacities"].grad, v_op.sum(0)) < 1e-3
    assert _rel(P["means"].grad, pb[0] + r_vmeans_sh) < 2e-3
    assert _rel(P["quats"].grad, pb[2]) < 5e-3 and _rel(P["scales"].grad, pb[3]) < 5e-3


def test_adam_and_relocation_vs_reference(ref, gs):
    """Trainer-side ops against the reference kernels (csrc/AdamCUDA.cu, RelocationCUDA.cu, MCMCPerturbCUDA.cu)."""
    g = torch.Generator(device=DEV).manual_seed(3)
    N = 20000
    for shape in ((N, 3), (N, 16, 3)):
        p = torch.randn(shape, device=DEV, generator=g)
        grad = torch.randn(shape, device=DEV, generator=g) * 0.1
        m = torch.randn(shape, device=DEV, generator=g) * 0.01
        v = torch.rand(shape, device=DEV, generator=g) * 1e-3
        vis = torch.rand(N, device=DEV, generator=g) < 0.3
        a = [x.clone() for x in (p, m, v)]
        b = [x.clone() for x in (p, m, v)]
        ref.adam(a[0], grad, a[1], a[2], vis, 1e-2, 0.9, 0.999, 1e-8)
        gs.adam(b[0], grad, b[1], b[2], vis, 1e-2, 0.9, 0.999, 1e-8)
        for x, y in zip(a, b):
            torch.testing.assert_close(y, x, rtol=1e-5, atol=1e-7)  # the reference is built with fast-math (FMA: cancellation in m)
        assert torch.equal(b[0][~vis], p[~vis])
    # relocation
    n_max = 51
    binoms = torch.zeros((n_max, n_max), device=DEV)
    for n in range(n_max):
        for k in range(n + 1):
            binoms[n, k] = math.comb(n, k)
    opac = torch.rand(N, device=DEV, generator=g) * 0.98 + 0.01
    scales = torch.rand((N, 3), device=DEV, generator=g) * 0.1 + 1e-3
    ratios = torch.randint(1, 12, (N,), device=DEV, generator=g)
    ro, rs = ref.relocation(opac, scales, ratios.int(), binoms, n_max, 0.005)
    o, s = gs.compute_relocation(opac, scales, ratios.clone(), binoms, 0.005)
    torch.testing.assert_close(o, ro, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(s, rs, rtol=1e-4, atol=1e-8)


def test_packed_projection_vs_reference(ref, gs):
    """Two-pass compacting projection fwd + bwd against the reference's projection_ewa_3dgs_packed kernels."""
    W, H, C = 960, 540, 3
    sc, vm, Ks = _scene(70000, W, H, C)
    means, quats, scales, opac = (_t(sc[k]) for k in ("means", "quats", "scales", "opacities"))
    r = ref.projection_ewa_3dgs_packed(means, None, quats, scales, opac, vm, Ks, W, H, 0.3, 0.01, 1e10, 0.0, False, True, 0)
    o = gs.fully_fused_projection(means, None, quats, scales, vm, Ks, W, H, opacities=opac, packed=True, calc_compensations=True)
    rb, rc_, rg, rind, rrad, rm2, rdep, rcon, rcomp = r
    # visibility differs on a handful of borderline rows (fast-math radius): compare on the common (camera, gaussian) keys
    N = len(means)
    rk, ok = rc_ * N + rg, o[1] * N + o[2]
    assert abs(len(rk) - len(ok)) <= max(2, int(1e-4 * len(rk)))
    common, ri, oi = np.intersect1d(rk.cpu().numpy(), ok.cpu().numpy(), return_indices=True)
    assert len(common) > 0.999 * len(rk)
    ri, oi = torch.from_numpy(ri).to(DEV), torch.from_numpy(oi).to(DEV)
    assert (np.diff(ok.cpu().numpy()) > 0).all(), "rows must be in ascending (camera, gaussian) order"
    torch.testing.assert_close(o[5][oi], rm2[ri], rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(o[6][oi], rdep[ri], rtol=1e-5, atol=1e-6)
    assert _rel(o[7][oi], rcon[ri]) < 1e-4 and _rel(o[8][oi], rcomp[ri]) < 1e-5
    if len(rk) == len(ok):
        assert torch.equal(o[3], rind)
    # backward on the reference's own rows / conics (dense accumulation)
    g = torch.Generator(device=DEV).manual_seed(2)
    nnz = len(rk)
    v_m2, v_dep, v_con = (torch.randn(s, device=DEV, generator=g) for s in ((nnz, 2), (nnz,), (nnz, 3)))
    rbw = ref.projection_ewa_3dgs_packed_bwd(
        means, None, quats, scales, vm, Ks, W, H, 0.3, 0, False, rb, rc_, rg, rcon, None, v_m2, v_dep, v_con, None, True
    )
    L = gs._cabi.lib()
    from gsplat_b200._cabi import ptr, stream

    v_means, v_quats, v_scales, v_vm = (torch.empty_like(x) for x in (means, quats, scales, vm))
    rc = L.gsb200_projection_packed_bwd(
        1, C, N, nnz, ptr(means), None, ptr(quats), ptr(scales), ptr(vm), ptr(Ks), W, H, 0.3, 0, ptr(rb), ptr(rc_), ptr(rg),
        ptr(rcon), None, ptr(v_m2), 2, ptr(v_dep), 1, ptr(v_con), 3, None, 0, ptr(v_means), None, ptr(v_quats), ptr(v_scales),
        ptr(v_vm), stream(),
    )
    assert rc == 0
    assert _rel(v_means, rbw[0]) < 1e-4 and _rel(v_quats, rbw[2]) < 1e-4 and _rel(v_scales, rbw[3]) < 1e-4
    assert _rel(v_vm, rbw[4]) < 1e-3


def _stock_step(G, P, vm, Ks, W, H, deg, v_rc, v_ra):
    """fwd + bwd through a rasterization() callable G; returns render, alpha and the parameter gradients."""
    ins = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    rc, ra, _ = G(ins["means"], ins["quats"], ins["scales"], ins["opacities"], ins["sh"], vm, Ks, W, H, sh_degree=deg, packed=False)
    ((rc * v_rc).sum() + (ra * v_ra).sum()).backward()
    return rc.detach(), ra.detach(), {k: ins[k].grad.detach() for k in ins}


@pytest.mark.parametrize("cfg", ["cfg1_garden_256_sh0", "cfg2_100k_1080p_sh3"])
def test_stock_rasterization_with_gradient_spread(ref, gs, cfg):
    """BASELINE configs[0] / configs[1] against the reference's STOCK path -- the unmodified gsplat package of
    baseline/_ref: gsplat.rasterization() -> rasterization_3dgs orchestrator + its registered autograd.

    What is measured and written to gpurun_out/r02_grad_spread_<cfg>.json (committed under profiles/), per gradient tensor:
      * the reference's own run-to-run spread (two runs, same inputs);
      * ours vs the reference: relative L2, max |diff| / max |g|, and the fraction of elements outside
        rtol 1e-4 + atol 1e-5 * max|g| (north_star's per-element tolerance, the absolute part scaled to the tensor);
      * cfg1 only: BOTH implementations against the float64 CPU oracle (the ground truth of the reference's formulas).
    Measured on B200 (round 2): the reference's run-to-run spread is 3e-8 ... 1e-5 in relative L2, i.e. its float
    atomics are NOT what separates two correct implementations; ours and the reference differ by 2e-5 ... 4e-4 in
    relative L2 because the reference is a -use_fast_math build (approximate division / rsqrt / log in the projection
    and conic inversion) while the b200 per-gaussian kernels are bit-exact against the float32 oracle, and because
    discrete decisions (alpha >= 1/255, T <= 1e-4, radius / tile cuts) flip on a few (pixel, gaussian) pairs.
    Contract asserted here: render / alpha rtol 1e-4, atol 1e-5 on all but < 0.1 % decision-flip pixels; every
    gradient tensor within 1e-3 relative L2 of the reference and >= 99 % of its elements within
    rtol 1e-4 + atol 1e-5 * max|g|; on cfg1 ours must be at least as close to the float64 oracle as the reference is
    (factor 1.5 + 2e-5)."""
    import json

    from oracle import gso, refcuda

    if not refcuda.package_available():
        pytest.skip("baseline/_ref not installed")
    gsplat_ref = refcuda.import_package()
    if cfg.startswith("cfg1"):
        sc = scene.make_scene(sh_degree=0)
        W = H = 256
        vm, Ks, deg = _t(sc["viewmats"][:1]), _t(sc["Ks"][:1]), 0
    else:
        W, H, deg = 1920, 1080, 3
        sc, vm, Ks = _scene(100000, W, H, 1)
    P = {k: _t(sc[k]) for k in ("means", "quats", "scales", "opacities", "sh")}
    g = torch.Generator(device=DEV).manual_seed(11)
    v_rc, v_ra = torch.randn((1, H, W, 3), device=DEV, generator=g), torch.randn((1, H, W, 1), device=DEV, generator=g)
    r1 = _stock_step(gsplat_ref.rasterization, P, vm, Ks, W, H, deg, v_rc, v_ra)
    r2 = _stock_step(gsplat_ref.rasterization, P, vm, Ks, W, H, deg, v_rc, v_ra)
    o1 = _stock_step(gs.rasterization, P, vm, Ks, W, H, deg, v_rc, v_ra)
    oracle = None
    if cfg.startswith("cfg1"):
        f8 = lambda a: np.ascontiguousarray(a, np.float64)  # noqa: E731
        _, og = gso.rasterization_fwd_bwd(
            f8(sc["means"]), f8(sc["quats"]), f8(sc["scales"]), f8(sc["opacities"]), f8(sc["sh"]), f8(sc["viewmats"][:1]),
            f8(sc["Ks"][:1]), W, H, 0, f8(v_rc.cpu().numpy()), f8(v_ra.cpu().numpy()),
        )
        oracle = {k: torch.from_numpy(np.asarray(og["v_" + k])).to(DEV) for k in ("means", "quats", "scales", "opacities", "sh")}
    err = (o1[0] - r1[0]).abs() - (1e-4 * r1[0].abs() + 1e-5)
    bad = float((err.amax(-1) > 0).float().mean())
    bad_a = float((((o1[1] - r1[1]).abs() - (1e-4 * r1[1].abs() + 1e-5)) > 0).float().mean())
    report = {"cfg": cfg, "n_gaussians": int(P["means"].shape[0]), "render_pixels_out_of_1e-4_1e-5": bad, "alpha_pixels_out": bad_a, "grads": {}}
    fails = []
    for k in ("means", "quats", "scales", "opacities", "sh"):
        ref_g, ref_g2, our_g = r1[2][k], r2[2][k], o1[2][k]
        scale = float(ref_g.abs().max().clamp_min(1e-30))
        tol = 1e-4 * ref_g.abs() + 1e-5 * scale
        row = {
            "ref_run_to_run_rel_l2": _rel(ref_g2, ref_g), "ours_vs_ref_rel_l2": _rel(our_g, ref_g),
            "ref_run_to_run_max_over_scale": float((ref_g2 - ref_g).abs().max()) / scale,
            "ours_vs_ref_max_over_scale": float((our_g - ref_g).abs().max()) / scale,
            "ours_vs_ref_frac_outside_rtol1e-4_atol1e-5scale": float(((our_g - ref_g).abs() > tol).float().mean()),
        }
        if oracle is not None:
            og64 = oracle[k].reshape(ref_g.shape)
            row["ours_vs_oracle64_rel_l2"] = float((our_g.double() - og64).norm() / og64.norm())
            row["ref_vs_oracle64_rel_l2"] = float((ref_g.double() - og64).norm() / og64.norm())
            # (sh is reported but not asserted against the oracle: the garden colours contain exact zeros, whose SH
            # value sits exactly on the relu edge of clamp_min(sh + 0.5, 0) -- float32 and float64 evaluations fall on
            # different sides of it, for the reference and for us alike: both read 0.177 against the float64 oracle)
            if k != "sh" and not row["ours_vs_oracle64_rel_l2"] <= 1.5 * row["ref_vs_oracle64_rel_l2"] + 2e-5:
                fails.append(f"{k}: ours vs float64 oracle {row['ours_vs_oracle64_rel_l2']:.3e}, reference vs oracle {row['ref_vs_oracle64_rel_l2']:.3e}")
        report["grads"][k] = row
        if not row["ours_vs_ref_rel_l2"] <= 1e-3:
            fails.append(f"{k}: rel L2 {row['ours_vs_ref_rel_l2']:.3e} vs the reference")
        if not row["ours_vs_ref_frac_outside_rtol1e-4_atol1e-5scale"] <= 1e-2:
            fails.append(f"{k}: {row['ours_vs_ref_frac_outside_rtol1e-4_atol1e-5scale'] * 100:.3f} % of the elements outside rtol 1e-4 + atol 1e-5 max|g|")
    os.makedirs(os.path.join(os.path.dirname(__file__), "..", "gpurun_out"), exist_ok=True)
    with open(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", f"r02_grad_spread_{cfg}.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report))
    assert bad < 1e-3 and bad_a < 1e-3, report
    assert (o1[0] - r1[0]).abs().max() < 5e-2
    assert not fails, (fails, report)
