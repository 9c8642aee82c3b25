"""Pins the CPU oracle (oracle/) against the golden vectors produced by the REFERENCE's own
Python twins (tests/golden/make_golden.py) and against the reference's known-answer tables.
CPU only -- runs in the `-m "not gpu"` suite."""
import math
import os

import numpy as np
import pytest

from oracle import gso


def _load(golden_dir, name):
    d = np.load(os.path.join(golden_dir, name))
    return {k: d[k] for k in d.files}


def _close(a, b, rtol, atol, what=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert err.max() <= 0, f"{what}: max violation {err.max():.3e} (max abs diff {np.abs(a - b).max():.3e})"


# ---- bits_for_count: known answers of /root/reference/tests/cpp/test_mathutils.cpp:53-87
@pytest.mark.parametrize(
    "count,bits", [(0, 0), (1, 0), (2, 1), (3, 2), (4, 2), (5, 3), (7, 3), (8, 3), (9, 4)]
)
def test_bits_for_count_small(count, bits):
    assert gso.bits_for_count(count) == bits


def test_bits_for_count_powers_of_two():
    for k in range(1, 32):
        assert gso.bits_for_count(1 << k) == k
        assert gso.bits_for_count((1 << k) + 1) == k + 1
    for count in range(2, 4097):
        b = gso.bits_for_count(count)
        assert count - 1 <= (1 << b) - 1 and count - 1 > (1 << (b - 1)) - 1


# ---- quat_scale_to_covar_preci fwd/bwd vs _math._quat_scale_to_covar_preci (+ torch autograd)
@pytest.mark.parametrize("triu", [False, True])
def test_quat_scale(golden_dir, triu):
    g = _load(golden_dir, "ref_quat_scale.npz")
    t = "_triu" if triu else ""
    q, s = g["quats"].astype(np.float64), g["scales"].astype(np.float64)
    cov, pre = gso.quat_scale_to_covar_preci(q, s, True, True, triu)
    _close(cov, g["covars" + t], 1e-12, 1e-12, "covars")
    _close(pre, g["precis" + t], 1e-11, 1e-11, "precis")
    vq, vs = gso.quat_scale_to_covar_preci_bwd(q, s, triu, g["v_covars" + t], g["v_precis" + t])
    _close(vq, g["v_quats" + t], 1e-9, 1e-9, "v_quats")
    _close(vs, g["v_scales" + t], 1e-9, 1e-9, "v_scales")
    # float32 instantiation stays within fp32 round-off of the float64 reference
    cov32, pre32 = gso.quat_scale_to_covar_preci(g["quats"], g["scales"], True, True, triu)
    _close(cov32, g["covars" + t], 1e-5, 1e-6, "covars f32")


# ---- projection fwd/bwd vs _torch_impl._fully_fused_projection (+ autograd)
def test_projection(golden_dir):
    g = _load(golden_dir, "ref_projection.npz")
    W, H = int(g["width"]), int(g["height"])
    f8 = lambda k: g[k].astype(np.float64)  # noqa: E731
    # opacities=None -> fixed 3.33 sigma radius, exactly the torch twin's rule (_torch_impl.py:335-351)
    radii, m2, dep, con, comp = gso.fully_fused_projection(
        f8("means"), None, f8("quats"), f8("scales"), f8("viewmats"), f8("Ks"), W, H, 0.3, 0.01, 1e10, 0.0, True
    )
    ref_valid = (g["radii"] > 0).all(-1)
    valid = (radii > 0).all(-1)
    # the CUDA rule culls `<=`/`>=` where torch uses strict inequalities: never MORE visible than torch
    assert not (valid & ~ref_valid).any()
    assert (valid != ref_valid).mean() < 1e-3
    both = valid & ref_valid
    assert both.sum() > 1000
    assert (radii[both] == g["radii"][both]).all()
    _close(m2[both], g["means2d"][both], 1e-10, 1e-9, "means2d")
    _close(dep[both], g["depths"][both], 1e-12, 1e-12, "depths")
    _close(con[both], g["conics"][both], 1e-9, 1e-12, "conics")
    _close(comp[both], g["compensations"][both], 1e-9, 1e-12, "compensations")

    # backward with the golden's cotangents; use the golden's own radii as the valid mask.
    # (a) without the compensation cotangent: identical formulas -> round-off agreement
    radii_ref = g["radii"].astype(np.int32)
    v = gso.fully_fused_projection_bwd(
        f8("means"), None, f8("quats"), f8("scales"), f8("viewmats"), f8("Ks"), W, H, 0.3, radii_ref,
        g["conics"], None, f8("v_means2d"), f8("v_depths"), f8("v_conics"), None, True,
    )
    for name, a in (("v_means", v[0]), ("v_quats", v[2]), ("v_scales", v[3]), ("v_viewmats", v[4])):
        ref = g[name + "_nc"]
        _close(a, ref, 1e-9, 1e-11 * np.abs(ref).max(), name + " (no comp)")
    # (b) with it: the CUDA rule divides by (comp + 1e-6) (include/Utils.cuh:486), torch autograd by comp
    v = gso.fully_fused_projection_bwd(
        f8("means"), None, f8("quats"), f8("scales"), f8("viewmats"), f8("Ks"), W, H, 0.3, radii_ref,
        g["conics"], g["compensations"], f8("v_means2d"), f8("v_depths"), f8("v_conics"), f8("v_compensations"), True,
    )
    for name, a in (("v_means", v[0]), ("v_quats", v[2]), ("v_scales", v[3]), ("v_viewmats", v[4])):
        ref = g[name]
        _close(a, ref, 5e-4, 1e-6 * np.abs(ref).max(), name)


@pytest.mark.parametrize("model", ["ortho", "fisheye"])
def test_projection_camera_models(golden_dir, model):
    """Oracle ortho / fisheye projection fwd + bwd vs the reference's torch twin (_ortho_proj / _fisheye_proj +
    autograd, which differentiates the Jacobian too -- what the CUDA closed form of Utils.cuh:733-846 does)."""
    g = _load(golden_dir, f"ref_projection_{model}.npz")
    W, H = int(g["width"]), int(g["height"])
    f8 = lambda k: g[k].astype(np.float64)  # noqa: E731
    radii, m2, dep, con, comp = gso.fully_fused_projection(
        f8("means"), None, f8("quats"), f8("scales"), f8("viewmats"), f8("Ks"), W, H, 0.3, 0.01, 1e10, 0.0, True,
        camera_model=model,
    )
    ref_valid = (g["radii"] > 0).all(-1)
    valid = (radii > 0).all(-1)
    assert not (valid & ~ref_valid).any() and (valid != ref_valid).mean() < 2e-3
    both = valid & ref_valid
    assert both.sum() > 1000
    assert (radii[both] == g["radii"][both]).all()
    _close(m2[both], g["means2d"][both], 1e-10, 1e-8, "means2d")
    _close(dep[both], g["depths"][both], 1e-12, 1e-12, "depths")
    _close(con[both], g["conics"][both], 1e-8, 1e-11, "conics")
    _close(comp[both], g["compensations"][both], 1e-9, 1e-12, "compensations")
    v = gso.fully_fused_projection_bwd(
        f8("means"), None, f8("quats"), f8("scales"), f8("viewmats"), f8("Ks"), W, H, 0.3, g["radii"].astype(np.int32),
        g["conics"], None, f8("v_means2d"), f8("v_depths"), f8("v_conics"), None, True, camera_model=model,
    )
    # fisheye: the closed form treats d mean2d / d xyz as J (ideal map) while the twin differentiates the
    # eps-regularised expression, and x^2 carries a +1e-7: agreement to ~1e-6 relative, not round-off
    rtol = 1e-9 if model == "ortho" else 2e-5
    for name, a in (("v_means", v[0]), ("v_quats", v[2]), ("v_scales", v[3]), ("v_viewmats", v[4])):
        ref = g[name + "_nc"]
        _close(a, ref, rtol, (1e-11 if model == "ortho" else 2e-7) * np.abs(ref).max(), name)
    # float32 instantiation stays close
    r32 = gso.fully_fused_projection(g["means"], None, g["quats"], g["scales"], g["viewmats"], g["Ks"], W, H,
                                     0.3, 0.01, 1e10, 0.0, True, camera_model=model)
    b32 = (r32[0] > 0).all(-1) & ref_valid
    _close(r32[1][b32], g["means2d"][b32], 1e-4, 2e-3, "means2d f32")


def test_projection_f32_close_to_f64(golden_dir):
    g = _load(golden_dir, "ref_projection.npz")
    W, H = int(g["width"]), int(g["height"])
    r32 = gso.fully_fused_projection(g["means"], None, g["quats"], g["scales"], g["viewmats"], g["Ks"], W, H,
                                     0.3, 0.01, 1e10, 0.0, True)
    both = (r32[0] > 0).all(-1) & (g["radii"] > 0).all(-1)
    _close(r32[1][both], g["means2d"][both], 1e-4, 1e-3, "means2d f32")   # pixels: abs 1e-3 px
    _close(r32[3][both], g["conics"][both], 2e-3, 1e-5, "conics f32")     # tiny gaussians are ill-conditioned in f32


# ---- SH fwd/bwd vs _torch_impl._spherical_harmonics (+ autograd), degrees 0..4
@pytest.mark.parametrize("deg", [0, 1, 2, 3, 4])
def test_sh(golden_dir, deg):
    g = _load(golden_dir, "ref_sh.npz")
    means, vm = g["means"].astype(np.float64), g["viewmats"].astype(np.float64)
    cf = g[f"coeffs{deg}"].astype(np.float64)
    colors = gso.spherical_harmonics(deg, means, vm, cf)
    _close(colors, g[f"colors{deg}"], 1e-10, 1e-10, "colors")
    v_cf, v_m = gso.spherical_harmonics_bwd(deg, means, vm, cf, None, g[f"v_colors{deg}"])
    _close(v_cf, g[f"v_coeffs{deg}"], 1e-10, 1e-10, "v_coeffs")
    _close(v_m, g[f"v_means{deg}"], 1e-9, 1e-10, "v_means")
    c32 = gso.spherical_harmonics(deg, g["means"], g["viewmats"], g[f"coeffs{deg}"])
    _close(c32, g[f"colors{deg}"], 1e-4, 1e-5, "colors f32")


def test_sh_masks_and_padding(golden_dir):
    g = _load(golden_dir, "ref_sh.npz")
    means, vm = g["means"].astype(np.float64), g["viewmats"].astype(np.float64)
    cf4 = g["coeffs4"].astype(np.float64)  # K = 25 but only degree 2 used
    full = gso.spherical_harmonics(2, means, vm, cf4[:, :9])
    padded = gso.spherical_harmonics(2, means, vm, cf4)
    assert np.array_equal(full, padded)
    mask = np.zeros(full.shape[:-1], bool)
    mask[:, ::2] = True
    m = gso.spherical_harmonics(2, means, vm, cf4, mask)
    assert np.array_equal(m[mask], full[mask]) and (m[~mask] == 0).all()
    v_cf, _ = gso.spherical_harmonics_bwd(2, means, vm, cf4, mask, np.ones_like(full))
    assert (v_cf[:, 9:] == 0).all() and (v_cf[1::2] == 0).all()


# ---- tile intersection (AABB mode) and offsets vs _torch_impl._isect_tiles/_isect_offset_encode: exact
def test_isect_aabb_exact(golden_dir):
    g = _load(golden_dir, "ref_isect.npz")
    ts, tw, th = int(g["tile_size"]), int(g["tile_width"]), int(g["tile_height"])
    for dt in (np.float32, np.float64):
        tpg, ids, fl = gso.isect_tiles(g["means2d"].astype(dt), g["radii"], g["depths"].astype(dt), ts, tw, th)
        assert np.array_equal(tpg, g["tiles_per_gauss"])
        assert np.array_equal(ids, g["isect_ids"])
        assert np.array_equal(fl, g["flatten_ids"])
        off = gso.isect_offset_encode(ids, g["means2d"].shape[0], tw, th)
        assert np.array_equal(off, g["isect_offsets"])


def test_isect_empty():
    tpg, ids, fl = gso.isect_tiles(np.zeros((1, 4, 2), np.float32), np.zeros((1, 4, 2), np.int32),
                                   np.ones((1, 4), np.float32), 16, 3, 2)
    assert tpg.sum() == 0 and ids.shape == (0,) and fl.shape == (0,)
    assert (gso.isect_offset_encode(ids, 1, 3, 2) == 0).all()


# ---- compositing fwd/bwd vs the reference's accumulate() (+ autograd)
def test_compositing_vs_reference_accumulate(golden_dir):
    g = _load(golden_dir, "ref_accumulate.npz")
    W, H, ts = int(g["width"]), int(g["height"]), int(g["tile_size"])
    f8 = lambda k: g[k].astype(np.float64)  # noqa: E731
    rc, ra, li, mg = gso.rasterize_to_pixels(
        f8("means2d"), f8("conics"), f8("colors"), f8("opacities"), W, H, ts, g["isect_offsets"], g["flatten_ids"],
        f8("backgrounds"), None, True,
    )
    # tolerance 1e-7: the oracle clamps alpha at 0.99f (the CUDA constant), python at 0.99
    _close(rc, g["render_colors"], 0, 1e-7, "render_colors")
    _close(ra, g["render_alphas"], 0, 1e-7, "render_alphas")
    out = gso.rasterize_to_pixels_bwd(
        f8("means2d"), f8("conics"), f8("colors"), f8("opacities"), W, H, ts, g["isect_offsets"], g["flatten_ids"],
        ra, li, f8("v_render_colors"), f8("v_render_alphas"), f8("backgrounds"),
    )
    # the saturated (alpha == 0.99) pairs get zero geometry gradient in BOTH (clamp_max in torch,
    # the `opac*vis <= MAX_ALPHA` gate in CUDA); elsewhere identical formulas
    for k in ("v_means2d", "v_conics", "v_opacities", "v_colors"):
        scale = np.abs(g[k]).max()
        _close(out[k], g[k], 1e-6, 1e-7 * scale, k)
    _close(out["v_backgrounds"], g["v_backgrounds"], 1e-7, 1e-7, "v_backgrounds")


def test_compositing_f32_within_tolerance(golden_dir):
    """float32 oracle vs float64 reference on non-marginal pixels: rtol 1e-4 / atol 1e-5."""
    g = _load(golden_dir, "ref_accumulate.npz")
    W, H, ts = int(g["width"]), int(g["height"]), int(g["tile_size"])
    rc, ra, li, mg = gso.rasterize_to_pixels(
        g["means2d"], g["conics"], g["colors"], g["opacities"], W, H, ts, g["isect_offsets"], g["flatten_ids"],
        g["backgrounds"], None, True,
    )
    ok = mg > 1e-4
    assert ok.mean() > 0.98
    _close(rc[ok], g["render_colors"][ok], 1e-4, 1e-5, "render_colors f32")
    _close(ra[ok], g["render_alphas"][ok], 1e-4, 1e-5, "render_alphas f32")


# ---- AccuTile (no reference Python twin exists): properties that pin it
def test_accutile_is_conservative_and_tighter_than_aabb():
    rng = np.random.RandomState(7)
    N, W, H, ts = 1500, 200, 120, 16
    tw, th = math.ceil(W / ts), math.ceil(H / ts)
    m2 = np.stack([rng.random_sample(N) * (W + 40) - 20, rng.random_sample(N) * (H + 40) - 20], -1)[None]
    sx, sy = rng.random_sample(N) * 9 + 0.6, rng.random_sample(N) * 9 + 0.6
    rho = rng.random_sample(N) * 1.8 - 0.9
    cov = np.stack([sx * sx, rho * sx * sy, sy * sy], -1)
    det = cov[:, 0] * cov[:, 2] - cov[:, 1] ** 2
    con = np.stack([cov[:, 2] / det, -cov[:, 1] / det, cov[:, 0] / det], -1)[None]
    op = (rng.random_sample(N) * 0.99 + 0.01)[None]
    ext = np.minimum(3.33, np.sqrt(np.maximum(2 * np.log(np.maximum(op[0] * 255, 1.0)), 0)))
    radii = np.stack([np.ceil(ext * sx), np.ceil(ext * sy)], -1).astype(np.int32)[None]
    radii[0, op[0] < 1 / 255.0] = 0
    dep = (rng.random_sample(N) + 0.1)[None]
    f4 = lambda a: np.ascontiguousarray(a, np.float32)  # noqa: E731
    tpg_a, ids_a, fl_a = gso.isect_tiles(f4(m2), radii, f4(dep), ts, tw, th, True, f4(con), f4(op))
    tpg_b, ids_b, fl_b = gso.isect_tiles(f4(m2), radii, f4(dep), ts, tw, th, True)
    tile_bits = gso.bits_for_count(tw * th)
    pa = set(zip(fl_a.tolist(), ((ids_a >> 32) & ((1 << tile_bits) - 1)).tolist()))
    pb = set(zip(fl_b.tolist(), ((ids_b >> 32) & ((1 << tile_bits) - 1)).tolist()))
    assert pa <= pb, "AccuTile emitted a tile outside the radius AABB"
    assert len(pa) < 0.9 * len(pb), "AccuTile should prune a good share of AABB tiles"
    # conservativeness: every (gaussian, tile) holding a pixel with alpha >= 1/255 must be emitted
    ys, xs = np.mgrid[0:H, 0:W]
    px, py = xs + 0.5, ys + 0.5
    tile_of = (ys // ts) * tw + (xs // ts)
    for n in range(0, N, 7):
        if radii[0, n, 0] <= 0:
            continue
        dx, dy = m2[0, n, 0] - px, m2[0, n, 1] - py
        sig = 0.5 * (con[0, n, 0] * dx * dx + con[0, n, 2] * dy * dy) + con[0, n, 1] * dx * dy
        alpha = np.minimum(0.99, op[0, n] * np.exp(-sig))
        hit = (alpha >= 1 / 255.0 * (1 + 1e-5))
        for t in np.unique(tile_of[hit]):
            assert (n, int(t)) in pa, f"gaussian {n} tile {t} has a visible pixel but was not emitted"
    # sortedness + offsets partition
    assert (np.diff(ids_a) >= 0).all()
    off = gso.isect_offset_encode(ids_a, 1, tw, th).reshape(-1)
    assert off[0] == 0 and (np.diff(off) >= 0).all() and off[-1] <= len(ids_a)


def test_full_pipeline_small_runs_and_is_consistent():
    """f32 vs f64 oracle on a tiny crop of the garden scene through the whole path."""
    from tests import scene

    sc = scene.make_scene(n_max=4000, sh_degree=1)
    W = H = 64
    Ks = sc["Ks"][:1].copy()
    Ks[0, 0, 2], Ks[0, 1, 2] = W / 2, H / 2
    rng = np.random.RandomState(0)
    v_rc, v_ra = rng.standard_normal((1, H, W, 3)), rng.standard_normal((1, H, W, 1))
    outs = {}
    for dt in (np.float32, np.float64):
        c = lambda k: sc[k].astype(dt)  # noqa: E731
        outs[dt] = gso.rasterization_fwd_bwd(
            c("means"), c("quats"), c("scales"), c("opacities"), c("sh"), sc["viewmats"][:1].astype(dt), Ks.astype(dt),
            W, H, 1, v_rc.astype(dt), v_ra.astype(dt),
        )
    f32, f64 = outs[np.float32][0], outs[np.float64][0]
    assert f32["flatten_ids"].shape[0] > 500
    ok = f64["margins"] > 1e-3
    same_list = np.array_equal(f32["flatten_ids"], f64["flatten_ids"])
    if same_list:
        _close(f32["render_colors"][ok], f64["render_colors"][ok], 1e-3, 1e-4, "pipeline colors")


# ---- MCMC strategy ops vs the reference's own Python restatements
def test_mcmc_ops(golden_dir):
    g = _load(golden_dir, "ref_mcmc.npz")
    f8 = lambda k: g[k].astype(np.float64)  # noqa: E731
    for tag, mo in (("", 0.005), ("_mo0", 0.0)):
        no, ns = gso.compute_relocation(f8("opacities"), f8("scales"), g["ratios"], f8("binoms"), mo)
        _close(no, g["new_opacities" + tag], 1e-12, 1e-14, "new_opacities")
        _close(ns, g["new_scales" + tag], 1e-9, 1e-12, "new_scales")
    newp = gso.mcmc_perturb_positions(f8("positions"), f8("quats"), f8("scales_log"), f8("opacities_logit"), f8("noise"),
                                      float(g["noise_scale"]), float(g["t"]), float(g["k"]))
    _close(newp, g["new_positions"], 1e-10, 1e-10, "new_positions")


def test_adam_l1_and_packed_bookkeeping(golden_dir):
    """Trainer-side oracle pieces against independent float64 torch formulas (CPU)."""
    import torch

    rng = np.random.RandomState(2)
    p, g = rng.standard_normal((50, 3)), rng.standard_normal((50, 3)) * 0.1
    m, v = rng.standard_normal((50, 3)) * 0.01, rng.random_sample((50, 3)) * 1e-3
    vis = rng.random_sample(50) < 0.5
    op, om, ov = gso.adam(p, g, m, v, vis, 1e-2, 0.9, 0.999, 1e-8)
    tp, tg, tm, tv = (torch.tensor(x) for x in (p, g, m, v))
    tm2 = 0.9 * tm + 0.1 * tg
    tv2 = 0.999 * tv + 0.001 * tg * tg
    tp2 = tp - 1e-2 * tm2 / (tv2.sqrt() + 1e-8)  # Adam without bias correction (csrc/AdamCUDA.cu:57-64)
    sel = torch.tensor(vis)[:, None]
    _close(op, torch.where(sel, tp2, tp).numpy(), 1e-12, 1e-14, "adam param")
    _close(om, torch.where(sel, tm2, tm).numpy(), 1e-12, 1e-14, "adam exp_avg")
    _close(ov, torch.where(sel, tv2, tv).numpy(), 1e-12, 1e-16, "adam exp_avg_sq")
    a, b = rng.standard_normal((4, 5, 3)), rng.standard_normal((4, 5, 3))
    b.reshape(-1)[::5] = a.reshape(-1)[::5]
    ta = torch.tensor(a, requires_grad=True)
    tl = (ta - torch.tensor(b)).abs().mean()
    tl.backward()
    loss, grad = gso.l1_loss(a, b)
    assert abs(loss - float(tl.detach())) < 1e-15
    _close(grad, ta.grad.numpy(), 1e-12, 1e-15, "l1 grad")
    # packed rows == the dense rows with radii > 0, ascending (camera, gaussian), CSR indptr per camera
    g = _load(golden_dir, "ref_projection.npz")
    W, H = int(g["width"]), int(g["height"])
    args = (g["means"], None, g["quats"], g["scales"], g["viewmats"], g["Ks"], W, H, 0.3, 0.01, 1e10, 0.0, True)
    d = gso.fully_fused_projection(*args)
    pk = gso.fully_fused_projection_packed(*args)
    vis = (d[0] > 0).all(-1)
    assert pk[3].tolist() == [0] + np.cumsum(vis.sum(-1)).tolist() and len(pk[0]) == vis.sum()
    key = pk[1] * vis.shape[1] + pk[2]
    assert (np.diff(key) > 0).all() and (pk[0] == 0).all()
    assert np.array_equal(pk[5], d[1][vis]) and np.array_equal(pk[8], d[4][vis])
