"""CPU model of phase B of the version-2 compositing backward (gsplat_b200/csrc/raster.cu, raster_bwd2_kernel::flush):
the per-gaussian gradients are recovered from the moments of o = vis * v_alpha about the centre of each 8x2 half of
the warp's 8x4 pixel block.  This restates the kernel's algebra in numpy (float64) and checks it against the direct
per-pixel sums of the reference's formulas (csrc/RasterizeToPixels3DGSDevice.cuh:104-173):

    v_sigma = -opacity * o,   v_xy = sum v_sigma * (conic . d),   v_conic = sum v_sigma * (dx^2/2, dx dy, dy^2/2),
    v_opacity = sum o,        d = mean - pixel centre.
"""
import numpy as np


def _direct(mx, my, a, b, c, op, o, px, py):
    dx, dy = mx - px, my - py
    vs = -op * o
    return np.array([
        (vs * (a * dx + b * dy)).sum(), (vs * (b * dx + c * dy)).sum(), (0.5 * vs * dx * dx).sum(), (vs * dx * dy).sum(),
        (0.5 * vs * dy * dy).sum(), o.sum(),
    ])


def _kernel_model(mx, my, a, b, c, op, o, bx0, by0):
    """o: [32] in lane order (lane = y * 8 + x inside the 8x4 block whose first pixel is (bx0, by0))."""
    cx, cy = bx0 + 4.0, by0 + 2.0
    out = np.zeros(6)
    for half in (0, 1):
        Dx, Dy = mx - cx, my - (cy - 1.0 + 2.0 * half)
        M = np.zeros(5)
        for i in range(16):
            xi, eta = (i & 7) - 3.5, (i >> 3) - 0.5
            v = o[half * 16 + i]
            M += v * np.array([1.0, xi, eta, xi * xi, xi * eta])
        U0, U1, U2, U3, U4 = (-op * M).tolist()
        U5 = 0.25 * U0  # eta^2 == 0.25 on both rows of a half
        Sdx, Sdy = Dx * U0 - U1, Dy * U0 - U2
        out += np.array([
            a * Sdx + b * Sdy, b * Sdx + c * Sdy, 0.5 * (Dx * (Sdx - U1) + U3), Dx * Sdy - Dy * U1 + U4,
            0.5 * (Dy * (Sdy - U2) + U5), M[0],
        ])
    return out


def test_half_block_moments_reproduce_the_direct_sums():
    rng = np.random.RandomState(0)
    for _ in range(200):
        bx0, by0 = 8 * rng.randint(0, 200), 4 * rng.randint(0, 200)
        mx, my = bx0 + rng.uniform(-300, 300), by0 + rng.uniform(-300, 300)
        a, c = rng.uniform(0.01, 2.0, 2)
        b = rng.uniform(-0.9, 0.9) * np.sqrt(a * c)
        op = rng.uniform(0.01, 1.0)
        o = rng.standard_normal(32) * (rng.random_sample(32) < 0.8)  # some pixels invalid (o = 0)
        lane = np.arange(32)
        px, py = bx0 + (lane & 7) + 0.5, by0 + (lane >> 3) + 0.5
        ref = _direct(mx, my, a, b, c, op, o, px, py)
        got = _kernel_model(mx, my, a, b, c, op, o, bx0, by0)
        scale = np.abs(ref).max() + 1e-30
        assert np.abs(got - ref).max() <= 1e-9 * scale + 1e-12 * (1 + abs(mx) + abs(my)) ** 2, (got, ref)


def test_scalar_behind_sum_equals_the_per_channel_buffer():
    """Phase A keeps sum_behind fac_j (c_j . v_c) as ONE scalar; the reference keeps buffer[k] = sum_behind fac_j c_j[k]
    and dots it with v_render_c afterwards -- the same number."""
    rng = np.random.RandomState(1)
    n, D = 40, 3
    col, fac = rng.random_sample((n, D)), rng.random_sample(n) * 0.1
    v_c = rng.standard_normal(D)
    buffer, behind = np.zeros(D), 0.0
    for j in range(n - 1, -1, -1):  # back to front
        assert abs(buffer @ v_c - behind) < 1e-12
        buffer += col[j] * fac[j]
        behind += (col[j] @ v_c) * fac[j]
