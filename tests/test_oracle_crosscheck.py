"""oracle/ (C++) against tests/pyref.py (an independently written pure-Python restatement of the same Rust / WGSL
lines) on the pieces no reference test and no SURVEY known-answer pins: fixed-step steppers, camera -> state, LUT
texels and samplers, g-factor, composite RGBA of whole pixels. CPU only; tens of rays, a few texels."""
import math

import numpy as np
import pytest

import pyref

SPIN = float(np.float32(0.999))
KS = 1


@pytest.fixture(scope="module")
def cam_small():
    from gravitas_b200 import camera
    cam, _ = camera.default_camera(64, 36)
    return cam


def test_steppers_bitwise(oracle):
    rng = np.random.default_rng(11)
    for _ in range(20):
        s = [0.0, rng.uniform(3, 40), rng.uniform(0.3, 2.8), rng.uniform(0, 6), -1.0, rng.uniform(-1, 1),
             rng.uniform(-3, 3), rng.uniform(-5, 5)]
        a = 0.9
        s = pyref.renormalize(1.0, a, s)
        np.testing.assert_array_equal(oracle.renormalize(1.0, a, KS, s), s)            # idempotent and identical
        np.testing.assert_array_equal(oracle.rhs(1.0, a, KS, s), pyref.rhs(1.0, a, s))
        assert oracle.hamiltonian(1.0, a, KS, s) == pyref.hamiltonian(1.0, a, s)
        for h in (0.05, 0.4, 1.0):
            np.testing.assert_array_equal(oracle.step_symplectic(1.0, a, KS, s, h), pyref.step_symplectic(1.0, a, s, h))
            np.testing.assert_array_equal(oracle.step_rk4(1.0, a, KS, s, h), pyref.step_rk4(1.0, a, s, h))


def test_camera_ray_bitwise(oracle, cam_small):
    opts = oracle.Options.default()
    rp, keep = oracle.make_render_params(64, 36, 1.0, SPIN, opts)
    for (x, y) in [(0, 0), (63, 35), (32, 18), (5, 30), (40, 2)]:
        np.testing.assert_array_equal(oracle.camera_ray(cam_small, rp, x, y), pyref.camera_ray(cam_small, 64, 36, x, y))


def test_g_factor_and_radii_bitwise(oracle):
    L = oracle.lib()
    for spin in (0.0, 0.5, -0.7, 0.999):
        assert L.orc_isco(1.0, spin, 1) == pyref.isco(1.0, spin)
        assert L.orc_event_horizon(1.0, spin, 0) == pyref.horizon(1.0, spin)
        for r, lam in ((3.0, 1.0), (7.5, -2.0), (20.0, 4.0), (1.3, 0.2)):
            assert L.orc_g_factor(r, 1.0, spin, lam) == pyref.g_factor(r, 1.0, spin, lam)


def test_spectrum_texels_bitwise(oracle):
    W, H = 24, 6
    lut = oracle.spectrum_lut(W, H, 1e7).reshape(H, W, 4)
    for (x, y) in [(0, 0), (1, 0), (5, 2), (23, 5), (12, 3), (23, 0)]:
        assert tuple(float(v) for v in lut[y, x, :3]) == pyref.spectrum_texel(x, y, W, H, 1e7)
        assert lut[y, x, 3] == 1.0


def test_composite_pixels(oracle, cam_small):
    """Whole pixels: same colour (to 1e-13 relative: the only non-identical operation order is inside pow/LUT
    arithmetic), same termination and step count."""
    W, H, steps, LW, LH = 64, 36, 200, 32, 8
    lut = oracle.spectrum_lut(LW, LH, 1e7)
    td = oracle.disk_lut(1.0, SPIN)
    opts = oracle.Options.default(method=oracle.METHOD_SYMPLECTIC, step_rule=1, max_steps=steps)
    rp, keep = oracle.make_render_params(W, H, 1.0, SPIN, opts, spectrum=lut, spec_w=LW, spec_h=LH, tdisk=td)
    ref = oracle.render(cam_small, rp)
    lit = 0
    for (x, y) in [(3, 3), (20, 10), (31, 17), (33, 18), (40, 22), (10, 25), (50, 14), (32, 20), (28, 19), (60, 30)]:
        rgb, term, n = pyref.render_pixel(cam_small, W, H, x, y, 1.0, SPIN, steps, lut, LW, LH, td)
        assert term == ref["term"][y, x] and n == ref["steps"][y, x], (x, y)
        np.testing.assert_allclose(rgb, ref["rgba"][y, x, :3], rtol=1e-13, atol=1e-300)
        lit += any(c > 0 for c in rgb)
    assert lit >= 4
