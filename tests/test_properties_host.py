"""Property-based tests of the PhysicsEngine host maths, mirroring the reference's fast-check suite on its TS Kerr
mirror (src/__tests__/physics/kerr-metric.test.ts:19-292: random mass in [0.1, 10], spin in [-1, 1], orderings of the
characteristic radii) and its Rust inequality tests (physics/redshift.rs:138-171, physics/disk.rs:226-308,
physics/shadow.rs:256-336) — here against the C ABI (CPU-side entry points; no GPU needed)."""
import math

import numpy as np
from hypothesis import given, settings, strategies as st

masses = st.floats(min_value=0.1, max_value=10.0, allow_nan=False)
spins = st.floats(min_value=-1.0, max_value=1.0, allow_nan=False)


@settings(max_examples=150, deadline=None)
@given(m=masses, a=spins)
def test_characteristic_radii_orderings(built, m, a):
    e = built.PhysicsEngine(m, a)
    rh, isco, rph = e.compute_horizon(), e.compute_isco(), e.compute_photon_sphere()
    assert m <= rh <= 2.0 * m + 1e-12                       # r+ in [M, 2M]
    assert rh - 1e-7 * m <= rph <= 4.0 * m + 1e-9               # photon orbit between r+ and 4M (3M at a = 0; 4M retrograde extremal)
    if a >= 0.0:
        assert rph - 1e-7 * m <= isco <= 6.0 * m + 1e-9          # r+ <= r_ph <= r_isco <= 6M: prograde orbits move inwards with spin
        assert rph <= 3.0 * m + 1e-9
    assert abs(rh - (m + math.sqrt(max(m * m - (a * m) ** 2, 0.0)))) < 1e-12
    # Kerr::isco(prograde = true) is even in a* (kerr.rs:100-123: it is prograde with respect to the hole's own spin),
    # while photon_sphere() takes the signed a* (kerr.rs:125-140) -- mirrored as is
    e2 = built.PhysicsEngine(m, -a)
    assert abs(e2.compute_horizon() - rh) < 1e-12 and abs(e2.compute_isco() - isco) < 1e-9 * m
    if a > 0.0:
        assert e2.compute_photon_sphere() >= rph - 1e-9      # prograde <= retrograde
    e.free(); e2.free()


@settings(max_examples=60, deadline=None)
@given(m=masses, a=st.floats(min_value=0.0, max_value=0.998), k=st.floats(min_value=1.05, max_value=40.0))
def test_redshift_and_flux_signs(built, m, a, k):
    e = built.PhysicsEngine(m, a)
    isco = e.compute_isco()
    r = isco * k
    assert e.compute_disk_flux(isco * 0.99) == 0.0 and e.compute_disk_flux(r) > 0.0     # disk.rs:226-245
    g0 = e.compute_g_factor(r, 0.0)
    assert 0.0 < g0 < 1.0                                    # pure gravitational + transverse redshift
    lam = 0.5 * math.sqrt(r * m)
    assert e.compute_g_factor(r, +lam) > g0 > e.compute_g_factor(r, -lam)               # approaching > receding (redshift.rs:160-171)
    assert abs(e.compute_g_factor(1000.0 * m, 0.0) - 1.0) < 5e-3                          # -> 1 far away (:138-146)
    assert e.compute_dilation(r) >= 1.0                       # dt/dtau of a static observer (lib.rs:97-105; 100 inside the ergoregion)
    e.free()


@settings(max_examples=40, deadline=None)
@given(m=masses, a=st.floats(min_value=0.05, max_value=0.999), th=st.floats(min_value=0.2, max_value=math.pi - 0.2))
def test_bardeen_curve_shape(built, m, a, th):
    e = built.PhysicsEngine(m, a)
    c = e.compute_shadow_curve(th, 32).reshape(-1, 2).astype(np.float64)
    assert c.shape == (64, 2) and np.all(np.isfinite(c))
    np.testing.assert_allclose(c[:32, 0], c[:31:-1, 0], rtol=1e-6, atol=1e-6)            # mirror symmetry in beta
    np.testing.assert_allclose(c[:32, 1], -c[:31:-1, 1], rtol=1e-6, atol=1e-6)
    width = c[:, 0].max() - c[:, 0].min()
    b = 3.0 * math.sqrt(3.0) * m
    assert 1.4 * b < width < 2.05 * b                       # between the extremal-Kerr and the Schwarzschild widths
    assert c[:, 0].mean() * a > -1e-9 or abs(math.sin(th)) < 0.3   # the curve shifts with the spin's sense (shadow.rs:300-336)
    sh = e.compute_shadow_shift(th)
    assert sh[0] <= c[:, 0].min() + 1e-5 and sh[1] >= c[:, 0].max() - 1e-5
    e.free()
