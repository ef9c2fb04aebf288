//! gen_fixtures.rs — pins the CPU oracle (oracle/gravitas_oracle.hpp) to the REFERENCE itself.
//!
//! This is an `examples/` program for the reference's own crate `gravitas-core` (library name `gravitas`). It calls
//! the reference's public API on fixed inputs and writes inputs + outputs as JSON. It contains no physics of its own.
//! The image this repository is built in has no cargo/rustc, so the file ships as a recipe; a maintainer with a Rust
//! toolchain runs (see oracle/ref_fixtures/README.md):
//!
//!   cp oracle/ref_fixtures/gen_fixtures.rs <reference>/physics-engine/gravitas-core/examples/
//!   (cd <reference>/physics-engine && cargo run --release -p gravitas-core --example gen_fixtures) \
//!       > tests/golden/ref_gravitas_core.json
//!
//! tests/test_oracle_ref_fixtures.py then replays every recorded input through the oracle (and, on a GPU box,
//! through the CUDA kernels) and compares with the recorded reference outputs.
//!
//! Reference entry points exercised (gravitas-core/src): metric/kerr.rs:48-123 (new, kerr_schild, event_horizon,
//! photon_sphere, isco), geodesic/hamiltonian.rs:13-35, invariants/mod.rs:25-37, invariants/renormalization.rs:13-45,
//! geodesic/integrator.rs:53-226 (AdaptiveStepper, adaptive_rkf45_step, step_rk4, step_symplectic),
//! geodesic/mod.rs:180-253 (integrate), physics/redshift.rs:65-95, physics/spectrum.rs:76-102,
//! physics/disk.rs:90-201, physics/shadow.rs:81-183.
//!
//! JSON schema ("schema": 1): every f64 is printed with Rust's shortest round-trip formatting ({:?}); non-finite
//! values are written as the strings "nan" / "inf" / "-inf". f32 arrays are widened to f64 exactly first.

use gravitas::geodesic::{
    adaptive_rkf45_step, get_state_derivative, integrate, step_rk4, step_symplectic, AdaptiveStepper, GeodesicState,
    IntegrationMethod, IntegrationOptions,
};
use gravitas::invariants::{hamiltonian, renormalize_null};
use gravitas::metric::{Kerr, Metric, Orbit};
use gravitas::physics::{disk, redshift, shadow, spectrum};

fn num(v: f64) -> String {
    if v.is_nan() {
        "\"nan\"".to_string()
    } else if v.is_infinite() {
        if v > 0.0 { "\"inf\"".to_string() } else { "\"-inf\"".to_string() }
    } else {
        format!("{:?}", v)
    }
}
fn arr(v: &[f64]) -> String {
    format!("[{}]", v.iter().map(|x| num(*x)).collect::<Vec<_>>().join(","))
}
fn arr32(v: &[f32]) -> String {
    format!("[{}]", v.iter().map(|x| num(*x as f64)).collect::<Vec<_>>().join(","))
}
fn st(s: &GeodesicState) -> String {
    arr(&[s.x[0], s.x[1], s.x[2], s.x[3], s.p[0], s.p[1], s.p[2], s.p[3]])
}
fn metric_of(m: f64, a: f64, ks: bool) -> Kerr {
    if ks { Kerr::kerr_schild(m, a) } else { Kerr::new(m, a) }
}

/// A deterministic fan of null rays from a camera-like position: (r0, theta0) fixed, direction angles on an n x n
/// lattice (alpha: towards/around the hole in the phi direction, beta: in the theta direction). p_t = -1.
fn ray_fan(r0: f64, theta0: f64, n: usize, half_width: f64) -> Vec<GeodesicState> {
    let mut out = Vec::new();
    for j in 0..n {
        for i in 0..n {
            let alpha = half_width * (2.0 * (i as f64 + 0.5) / n as f64 - 1.0);
            let beta = half_width * (2.0 * (j as f64 + 0.5) / n as f64 - 1.0);
            // local direction (towards the hole = -e_r), flat-space far-field momenta as compute.wgsl.ts:174-187 builds them
            let dr = -(alpha.cos() * beta.cos());
            let dth = beta.sin();
            let dph = alpha.sin() * beta.cos();
            let s = theta0.sin();
            out.push(GeodesicState::null_ray(r0, theta0, std::f64::consts::PI, dr, dth * r0, dph * r0 * s));
        }
    }
    out
}

fn integrate_case(name: &str, m: f64, a: f64, ks: bool, opts: &IntegrationOptions, method: &str, step: f64, rays: &[GeodesicState]) -> String {
    let bh = metric_of(m, a, ks);
    let mut items = Vec::new();
    for r in rays {
        let t = integrate(r, &bh, opts);
        items.push(format!(
            "{{\"in\":{},\"out\":{},\"termination\":{},\"steps\":{},\"max_drift\":{}}}",
            st(r), st(&t.final_state), t.termination as u32, t.steps_taken, num(t.max_hamiltonian_drift)
        ));
    }
    format!(
        "{{\"name\":\"{}\",\"mass\":{},\"spin\":{},\"kerr_schild\":{},\"method\":\"{}\",\"step_size\":{},\"tolerance\":{},\"initial_step\":{},\"max_steps\":{},\"escape_radius\":{},\"renormalize_interval\":{},\"rays\":[{}]}}",
        name, num(m), num(a), ks, method, num(step), num(opts.tolerance), num(opts.initial_step), opts.max_steps,
        num(opts.escape_radius), opts.renormalize_interval, items.join(",")
    )
}

fn main() {
    let half_pi = std::f64::consts::FRAC_PI_2;
    let spin32 = 0.999f32 as f64; // PhysicsParams carries spin as f32 (src/types/webgpu.ts:42-64)
    let mut sections: Vec<String> = Vec::new();

    // ---- radii (metric/kerr.rs:31-33,507-554 hold a few of these as doctest constants) ----
    {
        let mut v = Vec::new();
        for &a in &[0.0, 0.5, 0.9, 0.998, spin32, 1.0, -0.7] {
            let bh = Kerr::new(1.0, a);
            v.push(format!(
                "{{\"spin\":{},\"horizon\":{},\"photon_sphere\":{},\"isco_pro\":{},\"isco_retro\":{}}}",
                num(a), num(bh.event_horizon()), num(bh.photon_sphere()), num(bh.isco(Orbit::Prograde)), num(bh.isco(Orbit::Retrograde))
            ));
        }
        sections.push(format!("\"radii\":[{}]", v.join(",")));
    }

    // ---- pointwise: rhs, H, renormalize_null, one RKF45 attempt, one controller step, RK4, implicit midpoint ----
    {
        let states = [
            (1.0, 0.9, GeodesicState::null_ray(20.0, half_pi, 0.0, -1.0, 0.0, 3.5)),
            (1.0, spin32, GeodesicState::new(0.0, 10.0, 1.2, 0.3, -1.0, -0.9, 0.5, 2.0)),
            (1.0, spin32, GeodesicState::new(0.0, 3.0, 0.4, 1.0, -1.0, 0.3, -1.5, 0.8)),
            (1.0, 0.0, GeodesicState::new(0.0, 30.0, 1.6929693744, 3.14159, -1.0, -0.97, 2.0, 4.0)),
            (2.5, -0.6, GeodesicState::new(0.0, 12.0, 2.2, 0.0, -1.0, -0.8, 1.0, -5.0)),
            (1.0, spin32, GeodesicState::new(0.0, 1.2, 1.0, 0.0, -1.0, -2.0, 0.3, 1.0)),
        ];
        let mut v = Vec::new();
        for (m, a, s0) in states.iter() {
            for &ks in &[false, true] {
                let bh = metric_of(*m, *a, ks);
                let d = get_state_derivative(s0, &bh);
                let h0 = hamiltonian(s0, &bh);
                let mut sr = *s0;
                renormalize_null(&mut sr, &bh);
                let (s45, err) = adaptive_rkf45_step(&sr, &bh, 0.1);
                let mut sa = sr;
                let mut stepper = AdaptiveStepper::new(1e-8);
                let h_next = stepper.step(&mut sa, &bh, 0.5);
                let mut s4 = sr;
                step_rk4(&mut s4, &bh, 0.1);
                let mut sm = sr;
                step_symplectic(&mut sm, &bh, 0.1);
                v.push(format!(
                    "{{\"mass\":{},\"spin\":{},\"kerr_schild\":{},\"state\":{},\"rhs\":{},\"hamiltonian\":{},\"renormalized\":{},\"hamiltonian_after\":{},\"rkf45_h\":0.1,\"rkf45_state\":{},\"rkf45_error\":{},\"stepper_h_try\":0.5,\"stepper_tol\":1e-8,\"stepper_state\":{},\"stepper_h_next\":{},\"rk4_h\":0.1,\"rk4_state\":{},\"symplectic_h\":0.1,\"symplectic_state\":{}}}",
                    num(*m), num(*a), ks, st(s0), st(&d), num(h0), st(&sr), num(hamiltonian(&sr, &bh)), st(&s45), num(err),
                    st(&sa), num(h_next), st(&s4), st(&sm)
                ));
            }
        }
        sections.push(format!("\"pointwise\":[{}]", v.join(",")));
    }

    // ---- integrate(): the doctest ray, and 8x8 ray fans for the BASELINE configs' schemes ----
    {
        let mut v = Vec::new();
        let doc = [GeodesicState::null_ray(20.0, half_pi, 0.0, -1.0, 0.0, 3.5)];
        let def = IntegrationOptions::default();
        v.push(integrate_case("doctest_bl", 1.0, 0.9, false, &def, "rkf45", 0.0, &doc));
        v.push(integrate_case("doctest_ks", 1.0, 0.9, true, &def, "rkf45", 0.0, &doc));
        let theta0 = 97.0f64.to_radians();
        let fan = ray_fan(30.0, theta0, 8, 0.35);
        // config 1: Schwarzschild, BL, 128 RKF45 steps (gravitas-wasm/src/lib.rs:444-452 options)
        let mut o1 = IntegrationOptions::default();
        o1.max_steps = 128;
        v.push(integrate_case("config1_schwarzschild_bl_rkf45_128", 1.0, 0.0, false, &o1, "rkf45", 0.0, &fan));
        // config 4: Kerr 0.999 (as f32), KS, <= 1024 RKF45 steps
        let mut o4 = IntegrationOptions::default();
        o4.max_steps = 1024;
        v.push(integrate_case("config4_kerr_ks_rkf45_1024", 1.0, spin32, true, &o4, "rkf45", 0.0, &fan));
        // configs 2/3 use the per-step rule of compute.wgsl.ts:213, which gravitas-core does not have; what it has is
        // the constant-step implicit midpoint and RK4 (integrator.rs:193-226): pin those.
        for &(h, n) in &[(0.25, 512usize), (0.05, 256usize)] {
            let mut o = IntegrationOptions::default();
            o.method = IntegrationMethod::Symplectic { step_size: h };
            o.max_steps = n;
            v.push(integrate_case(&format!("kerr_ks_symplectic_h{}_{}", h, n), 1.0, spin32, true, &o, "symplectic", h, &fan));
            let mut o = IntegrationOptions::default();
            o.method = IntegrationMethod::RK4 { step_size: h };
            o.max_steps = n;
            v.push(integrate_case(&format!("kerr_ks_rk4_h{}_{}", h, n), 1.0, spin32, true, &o, "rk4", h, &fan));
        }
        sections.push(format!("\"integrate\":[{}]", v.join(",")));
    }

    // ---- g-factor, spectral LUT, disk LUT, Page-Thorne flux, Bardeen shadow ----
    {
        let mut v = Vec::new();
        for &(r, a, lam) in &[(6.0, 0.0, 0.0), (6.0, 0.0, 3.0), (2.0, spin32, 1.5), (10.0, spin32, -4.0), (30.0, 0.9, 5.0), (1000.0, 0.5, 0.0), (1.3, spin32, 2.0)] {
            v.push(format!("{{\"r\":{},\"mass\":1.0,\"spin\":{},\"lambda\":{},\"g\":{}}}", num(r), num(a), num(lam), num(redshift::kerr_g_factor(r, 1.0, a, lam))));
        }
        sections.push(format!("\"g_factor\":[{}]", v.join(",")));
        let lut = spectrum::generate_blackbody_lut(64, 16, 1e7);
        sections.push(format!("\"blackbody_lut\":{{\"width\":64,\"height\":16,\"max_temp\":1e7,\"rgba\":{}}}", arr32(&lut)));
        let bh = Kerr::new(1.0, spin32);
        let tl = disk::generate_temperature_lut(&bh, 512);
        sections.push(format!("\"temperature_lut\":{{\"mass\":1.0,\"spin\":{},\"width\":512,\"values\":{}}}", num(spin32), arr32(&tl)));
        let mut f = Vec::new();
        for &r in &[1.5, 2.0, 4.0, 6.0, 10.0, 25.0, 49.0] {
            f.push(format!("{{\"r\":{},\"flux\":{}}}", num(r), num(disk::page_thorne_flux(r, &bh, 1.0))));
        }
        sections.push(format!("\"page_thorne_flux\":{{\"mass\":1.0,\"spin\":{},\"m_dot\":1.0,\"samples\":[{}]}}", num(spin32), f.join(",")));
        let mut sh = Vec::new();
        for &(a, th) in &[(spin32, half_pi), (0.9, 1.0), (0.0, half_pi)] {
            let b = Kerr::new(1.0, a);
            let c = shadow::bardeen_shadow(&b, th, 32);
            let flat: Vec<f64> = c.iter().flat_map(|p| vec![p.0, p.1]).collect();
            sh.push(format!("{{\"spin\":{},\"theta_obs\":{},\"n_points\":32,\"alpha_beta\":{}}}", num(a), num(th), arr(&flat)));
        }
        sections.push(format!("\"bardeen_shadow\":[{}]", sh.join(",")));
    }

    println!("{{\"schema\":1,\"generator\":\"oracle/ref_fixtures/gen_fixtures.rs\",\"crate\":\"gravitas-core 0.1.0\",{}}}", sections.join(","));
}
