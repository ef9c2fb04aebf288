// gravitas_oracle.hpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// A C++17 restatement of the reference's Rust `gravitas-core` f64 geodesic path, kept
// operation-for-operation in the order the Rust source writes them so that, compiled with
// `-ffp-contract=off -fno-fast-math`, it produces what the Rust would (up to libm differences in
// sin/cos/pow/exp/acos). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may build, load or call anything in oracle/. The product
// (blackhole-simulation_b200/) never includes this header.
//
// PARITY STATUS: "parity unpinned" for integrate / adaptive_rkf45_step / renormalize_null /
// generate_blackbody_lut texels / camera->state / LUT sampling / final RGBA — the reference holds no
// golden vector or test for any of those values and cannot be built here (no
// cargo/rustc/node). What IS pinned, and checked in tests/test_oracle_kat.py: the reference's doctest
// and unit-test constants (metric/kerr.rs:31-33, 507-597), the legacy integrator assertions
// (_legacy_src/integrator.rs:102-150, 352-441) and SURVEY §8c's provisional known-answer table (an
// independent Python restatement of the same Rust).
//
// Every function cites the reference file:line (relative to /root/reference/) that it follows.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>
#include <algorithm>

namespace orc {

// ---------------------------------------------------------------------------------------------
// Counted<double>: instrumented scalar for the algorithmic-flop census.
// add/sub/mul/div/sqrt = 1 flop each; trig / pow / exp are counted separately as "sfu".
// ---------------------------------------------------------------------------------------------
struct FlopCensus {
    uint64_t add = 0, mul = 0, div = 0, sqrt_ = 0, trig = 0, pow_ = 0, cmp = 0;
    uint64_t flops() const { return add + mul + div + sqrt_; }
};
inline FlopCensus& census() { static thread_local FlopCensus c; return c; }

struct Counted {
    double v;
    Counted() : v(0.0) {}
    Counted(double x) : v(x) {}
    explicit operator double() const { return v; }
};
inline Counted operator+(Counted a, Counted b) { census().add++; return Counted(a.v + b.v); }
inline Counted operator-(Counted a, Counted b) { census().add++; return Counted(a.v - b.v); }
inline Counted operator*(Counted a, Counted b) { census().mul++; return Counted(a.v * b.v); }
inline Counted operator/(Counted a, Counted b) { census().div++; return Counted(a.v / b.v); }
inline Counted operator-(Counted a) { return Counted(-a.v); }
inline Counted& operator+=(Counted& a, Counted b) { a = a + b; return a; }
inline Counted& operator-=(Counted& a, Counted b) { a = a - b; return a; }
inline Counted& operator*=(Counted& a, Counted b) { a = a * b; return a; }
inline bool operator<(Counted a, Counted b) { census().cmp++; return a.v < b.v; }
inline bool operator>(Counted a, Counted b) { census().cmp++; return a.v > b.v; }
inline bool operator<=(Counted a, Counted b) { census().cmp++; return a.v <= b.v; }
inline bool operator>=(Counted a, Counted b) { census().cmp++; return a.v >= b.v; }
inline bool operator==(Counted a, Counted b) { census().cmp++; return a.v == b.v; }
inline bool operator!=(Counted a, Counted b) { census().cmp++; return a.v != b.v; }
inline Counted sin(Counted a) { census().trig++; return Counted(std::sin(a.v)); }
inline Counted cos(Counted a) { census().trig++; return Counted(std::cos(a.v)); }
inline Counted acos(Counted a) { census().trig++; return Counted(std::acos(a.v)); }
inline Counted sqrt(Counted a) { census().sqrt_++; return Counted(std::sqrt(a.v)); }
inline Counted fabs(Counted a) { return Counted(std::fabs(a.v)); }
inline Counted pow(Counted a, Counted b) { census().pow_++; return Counted(std::pow(a.v, b.v)); }
inline Counted exp(Counted a) { census().pow_++; return Counted(std::exp(a.v)); }
inline Counted floor(Counted a) { return Counted(std::floor(a.v)); }
inline double to_double(Counted a) { return a.v; }
inline double to_double(double a) { return a; }
inline double to_double(float a) { return (double)a; }
inline double to_double(long double a) { return (double)a; }

using std::sin; using std::cos; using std::acos; using std::sqrt; using std::fabs; using std::pow;
using std::exp; using std::floor;

// Rust f64::max / f64::min / clamp / signum semantics for the (non-NaN) values on this path.
template <class R> inline R rmax(R a, R b) { return (a < b) ? b : a; }
template <class R> inline R rmin(R a, R b) { return (b < a) ? b : a; }
template <class R> inline R rclamp(R x, R lo, R hi) { R y = x; if (y < lo) y = lo; if (y > hi) y = hi; return y; }
template <class R> inline R rsignum(R x) { return (x < R(0.0)) ? R(-1.0) : R(1.0); }  // f64::signum(+0)=1

enum Coords : int { BOYER_LINDQUIST = 0, KERR_SCHILD = 1 };
// geodesic/termination.rs:4-17 (#[repr(C)] enum)
enum Termination : uint32_t { TERM_NONE = 0, TERM_HORIZON = 1, TERM_ESCAPE = 2, TERM_MAXSTEPS = 3, TERM_DISK = 4 };
enum Method : int { METHOD_RKF45 = 0, METHOD_RK4 = 1, METHOD_SYMPLECTIC = 2, METHOD_VERLET_GLSL = 3 };

// ---------------------------------------------------------------------------------------------
// metric/kerr.rs — Kerr metric (mass, clamped spin, coordinate system)
// ---------------------------------------------------------------------------------------------
template <class R>
struct Kerr {
    R m;      // mass_val
    R spin;   // spin_val, clamped to [-1,1] (kerr.rs:49-55,58-64)
    int coords;
    Kerr(R mass, R s, int c) : m(mass), spin(rclamp<R>(s, R(-1.0), R(1.0))), coords(c) {}
    R a() const { return spin * m; }  // kerr.rs:72-75

    // metric/mod.rs:75-84
    R event_horizon() const {
        R aa = spin * m;
        R disc = m * m - aa * aa;
        if (disc < R(0.0)) return m;
        return m + sqrt(disc);
    }
    // kerr.rs:91-94
    R photon_sphere() const {
        R term = R(2.0 / 3.0) * acos(-spin);
        return R(2.0) * m * (R(1.0) + cos(term));
    }
    // kerr.rs:100-123 (sign: prograde -1, retrograde +1)
    R isco(bool prograde) const {
        R a_star = spin;
        if (fabs(a_star) < R(1e-6)) return m * R(6.0);
        R a2 = a_star * a_star;
        R z1 = R(1.0) + pow(R(1.0) - a2, R(1.0 / 3.0)) *
                            (pow(R(1.0) + a_star, R(1.0 / 3.0)) + pow(R(1.0) - a_star, R(1.0 / 3.0)));
        R z2 = sqrt(R(3.0) * a2 + z1 * z1);
        R sign = prograde ? R(-1.0) : R(1.0);
        R disc = (R(3.0) - z1) * (R(3.0) + z1 + R(2.0) * z2);
        R root = (disc < R(0.0)) ? R(0.0) : sqrt(disc);
        return m * (R(3.0) + z2 + sign * root);
    }

    // kerr.rs:242-264 covariant_bl — only the entries time_dilation needs (g_tt).
    R covariant_bl_tt(R r, R theta) const {
        R aa = a();
        R r2 = r * r, a2 = aa * aa;
        R cos_theta = cos(theta);
        R cos2 = cos_theta * cos_theta;
        R sigma = r2 + a2 * cos2;
        return -(R(1.0) - (R(2.0) * m * r) / sigma);
    }
    // kerr.rs:180-188 time_dilation
    R time_dilation(R r, R theta) const {
        R g_tt = covariant_bl_tt(r, theta);
        if (g_tt >= R(0.0)) return R(0.0);
        return sqrt(-g_tt);
    }

    // kerr.rs:266-293 contravariant_bl; kerr.rs:412-440 contravariant_ks. g is row-major [16].
    void contravariant(R r, R theta, R g[16]) const {
        for (int i = 0; i < 16; i++) g[i] = R(0.0);
        R aa = a();
        R r2 = r * r, a2 = aa * aa;
        if (coords == BOYER_LINDQUIST) {
            R sin_theta = sin(theta);
            R cos_theta = cos(theta);
            R sin2 = sin_theta * sin_theta;
            R cos2 = cos_theta * cos_theta;
            R sigma = r2 + a2 * cos2;
            R delta = r2 - R(2.0) * m * r + a2;
            R g_tt = -((sigma * (r2 + a2) + R(2.0) * m * r * a2 * sin2) / (delta * sigma));
            R g_rr = delta / sigma;
            R g_thth = R(1.0) / sigma;
            R g_phph = (sin2 < R(1e-9)) ? R(0.0) : (delta - a2 * sin2) / (delta * sigma * sin2);
            R g_tph = -(R(2.0) * m * r * aa) / (delta * sigma);
            g[0] = g_tt; g[3] = g_tph; g[5] = g_rr; g[10] = g_thth; g[12] = g_tph; g[15] = g_phph;
        } else {
            R s = sin(theta);
            R sin2 = rmax<R>(s * s, R(1e-12));
            R cos2 = R(1.0) - sin2;
            R sigma = r2 + a2 * cos2;
            R delta = r2 - R(2.0) * m * r + a2;
            R g_tt = -(R(1.0) + R(2.0) * m * r / sigma);
            R g_tr = R(2.0) * m * r / sigma;
            R g_rr = delta / sigma;
            R g_thth = R(1.0) / sigma;
            R g_phph = R(1.0) / (sigma * sin2);
            R g_rph = aa / sigma;
            g[0] = g_tt; g[1] = g_tr; g[4] = g_tr; g[5] = g_rr; g[7] = g_rph; g[10] = g_thth;
            g[13] = g_rph; g[15] = g_phph;
        }
    }

    // kerr.rs:295-372 hamiltonian_derivs_bl; kerr.rs:442-499 hamiltonian_derivs_ks
    void hamiltonian_derivatives(R r, R theta, const R p[4], R& dh_dr, R& dh_dtheta) const {
        R aa = a();
        R r2 = r * r, a2 = aa * aa;
        if (coords == BOYER_LINDQUIST) {
            R cos_theta = cos(theta);
            R sin_theta = sin(theta);
            R sin2 = sin_theta * sin_theta;
            R cos2 = cos_theta * cos_theta;
            R sigma = r2 + a2 * cos2;
            R delta = r2 - R(2.0) * m * r + a2;
            R sigma_sq = sigma * sigma;
            R dsigma_dr = R(2.0) * r;
            R dsigma_dtheta = R(-2.0) * a2 * cos_theta * sin_theta;
            R ddelta_dr = R(2.0) * r - R(2.0) * m;

            R dg_rr_dr = (ddelta_dr * sigma - delta * dsigma_dr) / sigma_sq;
            R dg_rr_dtheta = -(delta * dsigma_dtheta) / sigma_sq;
            R dg_thth_dr = -dsigma_dr / sigma_sq;
            R dg_thth_dtheta = -dsigma_dtheta / sigma_sq;

            R num_tphi = R(-2.0) * m * r * aa;
            R den_tphi = delta * sigma;
            R dnum_tphi_dr = R(-2.0) * m * aa;
            R dden_tphi_dr = ddelta_dr * sigma + delta * dsigma_dr;
            R dg_tphi_dr = (dnum_tphi_dr * den_tphi - num_tphi * dden_tphi_dr) / (den_tphi * den_tphi);
            R dden_tphi_dtheta = delta * dsigma_dtheta;
            R dg_tphi_dtheta = -(num_tphi * dden_tphi_dtheta) / (den_tphi * den_tphi);

            R du_dr = dsigma_dr * (r2 + a2) + sigma * R(2.0) * r + R(2.0) * m * a2 * sin2;
            R dv_dr = dden_tphi_dr;
            R u_val = sigma * (r2 + a2) + R(2.0) * m * r * a2 * sin2;
            R dg_tt_dr = -(du_dr * den_tphi - u_val * dv_dr) / (den_tphi * den_tphi);

            R du_dtheta = dsigma_dtheta * (r2 + a2) + R(2.0) * m * r * a2 * R(2.0) * sin_theta * cos_theta;
            R dv_dtheta = dden_tphi_dtheta;
            R dg_tt_dtheta = -(du_dtheta * den_tphi - u_val * dv_dtheta) / (den_tphi * den_tphi);

            R da_dr = -dsigma_dr / (sigma_sq * sin2);
            R db_dr = -a2 * dden_tphi_dr / (den_tphi * den_tphi);
            R dg_phph_dr = da_dr - db_dr;

            R d_denom_a_dtheta = dsigma_dtheta * sin2 + sigma * R(2.0) * sin_theta * cos_theta;
            R da_dtheta = -d_denom_a_dtheta / (sigma_sq * sin2 * sin2);
            R db_dtheta = -a2 * dden_tphi_dtheta / (den_tphi * den_tphi);
            R dg_phph_dtheta = da_dtheta - db_dtheta;

            R p_t = p[0], p_r = p[1], p_th = p[2], p_ph = p[3];
            dh_dr = R(0.5) * (p_t * p_t * dg_tt_dr + p_r * p_r * dg_rr_dr + p_th * p_th * dg_thth_dr +
                              p_ph * p_ph * dg_phph_dr + R(2.0) * p_t * p_ph * dg_tphi_dr);
            dh_dtheta = R(0.5) * (p_t * p_t * dg_tt_dtheta + p_r * p_r * dg_rr_dtheta +
                                  p_th * p_th * dg_thth_dtheta + p_ph * p_ph * dg_phph_dtheta +
                                  R(2.0) * p_t * p_ph * dg_tphi_dtheta);
        } else {
            R sin_theta = sin(theta);
            R cos_theta = cos(theta);
            R sin2 = rmax<R>(sin_theta * sin_theta, R(1e-12));
            R cos2 = R(1.0) - sin2;
            R sigma = r2 + a2 * cos2;
            R sigma2 = sigma * sigma;
            R delta = r2 - R(2.0) * m * r + a2;

            R dsigma_dr = R(2.0) * r;
            R dsigma_dtheta = R(-2.0) * a2 * sin_theta * cos_theta;
            R ddelta_dr = R(2.0) * r - R(2.0) * m;

            R dg_tt_dr = -(R(2.0) * m * (sigma - r * dsigma_dr)) / sigma2;
            R dg_tt_dtheta = (R(2.0) * m * r * dsigma_dtheta) / sigma2;
            R dg_tr_dr = -dg_tt_dr;
            R dg_tr_dtheta = -dg_tt_dtheta;
            R dg_rr_dr = (ddelta_dr * sigma - delta * dsigma_dr) / sigma2;
            R dg_rr_dtheta = -(delta * dsigma_dtheta) / sigma2;
            R dg_thth_dr = -dsigma_dr / sigma2;
            R dg_thth_dtheta = -dsigma_dtheta / sigma2;
            R dg_phph_dr = -dsigma_dr / (sigma2 * sin2);
            R dg_phph_dtheta =
                -(dsigma_dtheta * sin2 + sigma * R(2.0) * sin_theta * cos_theta) / (sigma2 * sin2 * sin2);
            R dg_rph_dr = -(aa * dsigma_dr) / sigma2;
            R dg_rph_dtheta = -(aa * dsigma_dtheta) / sigma2;

            dh_dr = R(0.5) * (dg_tt_dr * p[0] * p[0] + dg_rr_dr * p[1] * p[1] + dg_thth_dr * p[2] * p[2] +
                              dg_phph_dr * p[3] * p[3] + R(2.0) * dg_tr_dr * p[0] * p[1] +
                              R(2.0) * dg_rph_dr * p[1] * p[3]);
            dh_dtheta = R(0.5) * (dg_tt_dtheta * p[0] * p[0] + dg_rr_dtheta * p[1] * p[1] +
                                  dg_thth_dtheta * p[2] * p[2] + dg_phph_dtheta * p[3] * p[3] +
                                  R(2.0) * dg_tr_dtheta * p[0] * p[1] + R(2.0) * dg_rph_dtheta * p[1] * p[3]);
            if (fabs(sin_theta) < R(1e-10)) dh_dtheta = R(0.0);
        }
    }
};

// ---------------------------------------------------------------------------------------------
// geodesic/mod.rs:23-30 GeodesicState (#[repr(C)] 2x[f64;4]) and :72-146 Butcher-combination helpers
// ---------------------------------------------------------------------------------------------
template <class R>
struct State {
    R x[4];
    R p[4];
};

template <class R>
inline State<R> add_scaled(const State<R>& s, const State<R>& k, R c) {  // mod.rs:73-80
    State<R> n = s;
    for (int i = 0; i < 4; i++) { n.x[i] += k.x[i] * c; n.p[i] += k.p[i] * c; }
    return n;
}
template <class R>
inline State<R> add_scaled_2(const State<R>& s, const State<R>& k1, R s1, const State<R>& k2, R s2) {  // :82-89
    State<R> n = s;
    for (int i = 0; i < 4; i++) {
        n.x[i] += k1.x[i] * s1 + k2.x[i] * s2;
        n.p[i] += k1.p[i] * s1 + k2.p[i] * s2;
    }
    return n;
}
template <class R>
inline State<R> add_scaled_3(const State<R>& s, const State<R>& k1, R s1, const State<R>& k2, R s2,
                             const State<R>& k3, R s3) {  // :91-106
    State<R> n = s;
    for (int i = 0; i < 4; i++) {
        n.x[i] += k1.x[i] * s1 + k2.x[i] * s2 + k3.x[i] * s3;
        n.p[i] += k1.p[i] * s1 + k2.p[i] * s2 + k3.p[i] * s3;
    }
    return n;
}
template <class R>
inline State<R> add_scaled_4(const State<R>& s, const State<R>& k1, R s1, const State<R>& k2, R s2,
                             const State<R>& k3, R s3, const State<R>& k4, R s4) {  // :108-125
    State<R> n = s;
    for (int i = 0; i < 4; i++) {
        n.x[i] += k1.x[i] * s1 + k2.x[i] * s2 + k3.x[i] * s3 + k4.x[i] * s4;
        n.p[i] += k1.p[i] * s1 + k2.p[i] * s2 + k3.p[i] * s3 + k4.p[i] * s4;
    }
    return n;
}
template <class R>
inline State<R> add_scaled_5(const State<R>& s, const State<R>& k1, R s1, const State<R>& k2, R s2,
                             const State<R>& k3, R s3, const State<R>& k4, R s4, const State<R>& k5,
                             R s5) {  // :127-146
    State<R> n = s;
    for (int i = 0; i < 4; i++) {
        n.x[i] += k1.x[i] * s1 + k2.x[i] * s2 + k3.x[i] * s3 + k4.x[i] * s4 + k5.x[i] * s5;
        n.p[i] += k1.p[i] * s1 + k2.p[i] * s2 + k3.p[i] * s3 + k4.p[i] * s4 + k5.p[i] * s5;
    }
    return n;
}

// geodesic/hamiltonian.rs:13-35 get_state_derivative
template <class R>
inline State<R> state_derivative(const State<R>& s, const Kerr<R>& metric) {
    R r = s.x[1], theta = s.x[2];
    R g[16];
    metric.contravariant(r, theta, g);
    const R* p = s.p;
    R dt = g[0] * p[0] + g[1] * p[1] + g[3] * p[3];
    R dr = g[4] * p[0] + g[5] * p[1] + g[7] * p[3];
    R dth = g[10] * p[2];
    R dph = g[12] * p[0] + g[13] * p[1] + g[15] * p[3];
    R dh_dr, dh_dtheta;
    metric.hamiltonian_derivatives(r, theta, s.p, dh_dr, dh_dtheta);
    State<R> d;
    d.x[0] = dt; d.x[1] = dr; d.x[2] = dth; d.x[3] = dph;
    d.p[0] = R(0.0); d.p[1] = -dh_dr; d.p[2] = -dh_dtheta; d.p[3] = R(0.0);
    return d;
}

// invariants/mod.rs:25-37 hamiltonian
template <class R>
inline R hamiltonian(const State<R>& s, const Kerr<R>& metric) {
    R g[16];
    metric.contravariant(s.x[1], s.x[2], g);
    const R* p = s.p;
    return R(0.5) * (g[0] * p[0] * p[0] + g[5] * p[1] * p[1] + g[10] * p[2] * p[2] + g[15] * p[3] * p[3] +
                     R(2.0) * g[3] * p[0] * p[3] + R(2.0) * g[1] * p[0] * p[1] + R(2.0) * g[7] * p[1] * p[3]);
}

// invariants/renormalization.rs:13-45 renormalize_null
template <class R>
inline void renormalize_null(State<R>& s, const Kerr<R>& metric) {
    R g[16];
    metric.contravariant(s.x[1], s.x[2], g);
    R p_t = s.p[0], p_r = s.p[1], p_th = s.p[2], p_ph = s.p[3];
    R a_quad = g[5];
    R b_quad = R(2.0) * (g[1] * p_t + g[7] * p_ph);
    R c_quad = g[0] * p_t * p_t + g[10] * p_th * p_th + g[15] * p_ph * p_ph + R(2.0) * g[3] * p_t * p_ph;
    if (fabs(a_quad) > R(1e-12)) {
        R discriminant = b_quad * b_quad - R(4.0) * a_quad * c_quad;
        if (discriminant >= R(0.0)) {
            R sqrt_d = sqrt(discriminant);
            R sol1 = (-b_quad + sqrt_d) / (R(2.0) * a_quad);
            R sol2 = (-b_quad - sqrt_d) / (R(2.0) * a_quad);
            s.p[1] = (fabs(sol1 - p_r) < fabs(sol2 - p_r)) ? sol1 : sol2;
        }
    }
}

// geodesic/integrator.rs:113-190 adaptive_rkf45_step (Fehlberg 4(5); returns 5th-order state; error =
// max over the 4 POSITION components of |h * sum (b5-b4)_j k_j.x[i]|, absolute).
template <class R>
inline State<R> rkf45_step(const State<R>& s, const Kerr<R>& metric, R h, R& error_out, uint64_t* rhs_evals = nullptr) {
    State<R> k1 = state_derivative(s, metric);
    State<R> k2 = state_derivative(add_scaled(s, k1, h / R(4.0)), metric);
    State<R> k3 = state_derivative(add_scaled_2(s, k1, R(3.0) * h / R(32.0), k2, R(9.0) * h / R(32.0)), metric);
    State<R> k4 = state_derivative(
        add_scaled_3(s, k1, R(1932.0) * h / R(2197.0), k2, R(-7200.0) * h / R(2197.0), k3, R(7296.0) * h / R(2197.0)),
        metric);
    State<R> k5 = state_derivative(
        add_scaled_4(s, k1, R(439.0) * h / R(216.0), k2, R(-8.0) * h, k3, R(3680.0) * h / R(513.0), k4,
                     R(-845.0) * h / R(4104.0)),
        metric);
    State<R> k6 = state_derivative(
        add_scaled_5(s, k1, R(-8.0) * h / R(27.0), k2, R(2.0) * h, k3, R(-3544.0) * h / R(2565.0), k4,
                     R(1859.0) * h / R(4104.0), k5, R(-11.0) * h / R(40.0)),
        metric);
    if (rhs_evals) *rhs_evals += 6;

    // The Rust constant-folds `16.0 / 135.0` etc. at compile time (f64 IEEE division) — same values here.
    const R b1(16.0 / 135.0), b3(6656.0 / 12825.0), b4(28561.0 / 56430.0), b5(9.0 / 50.0), b6(2.0 / 55.0);
    State<R> f = s;
    for (int i = 0; i < 4; i++) {
        f.x[i] += h * (b1 * k1.x[i] + b3 * k3.x[i] + b4 * k4.x[i] - b5 * k5.x[i] + b6 * k6.x[i]);
        f.p[i] += h * (b1 * k1.p[i] + b3 * k3.p[i] + b4 * k4.p[i] - b5 * k5.p[i] + b6 * k6.p[i]);
    }
    const R e1(16.0 / 135.0 - 25.0 / 216.0), e3(6656.0 / 12825.0 - 1408.0 / 2565.0),
        e4(28561.0 / 56430.0 - 2197.0 / 4104.0), e5(-9.0 / 50.0 + 1.0 / 5.0), e6(2.0 / 55.0);
    R error(0.0);
    for (int i = 0; i < 4; i++) {
        R err = h * (e1 * k1.x[i] + e3 * k3.x[i] + e4 * k4.x[i] + e5 * k5.x[i] + e6 * k6.x[i]);
        error = rmax<R>(error, fabs(err));
    }
    error_out = error;
    return f;
}

// geodesic/integrator.rs:53-108 AdaptiveStepper
template <class R>
struct AdaptiveStepper {
    R safety_factor, min_step, max_step, tolerance;
    uint64_t attempts = 0, rejects = 0, rhs_evals = 0;
    explicit AdaptiveStepper(R tol) : safety_factor(0.9), min_step(1e-5), max_step(10.0), tolerance(tol) {}

    // Updates `s` in place; returns the recommended next step (integrator.rs:72-107).
    R step(State<R>& s, const Kerr<R>& metric, R h_try) {
        R h = rclamp<R>(h_try, -max_step, max_step);
        for (;;) {
            R error_estimate;
            State<R> ns = rkf45_step(s, metric, h, error_estimate, &rhs_evals);
            attempts++;
            R error_ratio = (error_estimate == R(0.0)) ? R(0.0) : error_estimate / tolerance;
            if (error_ratio <= R(1.0)) {
                s = ns;
                R growth = (error_ratio < R(1e-4)) ? R(5.0) : safety_factor * pow(error_ratio, R(-0.2));
                R next_h = h * rmin<R>(growth, R(5.0));
                return rclamp<R>(next_h, -max_step, max_step);
            } else {
                rejects++;
                R shrink = safety_factor * pow(error_ratio, R(-0.25));
                h *= rmax<R>(shrink, R(0.1));
                if (fabs(h) < min_step) {
                    R e2;
                    State<R> forced = rkf45_step(s, metric, min_step * rsignum(h), e2, &rhs_evals);
                    attempts++;
                    s = forced;
                    return min_step * rsignum(h);
                }
            }
        }
    }
};

// geodesic/integrator.rs:193-203 step_rk4
template <class R>
inline void step_rk4(State<R>& s, const Kerr<R>& metric, R h) {
    State<R> k1 = state_derivative(s, metric);
    State<R> k2 = state_derivative(add_scaled(s, k1, R(0.5) * h), metric);
    State<R> k3 = state_derivative(add_scaled(s, k2, R(0.5) * h), metric);
    State<R> k4 = state_derivative(add_scaled(s, k3, h), metric);
    for (int i = 0; i < 4; i++) {
        s.x[i] += (h / R(6.0)) * (k1.x[i] + R(2.0) * k2.x[i] + R(2.0) * k3.x[i] + k4.x[i]);
        s.p[i] += (h / R(6.0)) * (k1.p[i] + R(2.0) * k2.p[i] + R(2.0) * k3.p[i] + k4.p[i]);
    }
}

// geodesic/integrator.rs:209-226 step_symplectic (implicit midpoint, 2 fixed-point iterations + final)
template <class R>
inline void step_symplectic(State<R>& s, const Kerr<R>& metric, R h) {
    State<R> s_mid = s;
    for (int it = 0; it < 2; it++) {
        State<R> d = state_derivative(s_mid, metric);
        State<R> s_next = s;
        for (int i = 0; i < 4; i++) {
            s_next.x[i] = s.x[i] + d.x[i] * h;
            s_next.p[i] = s.p[i] + d.p[i] * h;
            s_mid.x[i] = R(0.5) * (s.x[i] + s_next.x[i]);
            s_mid.p[i] = R(0.5) * (s.p[i] + s_next.p[i]);
        }
    }
    State<R> d_final = state_derivative(s_mid, metric);
    for (int i = 0; i < 4; i++) {
        s.x[i] += d_final.x[i] * h;
        s.p[i] += d_final.p[i] * h;
    }
}

// geodesic/integrator.rs:24-47 IntegrationOptions (+ the per-step h rule of shaders/compute.wgsl.ts:213
// that the fixed-step configs use: h = clamp(0.15 (r - r+), 0.05, 1.0); step_rule=0 keeps `step_size`).
struct Options {
    int method = METHOD_RKF45;
    double tolerance = 1e-8;
    double initial_step = 0.01;   // also the fixed step_size for RK4/Symplectic when step_rule==0
    uint64_t max_steps = 10000;
    double escape_radius = 1000.0;
    uint64_t renormalize_interval = 10;
    int step_rule = 0;            // 0: constant step_size (Rust); 1: compute.wgsl.ts:213 rule
};

template <class R>
struct Trajectory {  // geodesic/mod.rs:150-161
    State<R> final_state;
    uint32_t termination;
    uint64_t steps_taken;
    R max_hamiltonian_drift;
    uint64_t attempts, rejects, rhs_evals;
};

// geodesic/mod.rs:255-265 check_termination
template <class R>
inline uint32_t check_termination(const State<R>& s, R horizon, R escape_r) {
    R r = s.x[1];
    if (r < horizon * R(1.001)) return TERM_HORIZON;
    if (r > escape_r) return TERM_ESCAPE;
    return TERM_NONE;
}

// compute.wgsl.ts:213
template <class R>
inline R wgsl_step_rule(R r, R rh) { return rclamp<R>((r - rh) * R(0.15), R(0.05), R(1.0)); }

// Per-step observer hook used by the composite RGBA oracle (disk crossings). Return true to stop.
template <class R>
struct NoHook { bool operator()(const State<R>&, const State<R>&) { return false; } };

// geodesic/mod.rs:180-253 integrate
template <class R, class Hook = NoHook<R>>
inline Trajectory<R> integrate(const State<R>& initial, const Kerr<R>& metric, const Options& o, Hook hook = Hook()) {
    State<R> s = initial;
    AdaptiveStepper<R> stepper{R(o.tolerance)};
    R h(o.initial_step);
    R horizon = metric.event_horizon();
    R max_drift(0.0);
    uint64_t steps = 0;
    renormalize_null(s, metric);
    Trajectory<R> t;
    for (uint64_t it = 0; it < o.max_steps; it++) {
        uint32_t term = check_termination(s, horizon, R(o.escape_radius));
        if (term != TERM_NONE) {
            t.final_state = s; t.termination = term; t.steps_taken = steps; t.max_hamiltonian_drift = max_drift;
            t.attempts = stepper.attempts; t.rejects = stepper.rejects; t.rhs_evals = stepper.rhs_evals;
            return t;
        }
        State<R> prev = s;
        if (o.method == METHOD_RKF45) {
            h = stepper.step(s, metric, h);
        } else {
            R hs = (o.step_rule == 1) ? wgsl_step_rule<R>(s.x[1], horizon) : R(o.initial_step);
            if (o.method == METHOD_RK4) { step_rk4(s, metric, hs); stepper.rhs_evals += 4; }
            else { step_symplectic(s, metric, hs); stepper.rhs_evals += 3; }
            stepper.attempts++;
        }
        if (steps % o.renormalize_interval == 0) renormalize_null(s, metric);
        R h_val = fabs(hamiltonian(s, metric));
        if (h_val > max_drift) max_drift = h_val;
        steps += 1;
        if (hook(prev, s)) {
            t.final_state = s; t.termination = TERM_DISK; t.steps_taken = steps; t.max_hamiltonian_drift = max_drift;
            t.attempts = stepper.attempts; t.rejects = stepper.rejects; t.rhs_evals = stepper.rhs_evals;
            return t;
        }
    }
    t.final_state = s; t.termination = TERM_MAXSTEPS; t.steps_taken = steps; t.max_hamiltonian_drift = max_drift;
    t.attempts = stepper.attempts; t.rejects = stepper.rejects; t.rhs_evals = stepper.rhs_evals;
    return t;
}

// ---------------------------------------------------------------------------------------------
// physics/redshift.rs:65-95 kerr_g_factor
// ---------------------------------------------------------------------------------------------
template <class R>
inline R kerr_g_factor(R r, R mass, R spin, R lambda) {
    R a = spin * mass;
    R r2 = r * r;
    R a2 = a * a;
    R m = mass;
    R omega = sqrt(m) / (pow(r, R(1.5)) + a * sqrt(m));
    R sigma = r2;
    R g_tt = -(R(1.0) - R(2.0) * m * r / sigma);
    R g_tphi = -(R(2.0) * m * r * a) / sigma;
    R g_phiphi = r2 + a2 + R(2.0) * m * r * a2 / sigma;
    R ut_denom = -g_tt - R(2.0) * omega * g_tphi - omega * omega * g_phiphi;
    if (ut_denom <= R(0.0)) return R(0.0);
    R ut = R(1.0) / sqrt(ut_denom);
    R factor = R(1.0) - lambda * omega;
    if (fabs(factor) < R(1e-30)) return R(0.0);
    return R(1.0) / (ut * factor);
}

// ---------------------------------------------------------------------------------------------
// physics/spectrum.rs — Planck law, CIE fit, XYZ->linear sRGB, 2-D blackbody/redshift LUT (f64 only)
// ---------------------------------------------------------------------------------------------
namespace spectrum {
constexpr double SI_C = 299792458.0;       // constants.rs:22
constexpr double SI_KB = 1.380649e-23;     // constants.rs:31
constexpr double H = 6.62607015e-34;       // spectrum.rs:5
constexpr double C1 = 2.0 * H * SI_C * SI_C;   // spectrum.rs:6
constexpr double C2 = H * SI_C / SI_KB;        // spectrum.rs:7

inline double powi5(double x) { return x * x * x * x * x; }  // f64::powi(5) = repeated multiplication

inline double planck_law(double lambda, double temperature) {  // spectrum.rs:12-18
    double exponent = C2 / (lambda * temperature);
    if (exponent > 100.0) return 0.0;
    return (C1 / powi5(lambda)) / (std::exp(exponent) - 1.0);
}
inline void cie_1931(double lambda, double& cx, double& cy, double& cz) {  // spectrum.rs:50-63
    double l_nm = lambda * 1e9;
    auto g = [l_nm](double mean, double sd) { double x = (l_nm - mean) / sd; return std::exp(-0.5 * x * x); };
    double x = 1.056 * g(599.0, 37.9) + 0.362 * g(442.0, 16.0) - 0.065 * g(501.0, 20.4);
    double y = 0.821 * g(568.0, 46.9) + 0.286 * g(530.0, 22.1);
    double z = 1.217 * g(437.0, 11.8) + 0.681 * g(459.0, 26.0);
    cx = std::max(x, 0.0); cy = std::max(y, 0.0); cz = std::max(z, 0.0);
}
inline void integrate_planck_xyz(double temperature, double xyz[3]) {  // spectrum.rs:23-46
    xyz[0] = xyz[1] = xyz[2] = 0.0;
    if (temperature < 100.0) return;
    double x = 0.0, y = 0.0, z = 0.0;
    double lambda = 380.0e-9;
    const double end = 780.0e-9, step = 2.0e-9;
    while (lambda <= end) {
        double intensity = planck_law(lambda, temperature);
        double cx, cy, cz;
        cie_1931(lambda, cx, cy, cz);
        x += intensity * cx * step;
        y += intensity * cy * step;
        z += intensity * cz * step;
        lambda += step;
    }
    xyz[0] = x; xyz[1] = y; xyz[2] = z;
}
inline void xyz_to_linear_rgb(double x, double y, double z, float rgb[3]) {  // spectrum.rs:66-71
    double r = 3.2404542 * x - 1.5371385 * y - 0.4985314 * z;
    double g = -0.9692660 * x + 1.8760108 * y + 0.0415560 * z;
    double b = 0.0556434 * x - 0.2040259 * y + 1.0572252 * z;
    rgb[0] = (float)std::max(r, 0.0); rgb[1] = (float)std::max(g, 0.0); rgb[2] = (float)std::max(b, 0.0);
}
// spectrum.rs:76-102 generate_blackbody_lut -> width*height*4 f32 (RGBA, alpha 1)
inline void generate_blackbody_lut(size_t width, size_t height, double max_temp, float* out) {
    const double min_g = 0.05, max_g = 5.0;
    for (size_t y = 0; y < height; y++) {
        double g = min_g + (max_g - min_g) * ((double)y / (double)std::max<size_t>(height - 1, 1));
        for (size_t x = 0; x < width; x++) {
            double t = std::pow((double)x / (double)std::max<size_t>(width - 1, 1), 2.5) * max_temp;
            double t_eff = t * g;
            double xyz[3];
            integrate_planck_xyz(t_eff, xyz);
            float rgb[3];
            xyz_to_linear_rgb(xyz[0], xyz[1], xyz[2], rgb);
            double g4 = g * g * g * g;  // powi(4)
            double scale = 1.0e-14 * g4;
            float* px = out + 4 * (y * width + x);
            px[0] = rgb[0] * (float)scale;   // `rgb[0] * scale as f32`: cast binds tighter than `*`
            px[1] = rgb[1] * (float)scale;
            px[2] = rgb[2] * (float)scale;
            px[3] = 1.0f;
        }
    }
}
}  // namespace spectrum

// ---------------------------------------------------------------------------------------------
// physics/disk.rs — Page-Thorne flux and the 1-D disk temperature LUT (f64 only)
// ---------------------------------------------------------------------------------------------
namespace disk {
inline double specific_energy(double r, double m, double a) {  // disk.rs:24-36
    double rm = r / m;
    double sqrt_mr = std::sqrt(m / r);
    double am = a / m;
    double num = 1.0 - 2.0 / rm + am * sqrt_mr;
    double den_sq = 1.0 - 3.0 / rm + 2.0 * am * sqrt_mr;
    if (den_sq <= 0.0) return 1.0;
    return num / std::sqrt(den_sq);
}
inline double specific_angular_momentum(double r, double m, double a) {  // disk.rs:44-57
    double rm = r / m;
    double sqrt_mr = std::sqrt(m / r);
    double am = a / m;
    double ar = a / r;
    double a2r2 = ar * ar;
    double num = std::sqrt(m) * std::sqrt(r) * (1.0 - 2.0 * am * sqrt_mr + a2r2);
    double den_sq = 1.0 - 3.0 / rm + 2.0 * am * sqrt_mr;
    if (den_sq <= 0.0) return 0.0;
    return num / std::sqrt(den_sq);
}
inline double angular_velocity(double r, double m, double a) {  // disk.rs:62-64
    return std::sqrt(m) / (std::pow(r, 1.5) + a * std::sqrt(m));
}
inline double page_thorne_flux(double r, const Kerr<double>& bh, double m_dot) {  // disk.rs:90-151
    double m = bh.m, a = bh.a();
    double r_isco = bh.isco(true);
    if (r <= r_isco) return 0.0;
    double e_r = specific_energy(r, m, a);
    double lz_r = specific_angular_momentum(r, m, a);
    double omega_r = angular_velocity(r, m, a);
    double denom = e_r - omega_r * lz_r;
    if (std::fabs(denom) < 1e-30) return 0.0;
    double dr = r * 1e-5;
    double omega_dr = (angular_velocity(r + dr, m, a) - angular_velocity(r - dr, m, a)) / (2.0 * dr);
    const size_t n = 200;
    double h = (r - r_isco) / (double)n;
    if (h <= 0.0) return 0.0;
    auto integrand = [m, a](double rp) {
        double ep = specific_energy(rp, m, a);
        double lzp = specific_angular_momentum(rp, m, a);
        double omp = angular_velocity(rp, m, a);
        double drp = rp * 1e-5;
        double dlz_dr =
            (specific_angular_momentum(rp + drp, m, a) - specific_angular_momentum(rp - drp, m, a)) / (2.0 * drp);
        return (ep - omp * lzp) * dlz_dr;
    };
    double sum = integrand(r_isco) + integrand(r);
    for (size_t i = 1; i < n; i++) {
        double rp = r_isco + (double)i * h;
        double weight = (i % 2 == 0) ? 2.0 : 4.0;
        sum += weight * integrand(rp);
    }
    double integral = sum * h / 3.0;
    double flux = -(omega_dr / (denom * denom)) * integral;
    return std::fabs(flux) * m_dot;
}
inline double temperature(double r, const Kerr<double>& bh, double m_dot) {  // disk.rs:160-170
    double flux = page_thorne_flux(r, bh, m_dot);
    if (flux <= 0.0) return 0.0;
    double t_scale = 1e7 * std::pow(m_dot, 0.25);
    return t_scale * std::pow(flux, 0.25);
}
inline void generate_temperature_lut(const Kerr<double>& bh, size_t width, float* out) {  // disk.rs:175-201
    double rin = bh.isco(true);
    double rout = 50.0 * bh.m;
    double max_temp = 0.0;
    std::vector<double> temps(width);
    for (size_t i = 0; i < width; i++) {
        double t = (double)i / (double)std::max<size_t>(width - 1, 1);
        double r = rin + t * (rout - rin);
        double temp = temperature(r, bh, 1.0);
        if (temp > max_temp) max_temp = temp;
        temps[i] = temp;
    }
    double norm = (max_temp > 0.0) ? 1.0 / max_temp : 1.0;
    for (size_t i = 0; i < width; i++) out[i] = (float)(temps[i] * norm);
}
}  // namespace disk

// ---------------------------------------------------------------------------------------------
// Composite RGBA oracle (SURVEY §8c "reference-derived, unpinned" — this file is the spec for the two
// joints the reference never assembled in f64: camera->state and LUT sampling).
// ---------------------------------------------------------------------------------------------
// types/webgpu.ts:67-116 — 88 f32: view, proj, inv_view, inv_proj, prev_view_proj (column-major mat4 as
// gl-matrix writes them), position.xyz+pad, direction.xyz+pad.
struct CameraUniforms { float f[88]; };

struct RenderParams {
    double mass = 1.0, spin = 0.999;
    uint32_t width = 0, height = 0;
    uint32_t frame_index = 0;
    int jitter = 0;            // compute.wgsl.ts:153-157 Halton(2,3) sub-pixel jitter
    int coords = KERR_SCHILD;
    Options opts;
    double disk_r_out = 50.0;  // disk.rs:177 (rout = 50 M); inner edge = ISCO prograde (kerr.rs:100-123)
    double lut_max_temp = 1e7;
};

struct Luts {
    const float* spectrum = nullptr;  // W*H*4 f32 (a19)
    uint32_t spec_w = 0, spec_h = 0;
    const float* tdisk = nullptr;     // 512 f32 (a20)
    uint32_t tdisk_n = 0;
    double tdisk_rin = 0.0, tdisk_rout = 0.0;
};

// compute.wgsl.ts:135-145 halton
inline double halton(uint32_t index, uint32_t base) {
    double result = 0.0, f = 1.0 / (double)base;
    uint32_t i = index;
    while (i > 0u) { result += f * (double)(i % base); i = i / base; f = f / (double)base; }
    return result;
}

// compute.wgsl.ts:149-187 camera -> (x, p), evaluated in R from the f32 uniform block.
template <class R>
inline State<R> camera_ray(const CameraUniforms& cam, const RenderParams& rp, uint32_t px, uint32_t py) {
    const float* inv_view = cam.f + 32;
    const float* inv_proj = cam.f + 48;
    R width = R((double)rp.width), height = R((double)rp.height);
    R jx(0.0), jy(0.0);
    if (rp.jitter) {
        jx = (R(halton((rp.frame_index % 8u) + 1u, 2u)) - R(0.5)) / width;
        jy = (R(halton((rp.frame_index % 8u) + 1u, 3u)) - R(0.5)) / height;
    }
    R uvx = R((double)px) / width, uvy = R((double)py) / height;
    R ndcx = (uvx + jx) * R(2.0) - R(1.0);
    R ndcy = (uvy + jy) * R(2.0) - R(1.0);
    R clip[4] = {ndcx, -ndcy, R(1.0), R(1.0)};
    R vt[4];
    for (int row = 0; row < 4; row++) {
        vt[row] = R((double)inv_proj[0 * 4 + row]) * clip[0] + R((double)inv_proj[1 * 4 + row]) * clip[1] +
                  R((double)inv_proj[2 * 4 + row]) * clip[2] + R((double)inv_proj[3 * 4 + row]) * clip[3];
    }
    R vx = vt[0] / vt[3], vy = vt[1] / vt[3], vz = vt[2] / vt[3];
    R vn = sqrt(vx * vx + vy * vy + vz * vz);
    vx = vx / vn; vy = vy / vn; vz = vz / vn;
    R w[3];
    for (int row = 0; row < 3; row++) {
        w[row] = R((double)inv_view[0 * 4 + row]) * vx + R((double)inv_view[1 * 4 + row]) * vy +
                 R((double)inv_view[2 * 4 + row]) * vz;
    }
    R wn = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    R dx = w[0] / wn, dy = w[1] / wn, dz = w[2] / wn;

    R cx = R((double)cam.f[80]), cy = R((double)cam.f[81]), cz = R((double)cam.f[82]);
    R r0 = sqrt(cx * cx + cy * cy + cz * cz);
    R theta0 = acos(rclamp<R>(cy / r0, R(-1.0), R(1.0)));
    R phi0 = R(std::atan2(to_double(cz), to_double(cx)));
    R st = sin(theta0), ct = cos(theta0), sp = sin(phi0), cp = cos(phi0);
    R pr_far = dx * (st * cp) + dy * ct + dz * (st * sp);
    R pth_far = (dx * (ct * cp) + dy * (-st) + dz * (ct * sp)) / r0;
    R safe_st = rmax<R>(st, R(1e-4));
    R pph_far = (dx * (-sp) + dz * cp) / (r0 * safe_st);
    State<R> s;
    s.x[0] = R(0.0); s.x[1] = r0; s.x[2] = theta0; s.x[3] = phi0;
    s.p[0] = R(-1.0); s.p[1] = pr_far; s.p[2] = pth_far * r0 * r0; s.p[3] = pph_far * r0 * r0 * st * st;
    return s;
}

// Linear sample of the 1-D disk temperature LUT (a20): entry i <-> r = rin + i/(n-1) (rout-rin).
template <class R>
inline R sample_tdisk(const Luts& l, R r) {
    R t = (r - R(l.tdisk_rin)) / (R(l.tdisk_rout) - R(l.tdisk_rin)) * R((double)(l.tdisk_n - 1));
    t = rclamp<R>(t, R(0.0), R((double)(l.tdisk_n - 1)));
    R fl = floor(t);
    uint32_t i0 = (uint32_t)to_double(fl);
    uint32_t i1 = std::min<uint32_t>(i0 + 1, l.tdisk_n - 1);
    R f = t - fl;
    R a = R((double)l.tdisk[i0]), b = R((double)l.tdisk[i1]);
    return a + (b - a) * f;
}

// Bilinear, clamp-to-edge, texel-centre (GL LINEAR) sample of the RGBA f32 spectrum LUT (a19):
// x = u*W - 0.5, y = v*H - 0.5 (filter/wrap as rendering/spectral.ts:52-54 configures the texture).
template <class R>
inline void sample_spectrum(const Luts& l, R u, R v, R rgb[3]) {
    R x = u * R((double)l.spec_w) - R(0.5);
    R y = v * R((double)l.spec_h) - R(0.5);
    x = rclamp<R>(x, R(0.0), R((double)(l.spec_w - 1)));
    y = rclamp<R>(y, R(0.0), R((double)(l.spec_h - 1)));
    R xf = floor(x), yf = floor(y);
    uint32_t x0 = (uint32_t)to_double(xf), y0 = (uint32_t)to_double(yf);
    uint32_t x1 = std::min<uint32_t>(x0 + 1, l.spec_w - 1), y1 = std::min<uint32_t>(y0 + 1, l.spec_h - 1);
    R fx = x - xf, fy = y - yf;
    const float* t00 = l.spectrum + 4 * ((size_t)y0 * l.spec_w + x0);
    const float* t10 = l.spectrum + 4 * ((size_t)y0 * l.spec_w + x1);
    const float* t01 = l.spectrum + 4 * ((size_t)y1 * l.spec_w + x0);
    const float* t11 = l.spectrum + 4 * ((size_t)y1 * l.spec_w + x1);
    for (int c = 0; c < 3; c++) {
        R top = R((double)t00[c]) + (R((double)t10[c]) - R((double)t00[c])) * fx;
        R bot = R((double)t01[c]) + (R((double)t11[c]) - R((double)t01[c])) * fx;
        rgb[c] = top + (bot - top) * fy;
    }
}

// Thin-disk crossing shade + front-to-back composite (compute.wgsl.ts:216-254 with a18 + a19 + a20
// swapped in for the artistic ramp, per north_star).
template <class R>
struct DiskHook {
    const Luts* luts;
    R mass, spin, r_in, r_out;
    R color[3] = {R(0.0), R(0.0), R(0.0)};
    R alpha = R(0.0);
    uint32_t crossings = 0;
    bool operator()(const State<R>& prev, const State<R>& cur) {
        const R half_pi(1.5707963267948966);
        R d0 = prev.x[2] - half_pi, d1 = cur.x[2] - half_pi;
        if (d0 * d1 <= R(0.0)) {
            R dth = cur.x[2] - prev.x[2];
            R f = (dth == R(0.0)) ? R(0.0) : (half_pi - prev.x[2]) / dth;
            R r_c = prev.x[1] + f * (cur.x[1] - prev.x[1]);
            if (r_c > r_in && r_c < r_out) {
                crossings++;
                R lambda = cur.p[3] / (-cur.p[0]);
                R g = kerr_g_factor<R>(r_c, mass, spin, lambda);
                R tn = sample_tdisk<R>(*luts, r_c);
                R u = pow(tn, R(0.4));             // inverse of spectrum.rs:86 (x/(W-1))^2.5
                R v = (g - R(0.05)) / R(4.95);     // inverse of spectrum.rs:82
                R rgb[3];
                sample_spectrum<R>(*luts, u, v, rgb);
                R opacity = R(0.6) * tn * g;       // compute.wgsl.ts:228,233 with T(r) from a20
                R w = (R(1.0) - alpha) * opacity;
                for (int c = 0; c < 3; c++) color[c] += rgb[c] * w;   // :250-253
                alpha += opacity;                  // :254
            }
        }
        return alpha > R(0.99);                    // :256
    }
};

// ---------------------------------------------------------------------------------------------
// GLSL-semantics path: the production WebGL2 fragment shader's Cartesian Velocity-Verlet march on
// its pseudo-Kerr acceleration field. src/shaders/blackhole/chunks/metric.ts:96-149 (kerr_geodesic_accel),
// fragment.glsl.ts:129-221 (loop, step-size heuristics, ZAMO rotation), chunks/common.ts:40-49 (constants, rot()).
// Deterministic subset: blue-noise dither = 0, lensing strength 1, no volumetric disk / jets / stars (they sample
// Math.random noise textures); disk light comes from the same thin-disk crossing composite as the Hamiltonian path,
// evaluated at the equatorial crossing point the shader itself computes (chunks/disk.ts:22-30).
// ---------------------------------------------------------------------------------------------
template <class R> struct V3 { R x, y, z; };
template <class R> inline R dot3(V3<R> a, V3<R> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class R> inline V3<R> cross3(V3<R> a, V3<R> b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

template <class R>
inline void glsl_kerr_accel(V3<R> p, V3<R> v, R M, R a, V3<R>& accel, R& omega) {   // metric.ts:96-149
    R a2 = a * a;
    R rho2 = dot3(p, p);
    R diff = rho2 - a2;
    R disc = diff * diff + R(4.0) * a2 * p.y * p.y;
    R r2 = R(0.5) * (diff + sqrt(rmax<R>(R(0.0), disc)));
    R r_k = sqrt(rmax<R>(R(1e-8), r2));
    R sigma = r2 + a2 * (p.y * p.y / rmax<R>(R(1e-8), r2));
    V3<R> L = cross3(p, v);
    R Ly = L.y;
    R Ly_eff = Ly - a;
    R L2_eff = Ly_eff * Ly_eff + (dot3(L, L) - Ly * Ly);
    R r_inv = R(1.0) / r_k;
    R r2_inv = r_inv * r_inv;
    R r4_inv = r2_inv * r2_inv;
    R sigma_ratio = r2 / rmax<R>(R(1e-8), sigma);
    R pn = sqrt(dot3(p, p));
    V3<R> r_hat = {-(p.x / pn), -(p.y / pn), -(p.z / pn)};
    R f = M * r2_inv * sigma_ratio + R(3.0) * M * rmax<R>(R(0.0), L2_eff) * r4_inv * sigma_ratio;
    accel = {r_hat.x * f, r_hat.y * f, r_hat.z * f};
    R r3_p_a2r = r_k * r2 + a2 * r_k;
    R drag = R(2.0) * M * a / rmax<R>(R(1e-8), r3_p_a2r);
    accel.x += v.z * drag;          // cross((0,1,0), v) = (v.z, 0, -v.x)
    accel.z += -v.x * drag;
    omega = R(2.0) * M * a / rmax<R>(R(1e-8), r3_p_a2r);
}
template <class R> inline R glsl_smoothstep(R e0, R e1, R x) {
    R t = rclamp<R>((x - e0) / (e1 - e0), R(0.0), R(1.0));
    return t * t * (R(3.0) - R(2.0) * t);
}

struct PixelResult {
    double rgba[4];
    double xp[8];
    uint32_t termination;
    uint32_t steps;
    double max_drift;
    uint32_t crossings;
    uint64_t attempts, rhs_evals;
};

template <class R>
inline PixelResult render_pixel(const CameraUniforms& cam, const RenderParams& rp, const Luts& luts, uint32_t px,
                                uint32_t py) {
    Kerr<R> metric(R(rp.mass), R(rp.spin), rp.coords);
    State<R> s0 = camera_ray<R>(cam, rp, px, py);
    DiskHook<R> hook;
    hook.luts = &luts;
    hook.mass = R(rp.mass); hook.spin = metric.spin;
    hook.r_in = metric.isco(true); hook.r_out = R(rp.disk_r_out);
    // integrate() takes the hook by value; keep the accumulators via a reference wrapper
    struct Ref { DiskHook<R>* h; bool operator()(const State<R>& a, const State<R>& b) { return (*h)(a, b); } };
    Trajectory<R> t = integrate<R, Ref>(s0, metric, rp.opts, Ref{&hook});
    PixelResult out;
    for (int c = 0; c < 3; c++) out.rgba[c] = to_double(hook.color[c]);
    out.rgba[3] = 1.0;  // compute.wgsl.ts:257
    for (int i = 0; i < 4; i++) { out.xp[i] = to_double(t.final_state.x[i]); out.xp[4 + i] = to_double(t.final_state.p[i]); }
    out.termination = t.termination;
    out.steps = (uint32_t)t.steps_taken;
    out.max_drift = to_double(t.max_hamiltonian_drift);
    out.crossings = hook.crossings;
    out.attempts = t.attempts; out.rhs_evals = t.rhs_evals;
    return out;
}

// world-space ray of a pixel: the direction part of camera_ray() (compute.wgsl.ts:159-171)
template <class R>
inline void camera_dir(const CameraUniforms& cam, const RenderParams& rp, uint32_t px, uint32_t py, V3<R>& ro, V3<R>& rd) {
    const float* inv_view = cam.f + 32;
    const float* inv_proj = cam.f + 48;
    R width = R((double)rp.width), height = R((double)rp.height);
    R jx(0.0), jy(0.0);
    if (rp.jitter) {
        jx = (R(halton((rp.frame_index % 8u) + 1u, 2u)) - R(0.5)) / width;
        jy = (R(halton((rp.frame_index % 8u) + 1u, 3u)) - R(0.5)) / height;
    }
    R ndcx = (R((double)px) / width + jx) * R(2.0) - R(1.0);
    R ndcy = (R((double)py) / height + jy) * R(2.0) - R(1.0);
    R clip[4] = {ndcx, -ndcy, R(1.0), R(1.0)};
    R vt[4];
    for (int row = 0; row < 4; row++)
        vt[row] = R((double)inv_proj[row]) * clip[0] + R((double)inv_proj[4 + row]) * clip[1] +
                  R((double)inv_proj[8 + row]) * clip[2] + R((double)inv_proj[12 + row]) * clip[3];
    R vx = vt[0] / vt[3], vy = vt[1] / vt[3], vz = vt[2] / vt[3];
    R vn = sqrt(vx * vx + vy * vy + vz * vz);
    vx = vx / vn; vy = vy / vn; vz = vz / vn;
    R w[3];
    for (int row = 0; row < 3; row++)
        w[row] = R((double)inv_view[row]) * vx + R((double)inv_view[4 + row]) * vy + R((double)inv_view[8 + row]) * vz;
    R wn = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    rd = {w[0] / wn, w[1] / wn, w[2] / wn};
    ro = {R((double)cam.f[80]), R((double)cam.f[81]), R((double)cam.f[82])};
}

// fragment.glsl.ts:89-221 with the deterministic subset described above. xp = (p, v, photonCrossings, 0).
template <class R>
inline PixelResult render_pixel_verlet(const CameraUniforms& cam, const RenderParams& rp, const Luts& luts, uint32_t px,
                                       uint32_t py) {
    const R M = R(rp.mass);
    Kerr<R> metric(M, R(rp.spin), KERR_SCHILD);
    const R a = metric.spin * M;
    const R rh = M + sqrt(rmax<R>(R(0.0), M * M - a * a));                       // kerr_horizon, metric.ts:13-15
    R a_star = rclamp<R>(a / M, R(-0.9999), R(0.9999));                           // kerr_photon_sphere, metric.ts:32-37
    const R rph = R(2.0) * M * (R(1.0) + cos(R(2.0 / 3.0) * acos(rclamp<R>(-a_star, R(-1.0), R(1.0)))));
    const R r_in = metric.isco(true), r_out = R(rp.disk_r_out);
    const R MIN_STEP(0.01), MAX_STEP(1.2), max_dist = R(rp.opts.escape_radius);   // common.ts:41-43
    V3<R> ro, rd;
    camera_dir<R>(cam, rp, px, py, ro, rd);
    R ron = sqrt(dot3(ro, ro));
    if (ron < rh * R(1.5)) { R k = rh * R(1.5) / ron; ro = {ro.x * k, ro.y * k, ro.z * k}; }   // fragment.glsl.ts:94-97
    V3<R> p = ro, v = rd;
    R col[3] = {R(0.0), R(0.0), R(0.0)}, alpha(0.0);
    uint32_t term = TERM_MAXSTEPS, steps = 0, crossings = 0, photon = 0;
    R prevY = p.y;
    V3<R> c0 = cross3(ro, rd);
    bool hit = sqrt(dot3(c0, c0)) < rh * R(0.9);                                  // fragment.glsl.ts:124-126
    const uint64_t max_steps = rp.opts.max_steps < 500 ? rp.opts.max_steps : 500;  // :115
    for (uint64_t i = 0; i < max_steps; i++) {
        V3<R> p_prev = p;
        R r = sqrt(dot3(p, p));
        if (r < rh * R(1.15)) { hit = true; term = TERM_HORIZON; break; }          // :135-138 (horizonThreshold 1.15)
        if (r > max_dist) { term = TERM_ESCAPE; break; }
        R distFactor = R(1.0) + r * R(0.05);
        R dt = rclamp<R>((r - rh) * R(0.1) * distFactor, MIN_STEP, MAX_STEP * distFactor);
        if (r > R(30.0)) {
            R farBoost = (r - R(30.0)) * R(0.08);
            dt = rmax<R>(dt, MIN_STEP + farBoost);
            dt = rmin<R>(dt, MAX_STEP * R(2.5));
        }
        R sphereProx = fabs(r - rph);
        dt = rmin<R>(dt, MIN_STEP + sphereProx * R(0.15));
        R hRef = glsl_smoothstep<R>(R(0.2), R(0.0), fabs(p.y));
        R cdt = dt * (R(1.0) - hRef * R(0.7));
        V3<R> acc; R omega;
        glsl_kerr_accel<R>(p, v, M, a, acc, omega);
        {   // v.xz *= rot(omega * cdt), rot(t) = mat2(c, -s, s, c)  (common.ts:46-49)
            R ang = omega * cdt, sn = sin(ang), cs = cos(ang);
            R nx = v.x * cs - v.z * sn, nz = v.x * sn + v.z * cs;
            v.x = nx; v.z = nz;
        }
        p = {p.x + v.x * cdt + R(0.5) * acc.x * cdt * cdt, p.y + v.y * cdt + R(0.5) * acc.y * cdt * cdt,
             p.z + v.z * cdt + R(0.5) * acc.z * cdt * cdt};
        R r_new = sqrt(dot3(p, p));
        if (alpha < R(0.95)) {
            V3<R> acc2; R om2;
            glsl_kerr_accel<R>(p, v, M, a, acc2, om2);
            v = {v.x + R(0.5) * (acc.x + acc2.x) * cdt, v.y + R(0.5) * (acc.y + acc2.y) * cdt,
                 v.z + R(0.5) * (acc.z + acc2.z) * cdt};
        }
        R vn = sqrt(dot3(v, v));
        v = {v.x / vn, v.y / vn, v.z / vn};
        if (prevY * p.y < R(0.0) && r_new < rph * R(2.0) && r_new > rh) photon = photon + 1 < 3 ? photon + 1 : 3;
        prevY = p.y;
        steps++;
        if (p_prev.y * p.y < R(0.0)) {                                            // chunks/disk.ts:22-30
            R t = fabs(p_prev.y) / rmax<R>(R(0.0001), fabs(p_prev.y) + fabs(p.y));
            V3<R> sp = {p_prev.x + (p.x - p_prev.x) * t, p_prev.y + (p.y - p_prev.y) * t, p_prev.z + (p.z - p_prev.z) * t};
            R r_c = sqrt(dot3(sp, sp));
            if (r_c > r_in && r_c < r_out) {
                crossings++;
                R lambda = p.z * v.x - p.x * v.z;                                  // L_photon, chunks/disk.ts:92
                R g = kerr_g_factor<R>(r_c, M, metric.spin, lambda);
                R tn = sample_tdisk<R>(luts, r_c);
                R rgb[3];
                sample_spectrum<R>(luts, pow(tn, R(0.4)), (g - R(0.05)) / R(4.95), rgb);
                R opacity = R(0.6) * tn * g;
                R w = (R(1.0) - alpha) * opacity;
                for (int c = 0; c < 3; c++) col[c] += rgb[c] * w;
                alpha += opacity;
            }
        }
        if (alpha > R(0.99)) { term = TERM_DISK; break; }
    }
    PixelResult out;
    for (int c = 0; c < 3; c++) out.rgba[c] = to_double(col[c]);
    out.rgba[3] = 1.0;
    out.xp[0] = to_double(p.x); out.xp[1] = to_double(p.y); out.xp[2] = to_double(p.z);
    out.xp[3] = to_double(v.x); out.xp[4] = to_double(v.y); out.xp[5] = to_double(v.z);
    out.xp[6] = (double)photon; out.xp[7] = hit ? 1.0 : 0.0;
    out.termination = term; out.steps = steps; out.max_drift = 0.0; out.crossings = crossings;
    out.attempts = steps; out.rhs_evals = 2ull * steps;
    return out;
}

}  // namespace orc
