// gvt_hostmath.h — host-side, once-per-frame / once-per-parameter-change mathematics of the product:
// closed-form hole radii, the two LUT generators, the Bardeen shadow curve and the camera kinematic filter.
// None of this is per-pixel work (the reference runs it once per tick or once per parameter change on the
// worker thread); the per-pixel path lives in gvt_kernels.cu and has no host implementation.
// Each function cites the reference lines whose results it must reproduce.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <thread>
#include <utility>
#include <vector>

namespace gvt {
namespace host {

struct Hole {
    double mass, spin;  // spin clamped to [-1,1] (kerr.rs:49-55)
    Hole(double m, double s) : mass(m), spin(std::min(1.0, std::max(-1.0, s))) {}
    double a() const { return spin * mass; }
    double horizon() const {  // metric/mod.rs:75-84
        const double aa = spin * mass, disc = mass * mass - aa * aa;
        return disc < 0.0 ? mass : mass + std::sqrt(disc);
    }
    double photon_sphere() const {  // kerr.rs:91-94
        return 2.0 * mass * (1.0 + std::cos((2.0 / 3.0) * std::acos(-spin)));
    }
    double isco(bool prograde) const {  // kerr.rs:100-123 (Bardeen-Press-Teukolsky)
        if (std::fabs(spin) < 1e-6) return mass * 6.0;
        const double a2 = spin * spin;
        const double z1 = 1.0 + std::pow(1.0 - a2, 1.0 / 3.0) *
                                    (std::pow(1.0 + spin, 1.0 / 3.0) + std::pow(1.0 - spin, 1.0 / 3.0));
        const double z2 = std::sqrt(3.0 * a2 + z1 * z1);
        const double disc = (3.0 - z1) * (3.0 + z1 + 2.0 * z2);
        const double root = disc < 0.0 ? 0.0 : std::sqrt(disc);
        return mass * (3.0 + z2 + (prograde ? -1.0 : 1.0) * root);
    }
    // lib.rs:97-105 over kerr.rs:180-188,242-264: 1/sqrt(-g_tt) at the equator, 100 inside the ergoregion
    double dilation(double r) const {
        const double aa = a(), th = 1.5707963267948966;
        const double c = std::cos(th);
        const double sigma = r * r + aa * aa * (c * c);
        const double g_tt = -(1.0 - (2.0 * mass * r) / sigma);
        if (g_tt >= 0.0) return 100.0;
        const double td = std::sqrt(-g_tt);
        return td <= 0.0 ? 100.0 : 1.0 / td;
    }
};

inline void parallel_rows(size_t n, const std::function<void(size_t)>& body) {
    unsigned nt = std::max(1u, std::thread::hardware_concurrency());
    if (nt > n) nt = (unsigned)n;
    if (nt <= 1) { for (size_t i = 0; i < n; i++) body(i); return; }
    std::atomic<size_t> next{0};
    auto work = [&]() { for (;;) { size_t i = next.fetch_add(1); if (i >= n) return; body(i); } };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
}

// ---- physics/spectrum.rs:12-102 : Planck -> CIE 1931 fit -> linear sRGB, 2-D (T, g) LUT -------------
namespace spec {
constexpr double kC = 299792458.0, kKb = 1.380649e-23, kH = 6.62607015e-34;
constexpr double kC1 = 2.0 * kH * kC * kC, kC2 = kH * kC / kKb;
inline double planck(double lambda, double T) {
    const double ex = kC2 / (lambda * T);
    if (ex > 100.0) return 0.0;
    const double l5 = lambda * lambda * lambda * lambda * lambda;
    return (kC1 / l5) / (std::exp(ex) - 1.0);
}
inline double lobe(double l_nm, double mean, double sd) { const double x = (l_nm - mean) / sd; return std::exp(-0.5 * x * x); }
inline void texel(double T_eff, double g, float out[4]) {
    double X = 0.0, Y = 0.0, Z = 0.0;
    if (!(T_eff < 100.0)) {
        const double end = 780.0e-9, step = 2.0e-9;
        for (double lambda = 380.0e-9; lambda <= end; lambda += step) {
            const double I = planck(lambda, T_eff);
            const double l_nm = lambda * 1e9;
            const double cx = std::max(1.056 * lobe(l_nm, 599.0, 37.9) + 0.362 * lobe(l_nm, 442.0, 16.0) - 0.065 * lobe(l_nm, 501.0, 20.4), 0.0);
            const double cy = std::max(0.821 * lobe(l_nm, 568.0, 46.9) + 0.286 * lobe(l_nm, 530.0, 22.1), 0.0);
            const double cz = std::max(1.217 * lobe(l_nm, 437.0, 11.8) + 0.681 * lobe(l_nm, 459.0, 26.0), 0.0);
            X += I * cx * step; Y += I * cy * step; Z += I * cz * step;
        }
    }
    const double r = 3.2404542 * X - 1.5371385 * Y - 0.4985314 * Z;
    const double gg = -0.9692660 * X + 1.8760108 * Y + 0.0415560 * Z;
    const double b = 0.0556434 * X - 0.2040259 * Y + 1.0572252 * Z;
    const float scale = (float)(1.0e-14 * (g * g * g * g));
    out[0] = (float)std::max(r, 0.0) * scale;
    out[1] = (float)std::max(gg, 0.0) * scale;
    out[2] = (float)std::max(b, 0.0) * scale;
    out[3] = 1.0f;
}
}  // namespace spec

inline void spectrum_lut(uint32_t W, uint32_t H, double max_temp, float* out) {
    parallel_rows(H, [&](size_t y) {
        const double g = 0.05 + (5.0 - 0.05) * ((double)y / (double)std::max<size_t>((size_t)H - 1, 1));
        for (uint32_t x = 0; x < W; x++) {
            const double T = std::pow((double)x / (double)std::max<size_t>((size_t)W - 1, 1), 2.5) * max_temp;
            spec::texel(T * g, g, out + 4 * (y * W + x));
        }
    });
}

// ---- physics/disk.rs:24-201 : Page-Thorne flux -> normalised T(r) table over [r_isco, 50 M] ----------
namespace nt {
inline double den_sq(double r, double m, double a) { return 1.0 - 3.0 / (r / m) + 2.0 * (a / m) * std::sqrt(m / r); }
inline double energy(double r, double m, double a) {
    const double d = den_sq(r, m, a);
    if (d <= 0.0) return 1.0;
    return (1.0 - 2.0 / (r / m) + (a / m) * std::sqrt(m / r)) / std::sqrt(d);
}
inline double ang_mom(double r, double m, double a) {
    const double d = den_sq(r, m, a);
    const double ar = a / r;
    const double num = std::sqrt(m) * std::sqrt(r) * (1.0 - 2.0 * (a / m) * std::sqrt(m / r) + ar * ar);
    if (d <= 0.0) return 0.0;
    return num / std::sqrt(d);
}
inline double omega(double r, double m, double a) { return std::sqrt(m) / (std::pow(r, 1.5) + a * std::sqrt(m)); }
inline double integrand(double rp, double m, double a) {
    const double drp = rp * 1e-5;
    const double dl = (ang_mom(rp + drp, m, a) - ang_mom(rp - drp, m, a)) / (2.0 * drp);
    return (energy(rp, m, a) - omega(rp, m, a) * ang_mom(rp, m, a)) * dl;
}
inline double flux(double r, const Hole& bh, double m_dot) {
    const double m = bh.mass, a = bh.a(), r_isco = bh.isco(true);
    if (r <= r_isco) return 0.0;
    const double denom = energy(r, m, a) - omega(r, m, a) * ang_mom(r, m, a);
    if (std::fabs(denom) < 1e-30) return 0.0;
    const double dr = r * 1e-5;
    const double dom = (omega(r + dr, m, a) - omega(r - dr, m, a)) / (2.0 * dr);
    const size_t n = 200;
    const double h = (r - r_isco) / (double)n;
    if (h <= 0.0) return 0.0;
    double sum = integrand(r_isco, m, a) + integrand(r, m, a);
    for (size_t i = 1; i < n; i++) sum += ((i % 2 == 0) ? 2.0 : 4.0) * integrand(r_isco + (double)i * h, m, a);
    const double integral = sum * h / 3.0;
    return std::fabs(-(dom / (denom * denom)) * integral) * m_dot;
}
inline double temperature(double r, const Hole& bh, double m_dot) {
    const double f = flux(r, bh, m_dot);
    if (f <= 0.0) return 0.0;
    return (1e7 * std::pow(m_dot, 0.25)) * std::pow(f, 0.25);
}
}  // namespace nt

inline void disk_lut(const Hole& bh, uint32_t n, float* out) {
    const double rin = bh.isco(true), rout = 50.0 * bh.mass;
    std::vector<double> T(n);
    double tmax = 0.0;
    for (uint32_t i = 0; i < n; i++) {
        const double t = (double)i / (double)std::max<size_t>((size_t)n - 1, 1);
        T[i] = nt::temperature(rin + t * (rout - rin), bh, 1.0);
        if (T[i] > tmax) tmax = T[i];
    }
    const double norm = tmax > 0.0 ? 1.0 / tmax : 1.0;
    for (uint32_t i = 0; i < n; i++) out[i] = (float)(T[i] * norm);
}

// ---- physics/redshift.rs:65-95 (scalar getter of the PhysicsEngine seam, lib.rs:203-205) -------------
inline double g_factor(double r, double mass, double spin, double lambda) {
    const double a = spin * mass, r2 = r * r, a2 = a * a, m = mass;
    const double om = std::sqrt(m) / (std::pow(r, 1.5) + a * std::sqrt(m));
    const double sigma = r2;
    const double g_tt = -(1.0 - 2.0 * m * r / sigma), g_tphi = -(2.0 * m * r * a) / sigma;
    const double g_pp = r2 + a2 + 2.0 * m * r * a2 / sigma;
    const double den = -g_tt - 2.0 * om * g_tphi - om * om * g_pp;
    if (den <= 0.0) return 0.0;
    const double ut = 1.0 / std::sqrt(den), f = 1.0 - lambda * om;
    if (std::fabs(f) < 1e-30) return 0.0;
    return 1.0 / (ut * f);
}

// ---- physics/shadow.rs:38-58,81-183 : Bardeen critical curve ------------------------------------------
struct Crit { double xi, eta; };
inline Crit critical_params(double r, double m, double a) {
    const double r2 = r * r, r3 = r2 * r, a2 = a * a;
    const double denom = a * (r - m);
    if (std::fabs(denom) < 1e-30) return {0.0, 0.0};
    const double xi = -(r3 - 3.0 * m * r2 + a2 * r + a2 * m) / denom;
    const double denom2 = a2 * (r - m) * (r - m);
    if (std::fabs(denom2) < 1e-30) return {xi, 0.0};
    const double q = r - 3.0 * m;
    return {xi, r3 * (4.0 * m * a2 - r * (q * q)) / denom2};
}
inline std::vector<std::pair<double, double>> bardeen_shadow(const Hole& bh, double theta_obs, size_t n_points) {
    const double PI = 3.14159265358979323846;
    const double m = bh.mass, a = bh.a();
    const double so = std::sin(theta_obs), co = std::cos(theta_obs);
    std::vector<std::pair<double, double>> pts;
    if (std::fabs(a) < 1e-10) {
        const double radius = 3.0 * std::sqrt(3.0) * m;
        for (size_t i = 0; i < n_points; i++) {
            const double phi = 2.0 * PI * (double)i / (double)n_points;
            pts.emplace_back(radius * std::cos(phi), radius * std::sin(phi));
        }
        return pts;
    }
    if (std::fabs(so) < 1e-10) {
        const Crit p = critical_params(bh.photon_sphere(), m, a);
        const double radius = std::sqrt(std::max(p.eta + a * a, 0.0));
        for (size_t i = 0; i < 2 * n_points; i++) {
            const double phi = 2.0 * PI * (double)i / (2.0 * (double)n_points);
            pts.emplace_back(radius * std::cos(phi), radius * std::sin(phi));
        }
        return pts;
    }
    const double a_star = a / m;
    const double r_pro = 2.0 * m * (1.0 + std::cos((2.0 / 3.0) * std::acos(-std::fabs(a_star))));
    const double r_ret = 2.0 * m * (1.0 + std::cos((2.0 / 3.0) * std::acos(std::fabs(a_star))));
    auto beta_sq = [&](double r, Crit& p) {
        p = critical_params(r, m, a);
        return p.eta + a * a * co * co - p.xi * p.xi * co * co / (so * so);
    };
    double r_min = r_pro, r_max = r_ret;
    const int steps = 1000;
    Crit p;
    for (int i = 0; i <= steps; i++) {
        const double r = r_pro + ((double)i / (double)steps) * (r_ret - r_pro);
        if (beta_sq(r, p) >= 0.0) { r_min = r; break; }
    }
    for (int i = steps; i >= 0; i--) {
        const double r = r_pro + ((double)i / (double)steps) * (r_ret - r_pro);
        if (beta_sq(r, p) >= 0.0) { r_max = r; break; }
    }
    auto point = [&](size_t i, double sign) {
        const double phase = PI * (double)i / (double)std::max<size_t>(n_points - 1, 1);
        const double t = 0.5 - 0.5 * std::cos(phase);
        const double r = r_min + t * (r_max - r_min);
        const double b2 = beta_sq(r, p);
        const double alpha = a * so - p.xi / so;
        pts.emplace_back(alpha, sign * std::sqrt(std::max(b2, 0.0)));
    };
    for (size_t i = 0; i < n_points; i++) point(i, -1.0);
    for (size_t i = n_points; i-- > 0;) point(i, 1.0);
    return pts;
}

// ---- gravitas-core/src/spacetime/{curvature,lightcone,frame_drag,embedding}.rs : the visualisation helpers the
// PhysicsEngine exposes (lib.rs:139-305). One-off host maths in the reference too; all on the Boyer-Lindquist metric
// (lib.rs:63 `metric_bl`), covariant components as kerr.rs:241-264. -----------------------------------------------
namespace viz {
struct CovBL { double g_tt, g_rr, g_thth, g_phph, g_tph; };
inline CovBL covariant_bl(const Hole& bh, double r, double theta) {
    const double m = bh.mass, a = bh.a(), r2 = r * r, a2 = a * a;
    const double st = std::sin(theta), ct = std::cos(theta), sin2 = st * st, cos2 = ct * ct;
    const double sigma = r2 + a2 * cos2, delta = r2 - 2.0 * m * r + a2;
    CovBL g;
    g.g_tt = -(1.0 - (2.0 * m * r) / sigma);
    g.g_rr = sigma / delta;
    g.g_thth = sigma;
    g.g_phph = (r2 + a2 + (2.0 * m * r * a2 * sin2) / sigma) * sin2;
    g.g_tph = -(2.0 * m * r * a * sin2) / sigma;
    return g;
}
// curvature.rs:13-36 (raw mass / spin, as lib.rs:213-215 passes them)
inline double kretschner(double r, double theta, double mass, double spin) {
    const double a = spin * mass, r2 = r * r, a2 = a * a, c = std::cos(theta);
    const double cos2 = c * c, cos4 = cos2 * cos2, cos6 = cos4 * cos2;
    const double r4 = r2 * r2, r6 = r4 * r2, a4 = a2 * a2, a6 = a4 * a2;
    const double sigma = r2 + a2 * cos2, s2 = sigma * sigma, s4 = s2 * s2, sigma6 = s2 * s4;   // powi(6) by squaring
    if (sigma6 < 1e-30) return INFINITY;
    const double num = r6 - 15.0 * r4 * a2 * cos2 + 15.0 * r2 * a4 * cos4 - a6 * cos6;
    return 48.0 * mass * mass * num / sigma6;
}
// lightcone.rs:18-49 (the BL metric is diagonal in (t, r): only the first branch can be taken)
inline double light_cone_tilt(const Hole& bh, double r, double theta) {
    const CovBL g = covariant_bl(bh, r, theta);
    if (g.g_tt >= 0.0) return 1.5707963267948966;
    return std::atan(std::sqrt(std::max(-g.g_tt / g.g_rr, 0.0)));
}
// frame_drag.rs:13-15 over kerr.rs:143-152
inline double frame_drag_omega(const Hole& bh, double r, double theta) {
    const CovBL g = covariant_bl(bh, r, theta);
    return std::fabs(g.g_phph) < 1e-30 ? 0.0 : -g.g_tph / g.g_phph;
}
// the (r, theta, value) sampling lattice shared by curvature_field / tilt_field / frame_drag_field
template <class F> inline void field(double r_min, double r_max, size_t n_radial, size_t n_polar, float* out, F f) {
    const double PI = 3.14159265358979323846;
    for (size_t i = 0; i < n_radial; i++) {
        const double r = r_min + (r_max - r_min) * (double)i / (double)(n_radial - 1);
        for (size_t j = 0; j < n_polar; j++) {
            const double theta = 0.1 + (PI - 0.2) * (double)j / (double)(n_polar - 1);
            float* o = out + 3 * (i * n_polar + j);
            o[0] = (float)r; o[1] = (float)theta; o[2] = (float)f(r, theta);
        }
    }
}
inline double flamm_height(double r, double mass) {   // embedding.rs:14-20
    const double rs = 2.0 * mass;
    return r <= rs ? 0.0 : 2.0 * std::sqrt(rs * (r - rs));
}
inline double kerr_embedding_height(const Hole& bh, double r, double r_ref, size_t n_steps) {   // embedding.rs:28-44
    const double dr = (r_ref - r) / (double)n_steps;
    double z = 0.0;
    for (size_t i = 0; i < n_steps; i++) {
        const double r_i = r + ((double)i + 0.5) * dr;
        z += std::sqrt(std::fabs(covariant_bl(bh, r_i, 1.5707963267948966).g_rr - 1.0)) * dr;
    }
    return z;
}
inline double proper_distance(const Hole& bh, double r1, double r2, size_t n_steps) {   // embedding.rs:49-63
    const double lo = r1 < r2 ? r1 : r2, hi = r1 < r2 ? r2 : r1, dr = (hi - lo) / (double)n_steps;
    double d = 0.0;
    for (size_t i = 0; i < n_steps; i++) {
        const double r_i = lo + ((double)i + 0.5) * dr;
        d += std::sqrt(std::fabs(covariant_bl(bh, r_i, 1.5707963267948966).g_rr)) * dr;
    }
    return d;
}
inline void embedding_mesh(double mass, double spin, double r_min, double r_max, size_t n_radial, size_t n_angular, float* out) {
    const double PI = 3.14159265358979323846;   // embedding.rs:72-110
    const Hole bh(mass, spin);
    for (size_t i = 0; i < n_radial; i++) {
        const double t = (double)i / (double)(n_radial - 1), r = r_min + t * (r_max - r_min);
        const double height = std::fabs(spin) < 1e-6 ? flamm_height(r, mass) : kerr_embedding_height(bh, r, r_max, 100);
        for (size_t j = 0; j < n_angular; j++) {
            const double phi = 2.0 * PI * (double)j / (double)n_angular;
            float* o = out + 3 * (i * n_angular + j);
            o[0] = (float)(r * std::cos(phi)); o[1] = (float)(-height); o[2] = (float)(r * std::sin(phi));
        }
    }
}
inline void ergosphere_mesh(const Hole& bh, size_t n_polar, size_t n_azimuthal, float* out) {   // frame_drag.rs:48-68
    const double PI = 3.14159265358979323846, m = bh.mass, a = bh.a();
    for (size_t i = 0; i < n_polar; i++) {
        const double theta = PI * (double)i / (double)(n_polar - 1), c = std::cos(theta);
        const double disc = m * m - a * a * c * c, r_ergo = disc < 0.0 ? m : m + std::sqrt(disc);   // kerr.rs:157-167
        for (size_t j = 0; j < n_azimuthal; j++) {
            const double phi = 2.0 * PI * (double)j / (double)n_azimuthal;
            float* o = out + 3 * (i * n_azimuthal + j);
            o[0] = (float)(r_ergo * std::sin(theta) * std::cos(phi)); o[1] = (float)(r_ergo * std::cos(theta));
            o[2] = (float)(r_ergo * std::sin(theta) * std::sin(phi));
        }
    }
}
}  // namespace viz

// ---- gravitas-wasm/src/camera.rs:9-70 : camera state + kinematic filter (glam DVec3/DQuat maths inlined) --
struct Vec3 { double x, y, z; };
struct CameraState {
    Vec3 position{0.0, 0.0, 20.0};
    Vec3 velocity{0.0, 0.0, 0.0};
    double quat[4] = {0.0, 1.0, 0.0, 0.0};  // xyzw
    bool auto_spin = false;
    bool valid() const {
        auto f = [](double v) { return std::isfinite(v); };
        return f(position.x) && f(position.y) && f(position.z) && f(velocity.x) && f(velocity.y) && f(velocity.z) &&
               f(quat[0]) && f(quat[1]) && f(quat[2]) && f(quat[3]);
    }
};
// glam DQuat::from_rotation_y(angle).mul_vec3(v): q = (0, sin(angle/2), 0, cos(angle/2));
// v' = v (w^2 - b.b) + b (2 v.b) + (b x v)(2 w), b = (0, s, 0)
inline Vec3 rotate_y(double angle, Vec3 v) {
    const double s = std::sin(angle * 0.5), w = std::cos(angle * 0.5);
    const double b2 = s * s, k = w * w - b2, d2 = (v.y * s) * 2.0, w2 = w * 2.0;
    // b x v = (s*v.z, 0, -s*v.x)
    return {v.x * k + (s * v.z) * w2, v.y * k + s * d2, v.z * k + (-(s * v.x)) * w2};
}
inline void update_camera(double mouse_dx, double /*mouse_dy*/, double zoom_delta, double dt, CameraState& st) {
    if (dt <= 0.0) return;
    const double friction = std::exp(-5.0 * dt);
    st.velocity = {st.velocity.x * friction, st.velocity.y * friction, st.velocity.z * friction};
    st.position = {st.position.x + st.velocity.x * dt, st.position.y + st.velocity.y * dt, st.position.z + st.velocity.z * dt};
    st.position = rotate_y(-mouse_dx * 2.0 * dt, st.position);
    if (st.auto_spin) st.position = rotate_y(0.15 * dt, st.position);
    const double zf = 1.0 + zoom_delta * dt;
    st.position = {st.position.x * zf, st.position.y * zf, st.position.z * zf};
}

}  // namespace host
}  // namespace gvt
