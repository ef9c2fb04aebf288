// oracle_capi.cpp — extern "C" surface of the CPU ORACLE for ctypes (tests/, smoke(), bench.py's
// cpu_baseline / --impl reference legs ONLY). TEST INFRASTRUCTURE; never linked into the product.
// PARITY STATUS: see gravitas_oracle.hpp header ("parity unpinned" for integrate/LUT/RGBA values).
#include "gravitas_oracle.hpp"
#include "glsl_fragment_oracle.hpp"
#include <cstring>
#include <chrono>
#include <atomic>
#include <thread>
#include <functional>
#include <cstdlib>

using namespace orc;

extern "C" {

struct orc_options {
    int32_t method;               // 0 RKF45, 1 RK4, 2 symplectic (implicit midpoint)
    int32_t step_rule;            // 0 constant, 1 compute.wgsl.ts:213 rule
    double tolerance;
    double initial_step;
    uint64_t max_steps;
    double escape_radius;
    uint64_t renormalize_interval;
};

struct orc_render_params {
    double mass, spin;
    uint32_t width, height;
    uint32_t frame_index;
    int32_t jitter;
    int32_t coords;
    int32_t precision;            // 0 = f64, 1 = f32 (same algorithm instantiated on float)
    orc_options opts;
    double disk_r_out;
    double lut_max_temp;
    // LUTs
    const float* spectrum; uint32_t spec_w, spec_h;
    const float* tdisk; uint32_t tdisk_n; uint32_t _pad;
    double tdisk_rin, tdisk_rout;
};

static Options to_opts(const orc_options* o) {
    Options r;
    r.method = o->method; r.step_rule = o->step_rule; r.tolerance = o->tolerance; r.initial_step = o->initial_step;
    r.max_steps = o->max_steps; r.escape_radius = o->escape_radius; r.renormalize_interval = o->renormalize_interval;
    return r;
}
static State<double> load(const double* xp) {
    State<double> s;
    for (int i = 0; i < 4; i++) { s.x[i] = xp[i]; s.p[i] = xp[4 + i]; }
    return s;
}
static void store(const State<double>& s, double* xp) {
    for (int i = 0; i < 4; i++) { xp[i] = s.x[i]; xp[4 + i] = s.p[i]; }
}

// Thread pool in miniature (the image's g++ wrapper has no libgomp.spec, so no OpenMP): dynamic
// scheduling over [0,n) in chunks; ORC_THREADS overrides the hardware thread count.
static int g_threads = 0;
int orc_num_threads() {
    if (g_threads > 0) return g_threads;
    const char* e = std::getenv("ORC_THREADS");
    int n = e ? std::atoi(e) : (int)std::thread::hardware_concurrency();
    return n > 0 ? n : 1;
}
void orc_set_num_threads(int n) { g_threads = n; }
static void parallel_for(int64_t n, int64_t chunk, const std::function<void(int64_t)>& body) {
    int nt = orc_num_threads();
    if (nt <= 1 || n <= chunk) { for (int64_t i = 0; i < n; i++) body(i); return; }
    std::atomic<int64_t> next{0};
    auto worker = [&]() {
        for (;;) {
            int64_t b = next.fetch_add(chunk);
            if (b >= n) return;
            int64_t e2 = std::min<int64_t>(b + chunk, n);
            for (int64_t i = b; i < e2; i++) body(i);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
}

double orc_event_horizon(double m, double a, int coords) { return Kerr<double>(m, a, coords).event_horizon(); }
double orc_isco(double m, double a, int prograde) { return Kerr<double>(m, a, 0).isco(prograde != 0); }
double orc_photon_sphere(double m, double a) { return Kerr<double>(m, a, 0).photon_sphere(); }
double orc_time_dilation(double m, double a, double r, double theta) { return Kerr<double>(m, a, 0).time_dilation(r, theta); }

void orc_contravariant(double m, double a, int coords, double r, double theta, double* g16) {
    Kerr<double>(m, a, coords).contravariant(r, theta, g16);
}
void orc_hamiltonian_derivs(double m, double a, int coords, double r, double theta, const double* p4, double* out2) {
    Kerr<double>(m, a, coords).hamiltonian_derivatives(r, theta, p4, out2[0], out2[1]);
}
double orc_hamiltonian(double m, double a, int coords, const double* xp) {
    return hamiltonian(load(xp), Kerr<double>(m, a, coords));
}
void orc_rhs(double m, double a, int coords, const double* xp, double* out) {
    store(state_derivative(load(xp), Kerr<double>(m, a, coords)), out);
}
void orc_renormalize(double m, double a, int coords, double* xp) {
    State<double> s = load(xp);
    renormalize_null(s, Kerr<double>(m, a, coords));
    store(s, xp);
}
double orc_rkf45_step(double m, double a, int coords, const double* xp, double h, double* out) {
    double err;
    store(rkf45_step(load(xp), Kerr<double>(m, a, coords), h, err), out);
    return err;
}
// One AdaptiveStepper::step call (integrator.rs:72-107): xp updated in place, returns next h.
double orc_stepper_step(double m, double a, int coords, double tol, double* xp, double h_try) {
    State<double> s = load(xp);
    AdaptiveStepper<double> st(tol);
    double hn = st.step(s, Kerr<double>(m, a, coords), h_try);
    store(s, xp);
    return hn;
}
void orc_step_symplectic(double m, double a, int coords, double* xp, double h) {
    State<double> s = load(xp);
    step_symplectic(s, Kerr<double>(m, a, coords), h);
    store(s, xp);
}
void orc_step_rk4(double m, double a, int coords, double* xp, double h) {
    State<double> s = load(xp);
    step_rk4(s, Kerr<double>(m, a, coords), h);
    store(s, xp);
}

// geodesic::integrate over n independent rays (OpenMP over rays). stats per ray: [term, steps, attempts, rhs]
void orc_integrate(double m, double a, int coords, const orc_options* o, uint64_t n, const double* xp_in,
                   double* xp_out, uint32_t* term, uint64_t* steps, double* max_drift, uint64_t* attempts,
                   uint64_t* rejects, uint64_t* rhs_evals) {
    Options opts = to_opts(o);
    Kerr<double> metric(m, a, coords);
    parallel_for((int64_t)n, 16, [&](int64_t i) {
        Trajectory<double> t = integrate<double>(load(xp_in + 8 * i), metric, opts);
        store(t.final_state, xp_out + 8 * i);
        if (term) term[i] = t.termination;
        if (steps) steps[i] = t.steps_taken;
        if (max_drift) max_drift[i] = t.max_hamiltonian_drift;
        if (attempts) attempts[i] = t.attempts;
        if (rejects) rejects[i] = t.rejects;
        if (rhs_evals) rhs_evals[i] = t.rhs_evals;
    });
}

double orc_g_factor(double r, double mass, double spin, double lambda) { return kerr_g_factor<double>(r, mass, spin, lambda); }

void orc_spectrum_lut(uint32_t w, uint32_t h, double max_temp, float* out) {
    // rows are independent: parallelise without changing any texel's arithmetic
    parallel_for((int64_t)h, 1, [&](int64_t y) {
        // generate row y exactly as spectrum.rs:80-99 does
        const double min_g = 0.05, max_g = 5.0;
        double g = min_g + (max_g - min_g) * ((double)y / (double)std::max<size_t>((size_t)h - 1, 1));
        for (size_t x = 0; x < w; x++) {
            double t = std::pow((double)x / (double)std::max<size_t>((size_t)w - 1, 1), 2.5) * max_temp;
            double t_eff = t * g;
            double xyz[3];
            spectrum::integrate_planck_xyz(t_eff, xyz);
            float rgb[3];
            spectrum::xyz_to_linear_rgb(xyz[0], xyz[1], xyz[2], rgb);
            double g4 = g * g * g * g;
            double scale = 1.0e-14 * g4;
            float* px = out + 4 * ((size_t)y * w + x);
            px[0] = rgb[0] * (float)scale; px[1] = rgb[1] * (float)scale; px[2] = rgb[2] * (float)scale; px[3] = 1.0f;
        }
    });
}
// serial, literal loop (used to check the parallel version above bit-for-bit)
void orc_spectrum_lut_serial(uint32_t w, uint32_t h, double max_temp, float* out) {
    spectrum::generate_blackbody_lut(w, h, max_temp, out);
}
void orc_disk_lut(double mass, double spin, uint32_t n, float* out) {
    disk::generate_temperature_lut(Kerr<double>(mass, spin, 0), n, out);
}
double orc_page_thorne_flux(double r, double mass, double spin, double m_dot) {
    return disk::page_thorne_flux(r, Kerr<double>(mass, spin, 0), m_dot);
}
double orc_planck(double lambda, double t) { return spectrum::planck_law(lambda, t); }

static void to_rp(const orc_render_params* p, RenderParams& rp, Luts& l) {
    rp.mass = p->mass; rp.spin = p->spin; rp.width = p->width; rp.height = p->height;
    rp.frame_index = p->frame_index; rp.jitter = p->jitter; rp.coords = p->coords; rp.opts = to_opts(&p->opts);
    rp.disk_r_out = p->disk_r_out; rp.lut_max_temp = p->lut_max_temp;
    l.spectrum = p->spectrum; l.spec_w = p->spec_w; l.spec_h = p->spec_h;
    l.tdisk = p->tdisk; l.tdisk_n = p->tdisk_n; l.tdisk_rin = p->tdisk_rin; l.tdisk_rout = p->tdisk_rout;
}

void orc_camera_ray(const float* cam88, const orc_render_params* p, uint32_t px, uint32_t py, double* xp) {
    RenderParams rp; Luts l; to_rp(p, rp, l);
    CameraUniforms cam; std::memcpy(cam.f, cam88, sizeof(cam.f));
    store(camera_ray<double>(cam, rp, px, py), xp);
}

// Composite RGBA oracle over the pixel lattice {x0 + i*xs < width} x {y0 + j*ys < y1}. Outputs are dense
// arrays over that lattice (row-major, nx = ceil((width-x0)/xs) columns). Any output pointer may be null.
// Returns wall seconds; *total_steps / *total_rhs accumulate accepted steps / RHS evaluations.
double orc_render(const float* cam88, const orc_render_params* p, uint32_t x0, uint32_t xs, uint32_t y0, uint32_t y1,
                  uint32_t ys, double* rgba, double* xp, uint32_t* term, uint32_t* steps, double* drift,
                  uint32_t* crossings, uint64_t* total_steps, uint64_t* total_rhs) {
    RenderParams rp; Luts l; to_rp(p, rp, l);
    CameraUniforms cam; std::memcpy(cam.f, cam88, sizeof(cam.f));
    uint32_t nx = (rp.width - x0 + xs - 1) / xs;
    uint32_t ny = (y1 - y0 + ys - 1) / ys;
    std::atomic<uint64_t> tsteps{0}, trhs{0};
    auto t0 = std::chrono::steady_clock::now();
    parallel_for((int64_t)ny, 1, [&](int64_t j) {
        uint64_t lsteps = 0, lrhs = 0;
        for (uint32_t i = 0; i < nx; i++) {
            uint32_t px = x0 + i * xs, py = y0 + (uint32_t)j * ys;
            // precision 2 = x87 80-bit long double: used only to measure how sensitive a pixel of the f64
            // algorithm is to rounding (tests' "oracle-unstable" mask), never as the reference value
            PixelResult r = (rp.opts.method == METHOD_VERLET_GLSL)
                                ? ((p->precision == 1) ? render_pixel_verlet<float>(cam, rp, l, px, py)
                                                       : render_pixel_verlet<double>(cam, rp, l, px, py))
                            : (p->precision == 1)   ? render_pixel<float>(cam, rp, l, px, py)
                            : (p->precision == 2) ? render_pixel<long double>(cam, rp, l, px, py)
                                                  : render_pixel<double>(cam, rp, l, px, py);
            size_t k = (size_t)j * nx + i;
            if (rgba) for (int c = 0; c < 4; c++) rgba[4 * k + c] = r.rgba[c];
            if (xp) for (int c = 0; c < 8; c++) xp[8 * k + c] = r.xp[c];
            if (term) term[k] = r.termination;
            if (steps) steps[k] = r.steps;
            if (drift) drift[k] = r.max_drift;
            if (crossings) crossings[k] = r.crossings;
            lsteps += r.steps; lrhs += r.rhs_evals;
        }
        tsteps += lsteps; trhs += lrhs;
    });
    auto t1 = std::chrono::steady_clock::now();
    if (total_steps) *total_steps = tsteps.load();
    if (total_rhs) *total_rhs = trhs.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

// The production WebGL2 fragment shader (glsl_fragment_oracle.hpp) over the pixel lattice
// {x0 + i*xs < width} x {y0 + j*ys < y1}; precision 0 = double, 1 = float (the shader's own arithmetic).
// Outputs are dense arrays over the lattice; any may be null. Returns wall seconds.
double orc_fragment_glsl(const orc::glsl::Uniforms* u, const uint8_t* noise_r, const uint8_t* blue_r, int precision,
                         uint32_t x0, uint32_t xs, uint32_t y0, uint32_t y1, uint32_t ys, double* rgba, uint32_t* steps,
                         uint32_t* hit, uint32_t* photon, uint64_t* total_steps) {
    const uint32_t width = (uint32_t)u->resolution[0];
    const uint32_t nx = (width - x0 + xs - 1) / xs, ny = (y1 - y0 + ys - 1) / ys;
    orc::glsl::Textures tex{noise_r, blue_r};
    std::atomic<uint64_t> tsteps{0};
    auto t0 = std::chrono::steady_clock::now();
    parallel_for((int64_t)ny, 1, [&](int64_t j) {
        uint64_t ls = 0;
        for (uint32_t i = 0; i < nx; i++) {
            const uint32_t px = x0 + i * xs, py = y0 + (uint32_t)j * ys;
            const size_t k = (size_t)j * nx + i;
            double c[4]; uint32_t st, h, ph;
            if (precision == 1) { auto o = orc::glsl::Shader<float>(*u, tex).main(px, py); for (int q = 0; q < 4; q++) c[q] = o.rgba[q]; st = o.steps; h = o.hit; ph = o.photon; }
            else { auto o = orc::glsl::Shader<double>(*u, tex).main(px, py); for (int q = 0; q < 4; q++) c[q] = o.rgba[q]; st = o.steps; h = o.hit; ph = o.photon; }
            if (rgba) for (int q = 0; q < 4; q++) rgba[4 * k + q] = c[q];
            if (steps) steps[k] = st;
            if (hit) hit[k] = h;
            if (photon) photon[k] = ph;
            ls += st;
        }
        tsteps += ls;
    });
    auto t1 = std::chrono::steady_clock::now();
    if (total_steps) *total_steps = tsteps.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

// Flop census with the instrumented scalar (SURVEY §8d "figure of record").
// what: 0 = one RHS (KS), 1 = one RHS (BL), 2 = one symplectic step (KS), 3 = one RKF45 attempt (KS),
//       4 = renormalize_null (KS), 5 = hamiltonian (KS), 6 = one RK4 step (KS)
// out: [add, mul, div, sqrt, trig, pow, cmp]
void orc_flop_census(int what, double spin, const double* xp, double h, uint64_t* out7) {
    State<Counted> s;
    for (int i = 0; i < 4; i++) { s.x[i] = Counted(xp[i]); s.p[i] = Counted(xp[4 + i]); }
    int coords = (what == 1) ? BOYER_LINDQUIST : KERR_SCHILD;
    Kerr<Counted> metric(Counted(1.0), Counted(spin), coords);
    census() = FlopCensus();
    Counted err;
    switch (what) {
        case 0: case 1: (void)state_derivative(s, metric); break;
        case 2: step_symplectic(s, metric, Counted(h)); break;
        case 3: (void)rkf45_step(s, metric, Counted(h), err); break;
        case 4: renormalize_null(s, metric); break;
        case 5: (void)hamiltonian(s, metric); break;
        case 6: step_rk4(s, metric, Counted(h)); break;
    }
    FlopCensus c = census();
    out7[0] = c.add; out7[1] = c.mul; out7[2] = c.div; out7[3] = c.sqrt_; out7[4] = c.trig; out7[5] = c.pow_; out7[6] = c.cmp;
}

}  // extern "C"
