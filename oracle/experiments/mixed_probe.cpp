// mixed_probe.cpp — CPU experiment (test infrastructure, not product): how much of the implicit-midpoint step
// (geodesic/integrator.rs:209-226) can be evaluated in f32 while the STATE stays f64, before the composite RGBA of
// the headline frame (config 3: 3840x2160x512, a*=0.999) leaves the 1e-6 relative tolerance of north_star?
//   mode 0: all three RHS evaluations in f64 (the oracle itself)       mode 1: fixed-point iteration 1 in f32
//   mode 2: iterations 1 and 2 in f32, final evaluation in f64         mode 3: all three in f32, f64 accumulation
// Usage: mixed_probe <camera88.f32> <stride> [spin [mode r_switch]...]   -> one line per mode
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../gravitas_oracle.hpp"
using namespace orc;

static State<double> rhs_mixed(const State<double>& s, const Kerr<double>& m64, const Kerr<float>& m32, bool f32) {
    if (!f32) return state_derivative(s, m64);
    State<float> sf;
    for (int i = 0; i < 4; i++) { sf.x[i] = (float)s.x[i]; sf.p[i] = (float)s.p[i]; }
    State<float> df = state_derivative(sf, m32);
    State<double> d;
    for (int i = 0; i < 4; i++) { d.x[i] = (double)df.x[i]; d.p[i] = (double)df.p[i]; }
    return d;
}

static void step_mixed(State<double>& s, const Kerr<double>& m64, const Kerr<float>& m32, double h, int mode) {
    State<double> s_mid = s;
    for (int it = 0; it < 2; it++) {
        State<double> d = rhs_mixed(s_mid, m64, m32, mode >= 3 || (mode == 2) || (mode == 1 && it == 0));
        for (int i = 0; i < 4; i++) {
            double nx = s.x[i] + d.x[i] * h, np = s.p[i] + d.p[i] * h;
            s_mid.x[i] = 0.5 * (s.x[i] + nx);
            s_mid.p[i] = 0.5 * (s.p[i] + np);
        }
    }
    State<double> d = rhs_mixed(s_mid, m64, m32, mode >= 3);
    for (int i = 0; i < 4; i++) { s.x[i] += d.x[i] * h; s.p[i] += d.p[i] * h; }
}

static double g_rswitch = 1e30;
static std::atomic<uint64_t> g_f32steps{0}, g_steps{0};
struct Px { double rgb[3]; uint32_t term, steps; };

static Px render_mixed(const CameraUniforms& cam, const RenderParams& rp, const Luts& luts, uint32_t px, uint32_t py, int mode) {
    Kerr<double> m64(rp.mass, rp.spin, rp.coords);
    Kerr<float> m32((float)rp.mass, (float)rp.spin, rp.coords);
    State<double> s = camera_ray<double>(cam, rp, px, py);
    DiskHook<double> hook;
    hook.luts = &luts; hook.mass = rp.mass; hook.spin = m64.spin; hook.r_in = m64.isco(true); hook.r_out = rp.disk_r_out;
    const double horizon = m64.event_horizon();
    renormalize_null(s, m64);
    uint32_t steps = 0, term = TERM_MAXSTEPS;
    for (uint64_t it = 0; it < rp.opts.max_steps; it++) {
        uint32_t t = check_termination(s, horizon, rp.opts.escape_radius);
        if (t != TERM_NONE) { term = t; break; }
        State<double> prev = s;
        {
            // modes >= 10: radius-adaptive. Beyond r_switch the two fixed-point (predictor) evaluations run in f32
            // (their error reaches the step only through h J / 2), inside it everything is f64.
            int m = mode;
            if (mode >= 10) { m = (s.x[1] > g_rswitch) ? (mode - 10) : 0; if (m) g_f32steps++; g_steps++; }
            step_mixed(s, m64, m32, wgsl_step_rule<double>(s.x[1], horizon), m);
        }
        if (steps % rp.opts.renormalize_interval == 0) renormalize_null(s, m64);
        steps++;
        if (hook(prev, s)) { term = TERM_DISK; break; }
    }
    Px o;
    for (int c = 0; c < 3; c++) o.rgb[c] = hook.color[c];
    o.term = term; o.steps = steps;
    return o;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s camera88.f32 stride [spin]\n", argv[0]); return 2; }
    CameraUniforms cam;
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(cam.f, 4, 88, f) != 88) { fprintf(stderr, "cannot read camera\n"); return 2; }
    fclose(f);
    const uint32_t stride = (uint32_t)atoi(argv[2]);
    RenderParams rp;
    rp.mass = 1.0; rp.spin = argc > 3 ? atof(argv[3]) : (double)0.999f;
    rp.width = 3840; rp.height = 2160; rp.coords = KERR_SCHILD;
    rp.opts.method = METHOD_SYMPLECTIC; rp.opts.step_rule = 1; rp.opts.max_steps = 512;
    std::vector<float> spec(256 * 32 * 4), td(512);
    spectrum::generate_blackbody_lut(256, 32, 1e7, spec.data());
    Kerr<double> bh(rp.mass, rp.spin, BOYER_LINDQUIST);
    disk::generate_temperature_lut(bh, 512, td.data());
    Luts l;
    l.spectrum = spec.data(); l.spec_w = 256; l.spec_h = 32; l.tdisk = td.data(); l.tdisk_n = 512;
    l.tdisk_rin = bh.isco(true); l.tdisk_rout = 50.0 * rp.mass;
    const uint32_t nx = (rp.width + stride - 1) / stride, ny = (rp.height + stride - 1) / stride;
    std::vector<int> modes = {0, 1, 2, 3};
    std::vector<double> rsw = {0, 0, 0, 0};
    for (int a = 4; a + 1 < argc; a += 2) { modes.push_back(10 + atoi(argv[a])); rsw.push_back(atof(argv[a + 1])); }
    std::vector<std::vector<Px>> res(modes.size());
    for (size_t mi = 0; mi < modes.size(); mi++) {
        const int mode = modes[mi];
        g_rswitch = rsw[mi]; g_f32steps = 0; g_steps = 0;
        res[mi].resize((size_t)nx * ny);
        std::atomic<uint32_t> next{0};
        auto work = [&]() {
            for (;;) {
                uint32_t j = next.fetch_add(1);
                if (j >= ny) return;
                for (uint32_t i = 0; i < nx; i++)
                    res[mi][(size_t)j * nx + i] = render_mixed(cam, rp, l, stride / 2 + i * stride, stride / 2 + j * stride, mode);
            }
        };
        std::vector<std::thread> th;
        unsigned nt = std::thread::hardware_concurrency();
        for (unsigned t = 1; t < nt; t++) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
        double peak = 0;
        for (auto& p : res[0]) for (int c = 0; c < 3; c++) peak = std::max(peak, p.rgb[c]);
        size_t bad6 = 0, bad5 = 0, bad4 = 0, census = 0, lit = 0;
        double worst = 0;
        std::vector<double> errs;
        for (size_t k = 0; k < res[0].size(); k++) {
            const Px &a = res[0][k], &b = res[mi][k];
            if (a.term != b.term || a.steps != b.steps) { census++; if (getenv("PROBE_VERBOSE")) printf("  px %zu (%zu,%zu): term %u/%u steps %u/%u rgb %.3e/%.3e\n", k, k % nx, k / nx, a.term, b.term, a.steps, b.steps, a.rgb[0], b.rgb[0]); }
            double e = 0;
            for (int c = 0; c < 3; c++) e = std::max(e, std::fabs(a.rgb[c] - b.rgb[c]) / std::max(std::fabs(a.rgb[c]), 1e-3 * peak));
            if (a.rgb[0] + a.rgb[1] + a.rgb[2] > 0) { lit++; errs.push_back(e); }
            worst = std::max(worst, e);
            bad6 += e > 1e-6; bad5 += e > 1e-5; bad4 += e > 1e-4;
        }
        std::sort(errs.begin(), errs.end());
        if (mode >= 10) printf("[r_switch %.1f, %.1f%% of steps with f32 predictors] ", g_rswitch, 100.0 * g_f32steps / std::max<uint64_t>(g_steps, 1));
        printf("mode %d: %zu px (%zu lit)  >1e-6: %zu (%.3f%%)  >1e-5: %zu  >1e-4: %zu  census-diff: %zu  max %.3e  median(lit) %.3e  p99(lit) %.3e\n",
               mode, res[0].size(), lit, bad6, 100.0 * bad6 / res[0].size(), bad5, bad4, census, worst,
               errs.empty() ? 0.0 : errs[errs.size() / 2], errs.empty() ? 0.0 : errs[errs.size() * 99 / 100]);
        fflush(stdout);
    }
    return 0;
}
