// CPU ORACLE (test infrastructure, never linked into the product) for the production WebGL2 fragment shader
// of the reference: src/shaders/blackhole/fragment.glsl.ts + chunks/{common,metric,noise,blackbody,background,disk}.ts.
//
// A statement-by-statement restatement of the GLSL in C++, templated on the scalar type (float = the shader's own
// arithmetic, double = the rounding-insensitive reference used for tight parity). Every function cites the GLSL
// lines it follows. PARITY STATUS: "parity unpinned" -- the reference holds no fixture for any pixel of this
// shader and GLSL needs a browser + GPU, which this container does not have.
//
// Conventions that GLSL leaves to the implementation and that are fixed here (and in the CUDA kernel):
//  * texture(u_noiseTex, ..) : 256x256 RGBA8 (webgl-utils.ts:259-276), LINEAR filter, REPEAT wrap, .r channel only;
//    exact bilinear weights in the working precision (real GPUs quantise them to 8 bits);
//  * texture(u_blueNoiseTex, ..) : NEAREST, REPEAT (webgl-utils.ts:281-303);
//  * normalize(v) = v / sqrt(dot(v, v)); mix(a, b, t) = a (1 - t) + b t; smoothstep as in the GLSL ES 3.00 spec;
//  * row j of the output image is gl_FragCoord.y = j + 0.5 (bottom-up, what gl.readPixels returns).
#pragma once
#include <cmath>
#include <cstdint>

namespace orc {
namespace glsl {

// chunks/common.ts:9-38 (values as webgl/renderer.ts:296-358 uploads them) + the shader manager's #defines
// (shaders/manager.ts:55-82) as feature bits. Layout shared with include/gravitas_b200.h::GvtGlslUniforms.
struct Uniforms {
    uint32_t struct_size;
    uint32_t features;
    float resolution[2];
    float time, mass, spin;
    float disk_density, disk_temp;
    float mouse[2];
    float zoom, lensing_strength, disk_size, disk_scale_height;
    int32_t max_ray_steps;
    float debug, show_redshift, show_kerr_shadow, shadow_count;
    float cam_pos[3];
    float cam_quat[4];
    float shadow_curve[128];
};
enum Feature : uint32_t {
    F_LENSING = 1u, F_DISK = 2u, F_JETS = 4u, F_STARS = 8u, F_PHOTON_GLOW = 16u, F_DOPPLER = 32u, F_REDSHIFT = 64u,
    F_LINEAR_OUTPUT = 128u, F_QUALITY_LOW = 256u
};

template <class R> struct v3 {
    R x, y, z;
};
template <class R> inline v3<R> operator+(v3<R> a, v3<R> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class R> inline v3<R> operator-(v3<R> a, v3<R> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class R> inline v3<R> operator*(v3<R> a, R s) { return {a.x * s, a.y * s, a.z * s}; }
template <class R> inline v3<R> operator*(v3<R> a, v3<R> b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
template <class R> inline v3<R> operator-(v3<R> a) { return {-a.x, -a.y, -a.z}; }
template <class R> inline R dot(v3<R> a, v3<R> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class R> inline v3<R> cross(v3<R> a, v3<R> b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
template <class R> inline R length(v3<R> a) { return std::sqrt(dot(a, a)); }
template <class R> inline v3<R> normalize(v3<R> a) { R n = length(a); return {a.x / n, a.y / n, a.z / n}; }
template <class R> inline R gmax(R a, R b) { return a < b ? b : a; }      // GLSL max(x, y): y if x < y else x
template <class R> inline R gmin(R a, R b) { return b < a ? b : a; }
template <class R> inline R clamp(R x, R lo, R hi) { return gmin(gmax(x, lo), hi); }
template <class R> inline R mix(R a, R b, R t) { return a * (R(1) - t) + b * t; }
template <class R> inline v3<R> mix(v3<R> a, v3<R> b, R t) { return {mix(a.x, b.x, t), mix(a.y, b.y, t), mix(a.z, b.z, t)}; }
template <class R> inline R smoothstep(R e0, R e1, R x) {
    R t = clamp((x - e0) / (e1 - e0), R(0), R(1));
    return t * t * (R(3) - R(2) * t);
}
template <class R> inline R sign(R x) { return x > R(0) ? R(1) : (x < R(0) ? R(-1) : R(0)); }
template <class R> inline R fract(R x) { return x - std::floor(x); }
// v.ab *= rot(t) with rot(t) = mat2(c, -s, s, c) (chunks/common.ts:46-49): row vector times matrix
template <class R> inline void rot2(R ang, R& a, R& b) {
    R s = std::sin(ang), c = std::cos(ang);
    R na = a * c - b * s, nb = a * s + b * c;
    a = na; b = nb;
}

struct Textures {
    const uint8_t* noise_r;   // 256*256, the .r channel of u_noiseTex
    const uint8_t* blue_r;    // 256*256, the .r channel of u_blueNoiseTex
};

template <class R> struct Shader {
    const Uniforms& U;
    const Textures& T;
    Shader(const Uniforms& u, const Textures& t) : U(u), T(t) {}
    static constexpr double PI_LIT = 3.14159265359;   // chunks/common.ts:40
    R PI() const { return R(PI_LIT); }
    R MAX_DIST() const { return R(10000.0); }         // physics.config.ts:60
    R MIN_STEP() const { return R(0.01); }
    R MAX_STEP() const { return R(1.2); }
    R u_time() const { return R(U.time); }

    // ---- chunks/noise.ts ---------------------------------------------------------------------------------
    R tex_noise(R u, R v) const {   // texture(u_noiseTex, vec2(u, v)).r : LINEAR, REPEAT
        R x = u * R(256) - R(0.5), y = v * R(256) - R(0.5);
        R xf = std::floor(x), yf = std::floor(y);
        R fx = x - xf, fy = y - yf;
        long long xi = (long long)xf, yi = (long long)yf;
        auto wrap = [](long long i) { return (int)(((i % 256) + 256) % 256); };
        int x0 = wrap(xi), x1 = wrap(xi + 1), y0 = wrap(yi), y1 = wrap(yi + 1);
        R t00 = R(T.noise_r[y0 * 256 + x0]) / R(255), t10 = R(T.noise_r[y0 * 256 + x1]) / R(255);
        R t01 = R(T.noise_r[y1 * 256 + x0]) / R(255), t11 = R(T.noise_r[y1 * 256 + x1]) / R(255);
        return mix(mix(t00, t10, fx), mix(t01, t11, fx), fy);
    }
    R hash(v3<R> p) const {   // noise.ts:3-8
        R ux = p.x + p.z * R(37.0), uy = p.y + p.z * R(37.0);
        return tex_noise((ux + R(0.5)) / R(256.0), (uy + R(0.5)) / R(256.0));
    }
    R noise(v3<R> p) const {   // noise.ts:11-19
        v3<R> i = {std::floor(p.x), std::floor(p.y), std::floor(p.z)};
        v3<R> f = {fract(p.x), fract(p.y), fract(p.z)};
        f = {f.x * f.x * (R(3) - R(2) * f.x), f.y * f.y * (R(3) - R(2) * f.y), f.z * f.z * (R(3) - R(2) * f.z)};
        auto h = [&](R dx, R dy, R dz) { return hash(v3<R>{i.x + dx, i.y + dy, i.z + dz}); };
        return mix(mix(mix(h(0, 0, 0), h(1, 0, 0), f.x), mix(h(0, 1, 0), h(1, 1, 0), f.x), f.y),
                   mix(mix(h(0, 0, 1), h(1, 0, 1), f.x), mix(h(0, 1, 1), h(1, 1, 1), f.x), f.y), f.z);
    }
    R fbm(v3<R> p) const {   // noise.ts:22-31
        R f(0), amp(0.5);
        for (int i = 0; i < 4; i++) { f += amp * noise(p); p = p * R(2.0); amp *= R(0.5); }
        return f;
    }

    // ---- chunks/blackbody.ts -----------------------------------------------------------------------------
    v3<R> blackbody(R temp) const {   // blackbody.ts:9-35
        R t = gmax(temp, R(1.0)) / R(100.0);
        R r, g, b;
        if (t <= R(66.0)) {
            r = R(255.0);
            g = R(99.4708025861) * std::log(t) - R(161.1195681661);
            if (t <= R(19.0)) b = R(0.0);
            else b = R(138.5177312231) * std::log(t - R(10.0)) - R(305.0447927307);
        } else {
            r = R(329.698727446) * std::pow(t - R(60.0), R(-0.1332047592));
            g = R(288.1221695283) * std::pow(t - R(60.0), R(-0.0755148492));
            b = R(255.0);
        }
        v3<R> s = {r / R(255.0), g / R(255.0), b / R(255.0)};
        return {std::pow(gmax(s.x, R(0)), R(2.2)), std::pow(gmax(s.y, R(0)), R(2.2)), std::pow(gmax(s.z, R(0)), R(2.2))};
    }
    v3<R> starColor(R bv) const {   // blackbody.ts:38-47
        R t = clamp(bv, R(-0.4), R(2.0));
        if (t < R(0.0)) return {R(0.6), R(0.7), R(1.0)};
        if (t < R(0.3)) return {R(0.85), R(0.88), R(1.0)};
        if (t < R(0.6)) return {R(1.0), R(0.96), R(0.9)};
        if (t < R(1.0)) return {R(1.0), R(0.85), R(0.6)};
        return {R(1.0), R(0.6), R(0.4)};
    }

    // ---- chunks/background.ts:3-30 -----------------------------------------------------------------------
    v3<R> starfield(v3<R> dir) const {
        v3<R> stars = {R(0), R(0), R(0)};
        v3<R> cell = {std::floor(dir.x * R(200.0)), std::floor(dir.y * R(200.0)), std::floor(dir.z * R(200.0))};
        R starNoise = hash(cell);
        if (starNoise > R(0.998)) {
            R brightness = std::pow(starNoise, R(10.0)) * R(2.0);
            R bv = hash(v3<R>{cell.x + R(127.1), cell.y + R(127.1), cell.z + R(127.1)}) * R(2.4) - R(0.4);
            R tw = R(0.85) + R(0.15) * std::sin(u_time() * (R(3.0) + hash(v3<R>{cell.x + R(73.7), cell.y + R(73.7), cell.z + R(73.7)}) * R(2.0)));
            stars = starColor(bv) * brightness * tw;
        }
        cell = {std::floor(dir.x * R(500.0)), std::floor(dir.y * R(500.0)), std::floor(dir.z * R(500.0))};
        starNoise = hash(cell);
        if (starNoise > R(0.996)) {
            R brightness = std::pow(starNoise, R(20.0)) * R(1.5);
            R bv = hash(v3<R>{cell.x + R(217.3), cell.y + R(217.3), cell.z + R(217.3)}) * R(2.4) - R(0.4);
            stars = stars + starColor(bv) * brightness;
        }
        R tt = u_time() * R(0.01);
        R nebula = fbm(v3<R>{dir.x * R(2.0) + tt, dir.y * R(2.0) + tt, dir.z * R(2.0) + tt}) * R(0.03);
        R ln = std::fabs(nebula);
        stars = stars + (v3<R>{nebula * R(0.2), nebula * R(0.3), nebula * R(0.5)} + v3<R>{R(0.05) * ln, R(0.02) * ln, R(0.05) * ln});
        return stars;
    }

    // ---- chunks/metric.ts --------------------------------------------------------------------------------
    R kerr_horizon(R M, R a) const { return M + std::sqrt(gmax(R(0.0), M * M - a * a)); }   // :13-15
    R kerr_isco(R M, R a) const {                                                             // :18-29
        R rs = a / M;
        R absS = std::fabs(clamp(rs, R(-0.9999), R(0.9999)));
        R z1 = R(1.0) + std::pow(R(1.0) - absS * absS, R(1.0 / 3.0)) *
                            (std::pow(R(1.0) + absS, R(1.0 / 3.0)) + std::pow(R(1.0) - absS, R(1.0 / 3.0)));
        R z2 = std::sqrt(R(3.0) * absS * absS + z1 * z1);
        R signOfA = sign(a);
        if (signOfA == R(0.0)) signOfA = R(1.0);
        return M * (R(3.0) + z2 - signOfA * std::sqrt((R(3.0) - z1) * (R(3.0) + z1 + R(2.0) * z2)));
    }
    R kerr_photon_sphere(R M, R a) const {                                                    // :32-37
        R a_star = clamp(a / M, R(-0.9999), R(0.9999));
        R arg = clamp(-a_star, R(-1.0), R(1.0));
        R theta = R(2.0 / 3.0) * std::acos(arg);
        return R(2.0) * M * (R(1.0) + std::cos(theta));
    }
    R kerr_ergosphere(R M, R a, R cosTheta) const {                                           // :47-49
        return M + std::sqrt(gmax(R(0.0), M * M - a * a * cosTheta * cosTheta));
    }
    void kerr_geodesic_accel(v3<R> p, v3<R> v, R M, R a, v3<R>& accel, R& omega) const {      // :96-149
        R a2 = a * a;
        R rho2 = dot(p, p);
        R diff = rho2 - a2;
        R disc = diff * diff + R(4.0) * a2 * p.y * p.y;
        R r2 = R(0.5) * (diff + std::sqrt(gmax(R(0.0), disc)));
        R r_k = std::sqrt(gmax(R(1e-8), r2));
        R sigma = r2 + a2 * (p.y * p.y / gmax(R(1e-8), r2));
        v3<R> L = cross(p, v);
        R Ly = L.y;
        R Ly_eff = Ly - a;
        R L2_eff = Ly_eff * Ly_eff + (dot(L, L) - Ly * Ly);
        R r_inv = R(1.0) / r_k;
        R r2_inv = r_inv * r_inv;
        R r4_inv = r2_inv * r2_inv;
        R sigma_ratio = r2 / gmax(R(1e-8), sigma);
        v3<R> r_hat = -normalize(p);
        accel = r_hat * (M * r2_inv * sigma_ratio + R(3.0) * M * gmax(R(0.0), L2_eff) * r4_inv * sigma_ratio);
        R r3_p_a2r = r_k * r2 + a2 * r_k;
        R drag = R(2.0) * M * a / gmax(R(1e-8), r3_p_a2r);
        accel = accel + cross(v3<R>{R(0), R(1), R(0)}, v) * drag;
        omega = R(2.0) * M * a / gmax(R(1e-8), r3_p_a2r);
    }

    // ---- chunks/disk.ts:16-118 ---------------------------------------------------------------------------
    void sample_accretion_disk(v3<R> p, v3<R> p_prev, v3<R> v, R isco, R M, R a, R dt, v3<R>& accC, R& accA) const {
        if (!(R(U.show_redshift) < R(0.5))) return;
        bool crossedEquator = (p_prev.y * p.y < R(0.0));
        v3<R> sampleP = p;
        if (crossedEquator) {
            R t = std::fabs(p_prev.y) / gmax(R(0.0001), std::fabs(p_prev.y) + std::fabs(p.y));
            sampleP = mix(p_prev, p, t);
        }
        R sampleR = length(sampleP);
        R effectiveScaleHeight = gmin(R(U.disk_scale_height), R(0.450));
        R diskHeight = sampleR * effectiveScaleHeight;
        R diskInner = isco;
        R diskOuter = gmax(M * R(U.disk_size), diskInner * R(1.1));
        if (!((std::fabs(sampleP.y) < diskHeight || crossedEquator) && sampleR > diskInner && sampleR < diskOuter)) return;
        R u_spin = R(U.spin);
        R sqrt_M_phase = std::sqrt(M);
        R signSpinPhase = sign(u_spin + R(1e-8));
        R OmegaPhase = (signSpinPhase * sqrt_M_phase) / (sampleR * std::sqrt(sampleR) + a * sqrt_M_phase);
        R rotAngle = OmegaPhase * u_time() * R(0.12) * R(10.0);
        v3<R> noiseP = sampleP;
        {   // noiseP.xz *= mat2(cos, -sin, sin, cos)
            R c = std::cos(rotAngle), s = std::sin(rotAngle);
            R nx = noiseP.x * c - noiseP.z * s, nz = noiseP.x * s + noiseP.z * c;
            noiseP.x = nx; noiseP.z = nz;
        }
        noiseP = noiseP * R(0.75);
        R turbulence = noise(noiseP) * R(0.5) + noise(noiseP * R(2.5)) * R(0.25);
        R samplesDiskHeight = sampleR * effectiveScaleHeight;
        R heightFalloff = std::exp(-std::fabs(sampleP.y) / gmax(R(0.001), samplesDiskHeight * R(0.25)));
        R radialFalloff = smoothstep(diskOuter, diskInner, sampleR);
        R baseDensity = turbulence * heightFalloff * radialFalloff;
        if (!(baseDensity > R(0.001))) return;
        R r2 = sampleR * sampleR;
        R sqrt_M = std::sqrt(M);
        R signSpin = sign(u_spin + R(1e-8));
        R Omega = (signSpin * sqrt_M) / (sampleR * std::sqrt(sampleR) + a * sqrt_M);
        R g_tt = -(R(1.0) - R(2.0) * M / sampleR);
        R g_tphi = R(-2.0) * M * a / sampleR;
        R g_phiphi = r2 + a * a + R(2.0) * M * a * a / sampleR;
        R u_t_sq = -(g_tt + R(2.0) * Omega * g_tphi + Omega * Omega * g_phiphi);
        R u_t = R(1.0) / std::sqrt(gmax(R(1e-6), u_t_sq));
        R L_photon = p.z * v.x - p.x * v.z;
        R delta = R(1.0) / gmax(R(0.01), u_t * (R(1.0) - Omega * L_photon));
        R beaming = (U.features & F_DOPPLER) ? gmax(R(0.01), std::pow(delta, R(3.5))) : R(1.0);
        R isco_r = clamp(isco / sampleR, R(0.0), R(1.0));
        R nt_factor = gmax(R(0.0), R(1.0) - std::sqrt(isco_r));
        R radialTempGradient = std::pow(isco_r, R(0.75)) * std::pow(nt_factor, R(0.25));
        R temperature = R(U.disk_temp) * radialTempGradient * delta;
        v3<R> diskColor = blackbody(temperature) * beaming;
        R density = baseDensity * R(U.disk_density) * R(0.12) * dt;
        accC = accC + diskColor * density * (R(1.0) - accA);
        accA += density;
    }
    // ---- chunks/disk.ts:120-155 --------------------------------------------------------------------------
    void sample_relativistic_jets(v3<R> p, v3<R> v, R rh, R dt, v3<R>& accC, R& accA) const {
        R jetVerticalPos = std::fabs(p.y);
        if (!(jetVerticalPos > rh * R(1.8) && jetVerticalPos < MAX_DIST() * R(0.8))) return;
        R jetRadialDist = std::sqrt(p.x * p.x + p.z * p.z);
        R jetWidth = R(1.0) + jetVerticalPos * R(0.15);
        if (!(jetRadialDist < jetWidth * R(2.0))) return;
        R radialFalloff = std::exp(-(jetRadialDist * jetRadialDist) / (jetWidth * R(0.5)));
        R lengthFalloff = std::exp(-jetVerticalPos * R(0.05));
        R flowCombined = p.y * R(2.0) - u_time() * R(8.0);
        v3<R> uvJet = {p.x, flowCombined, p.z};
        R noiseVal = noise(uvJet * R(0.5)) * R(0.6) + noise(uvJet * R(1.5)) * R(0.4);
        R jetDensity = radialFalloff * lengthFalloff * gmax(R(0.0), noiseVal - R(0.2));
        if (!(jetDensity > R(0.001))) return;
        R jetVel = R(0.92) * sign(p.y);
        v3<R> jetVelVec = {R(0.0), jetVel, R(0.0)};
        R cosThetaJet = dot(normalize(jetVelVec), -v);
        R betaJet = std::fabs(jetVel);
        R gammaJet = R(1.0) / std::sqrt(R(1.0) - betaJet * betaJet);
        R deltaJet = R(1.0) / (gammaJet * (R(1.0) - betaJet * cosThetaJet));
        R beamingJet = std::pow(deltaJet, R(3.5));
        v3<R> baseJetColor = {R(0.4), R(0.7), R(1.0)};
        v3<R> jetEmission = baseJetColor * jetDensity * R(0.05) * beamingJet * dt;
        accC = accC + jetEmission * (R(1.0) - accA);
        accA += jetDensity * R(0.05) * dt;
    }

    v3<R> aces(v3<R> c) const {   // chunks/common.ts:52-59
        auto f = [](R x) { return clamp((x * (R(2.51) * x + R(0.03))) / (x * (R(2.43) * x + R(0.59)) + R(0.14)), R(0.0), R(1.0)); };
        return {f(c.x), f(c.y), f(c.z)};
    }

    struct Out {
        double rgba[4];
        uint32_t steps;
        uint32_t hit;
        uint32_t photon;
    };

    // ---- fragment.glsl.ts:41-333 main() ------------------------------------------------------------------
    Out main(uint32_t px, uint32_t py) const {
        Out o{};
        auto emit = [&](v3<R> c) { o.rgba[0] = (double)c.x; o.rgba[1] = (double)c.y; o.rgba[2] = (double)c.z; o.rgba[3] = 1.0; };
        const R resx = R(U.resolution[0]), resy = R(U.resolution[1]);
        const R fcx = R((double)px) + R(0.5), fcy = R((double)py) + R(0.5);   // gl_FragCoord.xy
        R minRes = gmin(resx, resy);
        R uvx = (fcx - R(0.5) * resx) / minRes, uvy = (fcy - R(0.5) * resy) / minRes;
        if (U.debug > 0.5f) { emit({uvx + R(0.5), uvy + R(0.5), R(0.0)}); return o; }
        v3<R> ro, rd;
        v3<R> camPos = {R(U.cam_pos[0]), R(U.cam_pos[1]), R(U.cam_pos[2])};
        if (length(camPos) > R(0.001)) {
            ro = camPos;
            v3<R> d = normalize(v3<R>{uvx, uvy, R(1.2)});
            v3<R> q = {R(U.cam_quat[0]), R(U.cam_quat[1]), R(U.cam_quat[2])};
            R qw = R(U.cam_quat[3]);
            rd = d + cross(q, cross(q, d) + d * qw) * R(2.0);   // qrot, common.ts:75-77
        } else {
            ro = {R(0.0), R(0.0), -R(U.zoom)};
            rd = normalize(v3<R>{uvx, uvy, R(1.5)});
            R ax = (R(U.mouse[1]) - R(0.5)) * PI();
            R ay = (R(U.mouse[0]) - R(0.5)) * PI() * R(2.0);
            rot2(ax, ro.y, ro.z); rot2(ax, rd.y, rd.z);
            rot2(ay, ro.x, ro.z); rot2(ay, rd.x, rd.z);
        }
        R M = R(U.mass);
        R rs = M * R(2.0);
        R a = R(U.spin) * M;
        R rh = kerr_horizon(M, a);
        R rph = kerr_photon_sphere(M, a);
        R isco = kerr_isco(M, a);
        R absA = std::fabs(R(U.spin));

        if (U.features & F_QUALITY_LOW) {   // fragment.glsl.ts:76-87
            v3<R> bg = starfield(rd);
            R d = length(cross(ro, rd));
            R shadow = smoothstep(rh * R(1.2), rh * R(0.9), d);
            R glow = std::exp(-std::fabs(d - rph) * R(12.0)) * R(0.8);
            v3<R> glowCol = v3<R>{R(0.3), R(0.6), R(1.0)} * glow;
            R diskMask = smoothstep(isco * R(2.0), isco * R(1.0), d) * (R(1.0) - smoothstep(isco * R(1.0), isco * R(0.8), d));
            v3<R> diskCol = v3<R>{R(1.0), R(0.7), R(0.3)} * diskMask * R(0.6);
            v3<R> col = bg * (R(1.0) - shadow) + glowCol + diskCol;
            emit({std::pow(col.x, R(0.4545)), std::pow(col.y, R(0.4545)), std::pow(col.z, R(0.4545))});
            return o;
        }

        v3<R> p = ro, v = rd;
        if (length(ro) < rh * R(1.5)) { ro = normalize(ro) * rh * R(1.5); p = ro; }   // :94-97
        v3<R> accC = {R(0), R(0), R(0)};
        R accA(0.0);
        bool hitHorizon = false;
        R maxRedshift(0.0);
        R bNoise = R(T.blue_r[(py % 256u) * 256u + (px % 256u)]) / R(255);              // :105-106 NEAREST at a texel centre
        R dt_init = MIN_STEP();
        p = p + v * bNoise * dt_init;
        int photonCrossings = 0;
        R prevY = p.y;
        R impactParam = length(cross(ro, rd));
        bool redshiftInitialized = false;
        int maxSteps = (int)gmin(R((double)U.max_ray_steps), R(500.0));
        v3<R> p_prev = p;
        if (impactParam < rh * R(0.9)) hitHorizon = true;                                // :124-126
        const R lens = R(U.lensing_strength);
        for (int i = 0; i < maxSteps; i++) {
            p_prev = p;
            R r = length(p);
            if (r < rh * R(1.15)) { hitHorizon = true; break; }
            if (r > MAX_DIST()) break;
            R distFactor = R(1.0) + r * R(0.05);
            R dt = clamp((r - rh) * R(0.1) * distFactor, MIN_STEP(), MAX_STEP() * distFactor);
            if (r > R(30.0)) {
                R farBoost = (r - R(30.0)) * R(0.08);
                dt = gmax(dt, MIN_STEP() + farBoost);
                dt = gmin(dt, MAX_STEP() * R(2.5));
            }
            R sphereProx = std::fabs(r - rph);
            dt = gmin(dt, MIN_STEP() + sphereProx * R(0.15));
            R hRefinement = smoothstep(R(0.2), R(0.0), std::fabs(p.y));
            R currentDt = dt * (R(1.0) - hRefinement * R(0.7));
            v3<R> accel = {R(0), R(0), R(0)};
            R omega(0.0);
            if (U.features & F_LENSING) {
                kerr_geodesic_accel(p, v, M, a, accel, omega);
                accel = accel * lens;
                rot2(omega * currentDt, v.x, v.z);
            }
            p = p + (v * currentDt + accel * R(0.5) * currentDt * currentDt);
            R r_new = length(p);
            if (U.features & F_LENSING) {
                if (accA < R(0.95)) {
                    v3<R> accel_new; R om2;
                    kerr_geodesic_accel(p, v, M, a, accel_new, om2);
                    accel_new = accel_new * lens;
                    v = v + (accel + accel_new) * R(0.5) * currentDt;
                }
            }
            v = normalize(v);
            if (prevY * p.y < R(0.0) && r_new < rph * R(2.0) && r_new > rh) photonCrossings = photonCrossings + 1 < 3 ? photonCrossings + 1 : 3;
            if (U.show_redshift > 0.5f) {
                R potential = std::sqrt(gmax(R(0.0), R(1.0) - rs / r_new));
                if (!redshiftInitialized) { maxRedshift = potential; redshiftInitialized = true; }
                else maxRedshift = gmin(maxRedshift, potential);
            }
            prevY = p.y;
            o.steps++;
            if (U.features & F_DISK) {
                sample_accretion_disk(p, p_prev, v, isco, M, a, currentDt, accC, accA);
                if (accA > R(0.99)) break;
            }
            if (U.features & F_JETS) sample_relativistic_jets(p, v, rh, dt, accC, accA);
        }
        o.hit = hitHorizon ? 1u : 0u;
        o.photon = (uint32_t)photonCrossings;
        if ((U.features & F_REDSHIFT) && U.show_redshift > 0.5f) {   // :223-236
            R val = maxRedshift;
            if (hitHorizon) val = R(0.0);
            v3<R> heat = mix(v3<R>{R(0), R(0), R(0)}, v3<R>{R(1), R(0), R(0)}, smoothstep(R(0.0), R(0.3), val));
            heat = mix(heat, v3<R>{R(1), R(1), R(0)}, smoothstep(R(0.3), R(0.7), val));
            heat = mix(heat, v3<R>{R(0), R(0), R(1)}, smoothstep(R(0.7), R(1.0), val));
            emit(heat);
            return o;
        }
        v3<R> background = {R(0), R(0), R(0)};
        if (U.features & F_STARS) background = starfield(v);
        v3<R> photonColor = {R(0), R(0), R(0)};
        if ((U.features & F_PHOTON_GLOW) && !hitHorizon) {            // :246-258
            R distToPhotonRing = std::fabs(length(p) - rph);
            R directRing = std::exp(-distToPhotonRing * R(40.0)) * R(1.8) * lens;
            R higherOrderRing(0.0);
            if (photonCrossings > 0) {
                R ringSharpness = R(60.0) + R((double)photonCrossings) * R(30.0);
                R ringBrightness = std::exp(-R((double)photonCrossings) * R(1.0)) * R(1.2);
                higherOrderRing = std::exp(-distToPhotonRing * ringSharpness) * ringBrightness * lens;
            }
            R s = directRing + higherOrderRing;
            photonColor = {s, s, s};
        }
        v3<R> ergoColor = {R(0), R(0), R(0)};
        if (absA > R(0.1) && !hitHorizon) {                            // :262-268
            R rFinal = length(p);
            R cosTheta = p.y / gmax(rFinal, R(0.001));
            R r_ergo = kerr_ergosphere(M, a, cosTheta);
            R ergoGlow = std::exp(-std::fabs(rFinal - r_ergo) * R(20.0)) * R(0.35) * absA;
            ergoColor = v3<R>{R(0.3), R(0.35), R(0.9)} * ergoGlow;
        }
        if (hitHorizon) background = {R(0), R(0), R(0)};
        R om = R(1.0) - accA;
        v3<R> finalColor = background * om + accC + photonColor * om + ergoColor * om;   // :276
        if (U.show_kerr_shadow > 0.5f) {                                // :279-322
            v3<R> spin_axis = {R(0), R(1), R(0)};
            v3<R> cam_dir = normalize(ro);
            v3<R> sky_right = normalize(cross(spin_axis, cam_dir));
            v3<R> sky_up = cross(cam_dir, sky_right);
            v3<R> impact_vec = cross(cam_dir, rd) * length(ro);
            R al = -dot(impact_vec, sky_up), be = dot(impact_vec, sky_right);
            R minDist(1e10);
            int count = (int)U.shadow_count;
            auto seg = [&](R p1x, R p1y, R p2x, R p2y) {
                R pax = al - p1x, pay = be - p1y, bax = p2x - p1x, bay = p2y - p1y;
                R h = clamp((pax * bax + pay * bay) / (bax * bax + bay * bay), R(0.0), R(1.0));
                R dx = pax - bax * h, dy = pay - bay * h;
                minDist = gmin(minDist, std::sqrt(dx * dx + dy * dy));
            };
            for (int j = 0; j < 63; j++) {
                if (j >= count - 1) break;
                seg(R(U.shadow_curve[2 * j]), R(U.shadow_curve[2 * j + 1]), R(U.shadow_curve[2 * j + 2]), R(U.shadow_curve[2 * j + 3]));
            }
            if (count > 2) {
                int l = (count < 64 ? count : 64) - 1;
                seg(R(U.shadow_curve[2 * l]), R(U.shadow_curve[2 * l + 1]), R(U.shadow_curve[0]), R(U.shadow_curve[1]));
            }
            R thickness = M * R(0.045);
            if (minDist < thickness) {
                R edge = smoothstep(thickness, thickness * R(0.5), minDist);
                finalColor = mix(finalColor, v3<R>{R(0), R(1), R(0)}, R(1.0) * edge);
            }
        }
        if (!(U.features & F_LINEAR_OUTPUT)) {                          // :325-328
            finalColor = aces(finalColor);
            finalColor = {std::pow(gmax(finalColor.x, R(0.0)), R(0.4545)), std::pow(gmax(finalColor.y, R(0.0)), R(0.4545)),
                          std::pow(gmax(finalColor.z, R(0.0)), R(0.4545))};
        }
        emit(finalColor);
        return o;
    }
};

}  // namespace glsl
}  // namespace orc
