// gvt_fragment.cu — the reference's production WebGL2 fragment shader as one fused sm_100a kernel.
//
// What it computes, per pixel (src/shaders/blackhole/fragment.glsl.ts:41-333 and its chunks):
//   camera ray (quaternion camera or the mouse/zoom orbit camera) -> blue-noise dithered start -> Cartesian
//   Velocity-Verlet march on the pseudo-Kerr acceleration field of chunks/metric.ts:96-149 with the shader's own
//   step-size heuristics -> per step: photon-ring crossing count, redshift potential, volumetric accretion-disk
//   emission (chunks/disk.ts:16-118: turbulence from the noise texture, Page-Thorne-style kinematics, blackbody
//   colour, Doppler beaming), relativistic jets (disk.ts:120-155) -> starfield + nebula (chunks/background.ts),
//   photon-ring and ergosphere glow, Kerr-shadow guide polyline, ACES + gamma (chunks/common.ts:52-59).
//
// B200 mapping: persistent CTAs, one warp per 8x4 pixel tile pulled from an atomic queue (march lengths differ by
// 10x between sky, disk and shadow pixels); the ray state lives in registers; the 64 KB red channel of the 256x256
// noise texture -- the only texture the inner loop samples (8 taps per noise() call, 2 noise() calls per disk
// sample) -- is staged once per CTA into shared memory by TMA (cp.async.bulk + mbarrier) and sampled with plain
// byte loads; uniforms and the 64-point shadow curve sit in the kernel parameter bank; one coalesced float4 store
// per pixel (plus the optional host / NVLink-peer copies of the same pixel). Templated on the scalar type: float is
// the shader's own arithmetic, double exists for tight parity against the f64 instantiation of the oracle.
//
// This file is compiled twice (Makefile): once as is (IEEE division / sqrt, CUDA's 1-2 ulp sinf/expf/powf/logf) ->
// launch_fragment_glsl, the parity instantiations; and once with `-use_fast_math -DGVT_FRAGMENT_FAST` ->
// launch_fragment_glsl_fast, float only, where the same source maps onto MUFU (rcp/rsq/sqrt/ex2/lg2/sin/cos) the
// way a GLSL compiler maps the shader: no slow-path branches, a third of the instructions.
#include "gvt_internal.h"
#include <cuda_fp16.h>

namespace gvt {
#ifdef GVT_FRAGMENT_FAST
namespace fragment_fast {
#else
namespace fragment_precise {
#endif

namespace {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// frame targets are RGBA32F or RGBA16F (the reference's texture format, reprojection.ts:120-140)
__device__ __forceinline__ void frag_store_px(float4* base, size_t idx, const float4& v, bool f16) {
    if (f16) {
        const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&lo); o.y = *reinterpret_cast<const uint32_t*>(&hi);
        reinterpret_cast<uint2*>(base)[idx] = o;
    } else {
        base[idx] = v;
    }
}

template <class R> struct GM;
template <> struct GM<float> {
    static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float sin_(float x) { return sinf(x); }
    static __device__ __forceinline__ float cos_(float x) { return cosf(x); }
    static __device__ __forceinline__ void sincos_(float x, float* s, float* c) { sincosf(x, s, c); }
    static __device__ __forceinline__ float exp_(float x) { return expf(x); }
    static __device__ __forceinline__ float log_(float x) { return logf(x); }
    static __device__ __forceinline__ float pow_(float x, float y) { return powf(x, y); }
    static __device__ __forceinline__ float floor_(float x) { return floorf(x); }
    static __device__ __forceinline__ float abs_(float x) { return fabsf(x); }
    static __device__ __forceinline__ float acos_(float x) { return acosf(x); }
};
template <> struct GM<double> {
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
    static __device__ __forceinline__ double sin_(double x) { return sin(x); }
    static __device__ __forceinline__ double cos_(double x) { return cos(x); }
    static __device__ __forceinline__ void sincos_(double x, double* s, double* c) { sincos(x, s, c); }
    static __device__ __forceinline__ double exp_(double x) { return exp(x); }
    static __device__ __forceinline__ double log_(double x) { return log(x); }
    static __device__ __forceinline__ double pow_(double x, double y) { return pow(x, y); }
    static __device__ __forceinline__ double floor_(double x) { return floor(x); }
    static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
    static __device__ __forceinline__ double acos_(double x) { return acos(x); }
};

// GLSL built-ins with the conventions documented in include/gravitas_b200.h (GvtGlslUniforms)
template <class R> __device__ __forceinline__ R gl_max(R a, R b) { return a < b ? b : a; }
template <class R> __device__ __forceinline__ R gl_min(R a, R b) { return b < a ? b : a; }
template <class R> __device__ __forceinline__ R gl_clamp(R x, R lo, R hi) { return gl_min(gl_max(x, lo), hi); }
template <class R> __device__ __forceinline__ R gl_mix(R a, R b, R t) { return a * (R(1) - t) + b * t; }
template <class R> __device__ __forceinline__ R gl_smoothstep(R e0, R e1, R x) {
    const R t = gl_clamp((x - e0) / (e1 - e0), R(0), R(1));
    return t * t * (R(3) - R(2) * t);
}
template <class R> __device__ __forceinline__ R gl_sign(R x) { return x > R(0) ? R(1) : (x < R(0) ? R(-1) : R(0)); }
template <class R> __device__ __forceinline__ R len3(const Vec3<R>& a) { return GM<R>::sqrt_(a.x * a.x + a.y * a.y + a.z * a.z); }
template <class R> __device__ __forceinline__ Vec3<R> cross3(const Vec3<R>& a, const Vec3<R>& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <class R> __device__ __forceinline__ Vec3<R> unit3(const Vec3<R>& a) {
    const R n = len3(a);
    return {a.x / n, a.y / n, a.z / n};
}
// (a, b) *= rot(t), rot(t) = mat2(c, -s, s, c) (chunks/common.ts:46-49)
template <class R> __device__ __forceinline__ void rot_pair(R ang, R& a, R& b) {
    R s, c;
    GM<R>::sincos_(ang, &s, &c);
    const R na = a * c - b * s, nb = a * s + b * c;
    a = na; b = nb;
}

// The sampling context: noise texture red channel in shared memory, uniforms in the parameter bank.
template <class R> struct Ctx {
    const uint8_t* noise;   // shared memory, 256*256
    R time;

    // texture(u_noiseTex, (uv + 0.5)/256).r  -- LINEAR + REPEAT (chunks/noise.ts:3-8; webgl-utils.ts:259-276)
    __device__ __forceinline__ R hash(R px, R py, R pz) const {
        const R ux = px + pz * R(37.0), uy = py + pz * R(37.0);
        const R x = ((ux + R(0.5)) / R(256.0)) * R(256) - R(0.5), y = ((uy + R(0.5)) / R(256.0)) * R(256) - R(0.5);
        const R xf = GM<R>::floor_(x), yf = GM<R>::floor_(y);
        const R fx = x - xf, fy = y - yf;
        const int xi = (int)(long long)xf, yi = (int)(long long)yf;
        const int x0 = xi & 255, x1 = (xi + 1) & 255, y0 = (yi & 255) << 8, y1 = ((yi + 1) & 255) << 8;
        // integer lattice points (every call from noise()) land on texel centres: fx = fy = 0 and one tap suffices
        if (fx == R(0) && fy == R(0)) return R((int)noise[y0 + x0]) / R(255);
        const R t00 = R((int)noise[y0 + x0]) / R(255), t10 = R((int)noise[y0 + x1]) / R(255);
        const R t01 = R((int)noise[y1 + x0]) / R(255), t11 = R((int)noise[y1 + x1]) / R(255);
        return gl_mix(gl_mix(t00, t10, fx), gl_mix(t01, t11, fx), fy);
    }
    __device__ __noinline__ R hash_far(R px, R py, R pz) const { return hash(px, py, pz); }   // rare path, kept out of line
    // chunks/noise.ts:11-19. The eight lattice hashes are texel-centre fetches: (i.xy + i.z*37) & 255.
    __device__ __forceinline__ R noise3(R px, R py, R pz) const {
        const R ix = GM<R>::floor_(px), iy = GM<R>::floor_(py), iz = GM<R>::floor_(pz);
        R fx = px - ix, fy = py - iy, fz = pz - iz;
        fx = fx * fx * (R(3) - R(2) * fx); fy = fy * fy * (R(3) - R(2) * fy); fz = fz * fz * (R(3) - R(2) * fz);
        R h[2][2][2];
        // The eight lattice points are integers, so (uv + 0.5)/256 is a texel centre and the LINEAR fetch is one exact
        // tap at ((i.x + 37 i.z) mod 256, (i.y + 37 i.z) mod 256): integer address arithmetic, as long as the float
        // expression p.xy + p.z*37 + 0.5 is itself exact (|.| < 2^23; otherwise take the generic path, which rounds
        // exactly like the shader's arithmetic would).
        if (GM<R>::abs_(ix) < R(2097152.0) && GM<R>::abs_(iy) < R(2097152.0) && GM<R>::abs_(iz) < R(65536.0)) {
            const int kz = 37 * (int)iz;
            const int bx = (int)ix + kz, by = (int)iy + kz;
#pragma unroll
            for (int dz = 0; dz < 2; dz++)
#pragma unroll
                for (int dy = 0; dy < 2; dy++)
#pragma unroll
                    for (int dx = 0; dx < 2; dx++)
                        h[dz][dy][dx] = R((int)noise[(((by + dy + 37 * dz) & 255) << 8) | ((bx + dx + 37 * dz) & 255)]) / R(255);
        } else {
#pragma unroll
            for (int q = 0; q < 8; q++) h[q >> 2][(q >> 1) & 1][q & 1] = hash_far(ix + R(q & 1), iy + R((q >> 1) & 1), iz + R(q >> 2));
        }
        const R a0 = gl_mix(gl_mix(h[0][0][0], h[0][0][1], fx), gl_mix(h[0][1][0], h[0][1][1], fx), fy);
        const R a1 = gl_mix(gl_mix(h[1][0][0], h[1][0][1], fx), gl_mix(h[1][1][0], h[1][1][1], fx), fy);
        return gl_mix(a0, a1, fz);
    }
    __device__ R fbm(R px, R py, R pz) const {   // chunks/noise.ts:22-31
        R f = R(0), amp = R(0.5);
        for (int i = 0; i < 4; i++) {
            f += amp * noise3(px, py, pz);
            px *= R(2.0); py *= R(2.0); pz *= R(2.0);
            amp *= R(0.5);
        }
        return f;
    }
    __device__ __forceinline__ void star_color(R bv, R c[3]) const {   // chunks/blackbody.ts:38-47
        const R t = gl_clamp(bv, R(-0.4), R(2.0));
        if (t < R(0.0)) { c[0] = R(0.6); c[1] = R(0.7); c[2] = R(1.0); }
        else if (t < R(0.3)) { c[0] = R(0.85); c[1] = R(0.88); c[2] = R(1.0); }
        else if (t < R(0.6)) { c[0] = R(1.0); c[1] = R(0.96); c[2] = R(0.9); }
        else if (t < R(1.0)) { c[0] = R(1.0); c[1] = R(0.85); c[2] = R(0.6); }
        else { c[0] = R(1.0); c[1] = R(0.6); c[2] = R(0.4); }
    }
    // chunks/background.ts:3-30
    __device__ void starfield(const Vec3<R>& d, R out[3]) const {
        using M = GM<R>;
        R st[3] = {R(0), R(0), R(0)};
        R cx = M::floor_(d.x * R(200.0)), cy = M::floor_(d.y * R(200.0)), cz = M::floor_(d.z * R(200.0));
        R sn = hash(cx, cy, cz);
        if (sn > R(0.998)) {
            const R brightness = M::pow_(sn, R(10.0)) * R(2.0);
            const R bv = hash(cx + R(127.1), cy + R(127.1), cz + R(127.1)) * R(2.4) - R(0.4);
            const R tw = R(0.85) + R(0.15) * M::sin_(time * (R(3.0) + hash(cx + R(73.7), cy + R(73.7), cz + R(73.7)) * R(2.0)));
            R c[3];
            star_color(bv, c);
            for (int k = 0; k < 3; k++) st[k] = c[k] * brightness * tw;
        }
        cx = M::floor_(d.x * R(500.0)); cy = M::floor_(d.y * R(500.0)); cz = M::floor_(d.z * R(500.0));
        sn = hash(cx, cy, cz);
        if (sn > R(0.996)) {
            const R brightness = M::pow_(sn, R(20.0)) * R(1.5);
            const R bv = hash(cx + R(217.3), cy + R(217.3), cz + R(217.3)) * R(2.4) - R(0.4);
            R c[3];
            star_color(bv, c);
            for (int k = 0; k < 3; k++) st[k] += c[k] * brightness;
        }
        const R tt = time * R(0.01);
        const R neb = fbm(d.x * R(2.0) + tt, d.y * R(2.0) + tt, d.z * R(2.0) + tt) * R(0.03);
        const R ln = M::abs_(neb);
        out[0] = st[0] + (neb * R(0.2) + R(0.05) * ln);
        out[1] = st[1] + (neb * R(0.3) + R(0.02) * ln);
        out[2] = st[2] + (neb * R(0.5) + R(0.05) * ln);
    }
};

// chunks/blackbody.ts:9-35 (Tanner-Helland fit, sRGB -> linear by pow 2.2)
template <class R> __device__ __forceinline__ void blackbody(R temp, R c[3]) {
    using M = GM<R>;
    const R t = gl_max(temp, R(1.0)) / R(100.0);
    R r, g, b;
    if (t <= R(66.0)) {
        r = R(255.0);
        g = R(99.4708025861) * M::log_(t) - R(161.1195681661);
        b = (t <= R(19.0)) ? R(0.0) : R(138.5177312231) * M::log_(t - R(10.0)) - R(305.0447927307);
    } else {
        r = R(329.698727446) * M::pow_(t - R(60.0), R(-0.1332047592));
        g = R(288.1221695283) * M::pow_(t - R(60.0), R(-0.0755148492));
        b = R(255.0);
    }
    c[0] = M::pow_(gl_max(r / R(255.0), R(0)), R(2.2));
    c[1] = M::pow_(gl_max(g / R(255.0), R(0)), R(2.2));
    c[2] = M::pow_(gl_max(b / R(255.0), R(0)), R(2.2));
}

// chunks/metric.ts:96-149 kerr_geodesic_accel, as the shader writes it (divisions kept: this path is a behavioural
// restatement of the f32 shader, not the headline kernel)
// `pn` = |p| = sqrt(dot(p, p)), which the march already holds (it is the step's r or r_new): normalize(p) reuses it.
template <class R>
__device__ __forceinline__ void shader_accel(const Vec3<R>& p, const Vec3<R>& v, R pn, R M, R a, Vec3<R>& acc, R& omega) {
    using G = GM<R>;
    const R a2 = a * a;
    const R rho2 = p.x * p.x + p.y * p.y + p.z * p.z;
    const R diff = rho2 - a2;
    const R disc = diff * diff + R(4.0) * a2 * p.y * p.y;
    const R r2 = R(0.5) * (diff + G::sqrt_(gl_max(R(0.0), disc)));
    const R r_k = G::sqrt_(gl_max(R(1e-8), r2));
    const R sigma = r2 + a2 * (p.y * p.y / gl_max(R(1e-8), r2));
    const Vec3<R> L = cross3(p, v);
    const R Ly_eff = L.y - a;
    const R L2_eff = Ly_eff * Ly_eff + ((L.x * L.x + L.y * L.y + L.z * L.z) - L.y * L.y);
    const R r_inv = R(1.0) / r_k;
    const R r2_inv = r_inv * r_inv;
    const R r4_inv = r2_inv * r2_inv;
    const R sigma_ratio = r2 / gl_max(R(1e-8), sigma);
    const R f = M * r2_inv * sigma_ratio + R(3.0) * M * gl_max(R(0.0), L2_eff) * r4_inv * sigma_ratio;
    const R r3_p_a2r = r_k * r2 + a2 * r_k;
    const R drag = R(2.0) * M * a / gl_max(R(1e-8), r3_p_a2r);
    acc.x = -(p.x / pn) * f + v.z * drag;
    acc.y = -(p.y / pn) * f + R(0) * drag;
    acc.z = -(p.z / pn) * f + (-v.x) * drag;
    omega = drag;
}

}  // namespace

// resident CTAs per SM: three 64 KB noise copies fit in shared memory; the MUFU build fits 3 x 256 threads in 77
// registers without spilling (15.6 ms vs 16.2 ms at 4K), the IEEE/libm builds need ~104-128 registers
#ifndef GVT_FRAG_THREADS
#define GVT_FRAG_THREADS 256
#endif
#ifndef GVT_FRAG_MINB
#ifdef GVT_FRAGMENT_FAST
#define GVT_FRAG_MINB 3
#else
#define GVT_FRAG_MINB 2
#endif
#endif
template <class R>
__global__ void __launch_bounds__(GVT_FRAG_THREADS, GVT_FRAG_MINB) k_fragment_glsl(const __grid_constant__ GlslParams P) {
    using G = GM<R>;
    extern __shared__ __align__(128) unsigned char smem_noise[];
    __shared__ uint64_t bar;
    const GvtGlslUniforms& U = P.u;

    // ---- stage the noise texture's red channel: one 64 KB TMA bulk copy per CTA ----
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)), "r"(65536u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_addr(smem_noise)),
                     "l"(P.noise_r), "r"(65536u), "r"(smem_addr(&bar))
                     : "memory");
    }
    {
        const uint32_t b = smem_addr(&bar);
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "NOISE_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra NOISE_DONE;\n"
            "bra NOISE_WAIT;\n"
            "NOISE_DONE:\n"
            "}\n" ::"r"(b),
            "r"(0u)
            : "memory");
    }
    Ctx<R> ctx;
    ctx.noise = smem_noise;
    ctx.time = R(U.time);

    // ---- per-frame constants (fragment.glsl.ts:63-73) ----
    const R M = R(U.mass), rs = M * R(2.0), a = R(U.spin) * M, absA = G::abs_(R(U.spin));
    const R rh = M + G::sqrt_(gl_max(R(0.0), M * M - a * a));                                   // kerr_horizon
    R rph, isco;
    {
        const R a_star = gl_clamp(a / M, R(-0.9999), R(0.9999));                                // kerr_photon_sphere
        rph = R(2.0) * M * (R(1.0) + G::cos_(R(2.0 / 3.0) * G::acos_(gl_clamp(-a_star, R(-1.0), R(1.0)))));
        const R absS = G::abs_(a_star);                                                          // kerr_isco
        const R z1 = R(1.0) + G::pow_(R(1.0) - absS * absS, R(1.0 / 3.0)) *
                                  (G::pow_(R(1.0) + absS, R(1.0 / 3.0)) + G::pow_(R(1.0) - absS, R(1.0 / 3.0)));
        const R z2 = G::sqrt_(R(3.0) * absS * absS + z1 * z1);
        R sg = gl_sign(a);
        if (sg == R(0.0)) sg = R(1.0);
        isco = M * (R(3.0) + z2 - sg * G::sqrt_((R(3.0) - z1) * (R(3.0) + z1 + R(2.0) * z2)));
    }
    const R PI = R(3.14159265359), MAX_DIST = R(10000.0), MIN_STEP = R(0.01), MAX_STEP = R(1.2);
    const R lens = R(U.lensing_strength);
    const uint32_t feat = U.features;
    const R resx = R(U.resolution[0]), resy = R(U.resolution[1]);
    const R minRes = gl_min(resx, resy);
    const int maxSteps = (int)gl_min(R((double)U.max_ray_steps), R(500.0));
    const bool show_red = U.show_redshift > 0.5f;

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_rows = P.stripe.s ? P.n_lattice_rows : (P.y1 - P.y0 + P.ys - 1u) / P.ys;
    const uint32_t tiles_x = (P.width + 7u) / 8u, tiles_y = (n_rows + 3u) / 4u;
    const uint32_t n_tiles = tiles_x * tiles_y;
    unsigned long long acc_steps = 0;
    uint32_t acc_hit = 0, acc_other = 0;

    for (;;) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(&P.counters->tile_counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= n_tiles) break;
        const uint32_t px = (tile % tiles_x) * 8u + (lane & 7u), lj = (tile / tiles_x) * 4u + (lane >> 3);
        uint32_t py = P.y0 + lj * P.ys;
        bool valid = px < P.width && py < P.y1;
        if (P.stripe.s) {   // GVT_FLAG_ROW_INTERLEAVE: this rank's stripes (+ TAA halo rows)
            valid = lj < n_rows && stripe_row(P.stripe, lj, P.height, py) && px < P.width;
        }
        R out[3] = {R(0), R(0), R(0)};
        uint32_t steps = 0;
        bool hitHorizon = false;
        if (valid) {
            const R uvx = ((R((double)px) + R(0.5)) - R(0.5) * resx) / minRes;
            const R uvy = ((R((double)py) + R(0.5)) - R(0.5) * resy) / minRes;
            if (U.debug > 0.5f) {
                out[0] = uvx + R(0.5); out[1] = uvy + R(0.5); out[2] = R(0.0);
            } else {
                // ---- camera (fragment.glsl.ts:50-61) ----
                Vec3<R> ro, rd;
                {
                    const Vec3<R> cp = {R(U.cam_pos[0]), R(U.cam_pos[1]), R(U.cam_pos[2])};
                    if (len3(cp) > R(0.001)) {
                        ro = cp;
                        const Vec3<R> d = unit3(Vec3<R>{uvx, uvy, R(1.2)});
                        const Vec3<R> q = {R(U.cam_quat[0]), R(U.cam_quat[1]), R(U.cam_quat[2])};
                        const R qw = R(U.cam_quat[3]);
                        const Vec3<R> c1 = cross3(q, d);
                        const Vec3<R> c2 = cross3(q, Vec3<R>{c1.x + d.x * qw, c1.y + d.y * qw, c1.z + d.z * qw});
                        rd = {d.x + c2.x * R(2.0), d.y + c2.y * R(2.0), d.z + c2.z * R(2.0)};        // qrot
                    } else {
                        ro = {R(0.0), R(0.0), -R(U.zoom)};
                        rd = unit3(Vec3<R>{uvx, uvy, R(1.5)});
                        const R ax = (R(U.mouse[1]) - R(0.5)) * PI, ay = (R(U.mouse[0]) - R(0.5)) * PI * R(2.0);
                        rot_pair(ax, ro.y, ro.z); rot_pair(ax, rd.y, rd.z);
                        rot_pair(ay, ro.x, ro.z); rot_pair(ay, rd.x, rd.z);
                    }
                }
                if (feat & GVT_GLSL_QUALITY_LOW) {
                    // ---- low-quality mode: no march (fragment.glsl.ts:76-87) ----
                    R bg[3];
                    ctx.starfield(rd, bg);
                    const R d = len3(cross3(ro, rd));
                    const R shadow = gl_smoothstep(rh * R(1.2), rh * R(0.9), d);
                    const R glow = G::exp_(-G::abs_(d - rph) * R(12.0)) * R(0.8);
                    const R mask = gl_smoothstep(isco * R(2.0), isco * R(1.0), d) * (R(1.0) - gl_smoothstep(isco * R(1.0), isco * R(0.8), d));
                    const R gc[3] = {R(0.3), R(0.6), R(1.0)}, dc[3] = {R(1.0), R(0.7), R(0.3)};
                    for (int k = 0; k < 3; k++)
                        out[k] = G::pow_(bg[k] * (R(1.0) - shadow) + gc[k] * glow + dc[k] * mask * R(0.6), R(0.4545));
                } else {
                    // ---- Kerr march (fragment.glsl.ts:89-221) ----
                    if (len3(ro) < rh * R(1.5)) { const Vec3<R> n = unit3(ro); ro = {n.x * rh * R(1.5), n.y * rh * R(1.5), n.z * rh * R(1.5)}; }
                    Vec3<R> p = ro, v = rd;
                    R col[3] = {R(0), R(0), R(0)}, alpha = R(0.0), maxRedshift = R(0.0);
                    const R bNoise = R((int)P.blue_r[((py & 255u) << 8) + (px & 255u)]) / R(255);     // NEAREST, texel centre
                    p.x += v.x * bNoise * MIN_STEP; p.y += v.y * bNoise * MIN_STEP; p.z += v.z * bNoise * MIN_STEP;
                    int photon = 0;
                    R prevY = p.y;
                    bool red_init = false;
                    if (len3(cross3(ro, rd)) < rh * R(0.9)) hitHorizon = true;                       // inner shadow culling
                    R r_next = len3(p);                                                              // |p| is carried from step to step
                    for (int i = 0; i < maxSteps; i++) {
                        const Vec3<R> pp = p;
                        const R r = r_next;
                        if (r < rh * R(1.15)) { hitHorizon = true; break; }
                        if (r > MAX_DIST) break;
                        const R distFactor = R(1.0) + r * R(0.05);
                        R dt = gl_clamp((r - rh) * R(0.1) * distFactor, MIN_STEP, MAX_STEP * distFactor);
                        if (r > R(30.0)) {
                            dt = gl_max(dt, MIN_STEP + (r - R(30.0)) * R(0.08));
                            dt = gl_min(dt, MAX_STEP * R(2.5));
                        }
                        dt = gl_min(dt, MIN_STEP + G::abs_(r - rph) * R(0.15));
                        const R cdt = dt * (R(1.0) - gl_smoothstep(R(0.2), R(0.0), G::abs_(p.y)) * R(0.7));
                        Vec3<R> acc = {R(0), R(0), R(0)};
                        if (feat & GVT_GLSL_LENSING) {
                            R omega;
                            shader_accel<R>(p, v, r, M, a, acc, omega);
                            acc.x *= lens; acc.y *= lens; acc.z *= lens;
                            rot_pair(omega * cdt, v.x, v.z);                                           // ZAMO twist of v.xz
                        }
                        p.x += v.x * cdt + acc.x * R(0.5) * cdt * cdt;
                        p.y += v.y * cdt + acc.y * R(0.5) * cdt * cdt;
                        p.z += v.z * cdt + acc.z * R(0.5) * cdt * cdt;
                        const R r_new = len3(p);
                        r_next = r_new;
                        if ((feat & GVT_GLSL_LENSING) && alpha < R(0.95)) {
                            Vec3<R> acc2; R om2;
                            shader_accel<R>(p, v, r_new, M, a, acc2, om2);
                            v.x += (acc.x + acc2.x * lens) * R(0.5) * cdt;
                            v.y += (acc.y + acc2.y * lens) * R(0.5) * cdt;
                            v.z += (acc.z + acc2.z * lens) * R(0.5) * cdt;
                        }
                        v = unit3(v);
                        if (prevY * p.y < R(0.0) && r_new < rph * R(2.0) && r_new > rh) photon = min(photon + 1, 3);
                        if (show_red) {
                            const R pot = G::sqrt_(gl_max(R(0.0), R(1.0) - rs / r_new));
                            maxRedshift = red_init ? gl_min(maxRedshift, pot) : pot;
                            red_init = true;
                        }
                        prevY = p.y;
                        steps++;
                        if (feat & GVT_GLSL_DISK) {
                            // ---- sample_accretion_disk (chunks/disk.ts:16-118) ----
                            if (!show_red) {
                                const bool crossed = pp.y * p.y < R(0.0);
                                Vec3<R> sp = p;
                                if (crossed) {
                                    const R t = G::abs_(pp.y) / gl_max(R(0.0001), G::abs_(pp.y) + G::abs_(p.y));
                                    sp = {gl_mix(pp.x, p.x, t), gl_mix(pp.y, p.y, t), gl_mix(pp.z, p.z, t)};
                                }
                                const R sr = crossed ? len3(sp) : r_new;
                                const R esh = gl_min(R(U.disk_scale_height), R(0.450));
                                const R diskOuter = gl_max(M * R(U.disk_size), isco * R(1.1));
                                if ((G::abs_(sp.y) < sr * esh || crossed) && sr > isco && sr < diskOuter) {
                                    const R sqM = G::sqrt_(M);
                                    const R sgn = gl_sign(R(U.spin) + R(1e-8));
                                    const R Omega = (sgn * sqM) / (sr * G::sqrt_(sr) + a * sqM);
                                    R nx = sp.x, nz = sp.z;
                                    rot_pair(Omega * ctx.time * R(0.12) * R(10.0), nx, nz);
                                    nx *= R(0.75); nz *= R(0.75);
                                    const R ny = sp.y * R(0.75);
                                    const R turb = ctx.noise3(nx, ny, nz) * R(0.5) + ctx.noise3(nx * R(2.5), ny * R(2.5), nz * R(2.5)) * R(0.25);
                                    const R hfall = G::exp_(-G::abs_(sp.y) / gl_max(R(0.001), (sr * esh) * R(0.25)));
                                    const R base = turb * hfall * gl_smoothstep(diskOuter, isco, sr);
                                    if (base > R(0.001)) {
                                        const R g_tt = -(R(1.0) - R(2.0) * M / sr);
                                        const R g_tphi = R(-2.0) * M * a / sr;
                                        const R g_pp = sr * sr + a * a + R(2.0) * M * a * a / sr;
                                        const R ut2 = -(g_tt + R(2.0) * Omega * g_tphi + Omega * Omega * g_pp);
                                        const R u_t = R(1.0) / G::sqrt_(gl_max(R(1e-6), ut2));
                                        const R Lph = p.z * v.x - p.x * v.z;
                                        const R delta = R(1.0) / gl_max(R(0.01), u_t * (R(1.0) - Omega * Lph));
                                        const R beaming = (feat & GVT_GLSL_DOPPLER) ? gl_max(R(0.01), G::pow_(delta, R(3.5))) : R(1.0);
                                        const R ir = gl_clamp(isco / sr, R(0.0), R(1.0));
                                        const R ntf = gl_max(R(0.0), R(1.0) - G::sqrt_(ir));
                                        const R grad = G::pow_(ir, R(0.75)) * G::pow_(ntf, R(0.25));
                                        R bb[3];
                                        blackbody<R>(R(U.disk_temp) * grad * delta, bb);
                                        const R density = base * R(U.disk_density) * R(0.12) * cdt;
                                        const R w = R(1.0) - alpha;
                                        for (int k = 0; k < 3; k++) col[k] += bb[k] * beaming * density * w;
                                        alpha += density;
                                    }
                                }
                            }
                            if (alpha > R(0.99)) break;
                        }
                        if (feat & GVT_GLSL_JETS) {
                            // ---- sample_relativistic_jets (chunks/disk.ts:120-155); note: pre-step r's dt, not cdt ----
                            const R jv = G::abs_(p.y);
                            if (jv > rh * R(1.8) && jv < MAX_DIST * R(0.8)) {
                                const R jr = G::sqrt_(p.x * p.x + p.z * p.z);
                                const R jw = R(1.0) + jv * R(0.15);
                                if (jr < jw * R(2.0)) {
                                    const R rf = G::exp_(-(jr * jr) / (jw * R(0.5)));
                                    const R lf = G::exp_(-jv * R(0.05));
                                    const R flow = p.y * R(2.0) - ctx.time * R(8.0);
                                    const R nv = ctx.noise3(p.x * R(0.5), flow * R(0.5), p.z * R(0.5)) * R(0.6) +
                                                 ctx.noise3(p.x * R(1.5), flow * R(1.5), p.z * R(1.5)) * R(0.4);
                                    const R jd = rf * lf * gl_max(R(0.0), nv - R(0.2));
                                    if (jd > R(0.001)) {
                                        const R jetVel = R(0.92) * gl_sign(p.y);
                                        const R beta = G::abs_(jetVel);
                                        const R cosT = R(0.0) * -v.x + (jetVel / G::sqrt_(R(0.0) * R(0.0) + jetVel * jetVel + R(0.0) * R(0.0))) * -v.y + R(0.0) * -v.z;
                                        const R gam = R(1.0) / G::sqrt_(R(1.0) - beta * beta);
                                        const R dj = R(1.0) / (gam * (R(1.0) - beta * cosT));
                                        const R bj = G::pow_(dj, R(3.5));
                                        const R w = R(1.0) - alpha;
                                        const R jc[3] = {R(0.4), R(0.7), R(1.0)};
                                        for (int k = 0; k < 3; k++) col[k] += jc[k] * jd * R(0.05) * bj * dt * w;
                                        alpha += jd * R(0.05) * dt;
                                    }
                                }
                            }
                        }
                    }
                    // ---- compositing (fragment.glsl.ts:223-330) ----
                    if ((feat & GVT_GLSL_REDSHIFT) && show_red) {
                        const R val = hitHorizon ? R(0.0) : maxRedshift;
                        const R s1 = gl_smoothstep(R(0.0), R(0.3), val), s2 = gl_smoothstep(R(0.3), R(0.7), val),
                                s3 = gl_smoothstep(R(0.7), R(1.0), val);
                        R h[3] = {gl_mix(R(0), R(1), s1), gl_mix(R(0), R(0), s1), gl_mix(R(0), R(0), s1)};
                        h[0] = gl_mix(h[0], R(1), s2); h[1] = gl_mix(h[1], R(1), s2); h[2] = gl_mix(h[2], R(0), s2);
                        h[0] = gl_mix(h[0], R(0), s3); h[1] = gl_mix(h[1], R(0), s3); h[2] = gl_mix(h[2], R(1), s3);
                        out[0] = h[0]; out[1] = h[1]; out[2] = h[2];
                    } else {
                        R bg[3] = {R(0), R(0), R(0)};
                        if (feat & GVT_GLSL_STARS) ctx.starfield(v, bg);
                        R ring = R(0.0);
                        if ((feat & GVT_GLSL_PHOTON_GLOW) && !hitHorizon) {
                            const R dpr = G::abs_(len3(p) - rph);
                            ring = G::exp_(-dpr * R(40.0)) * R(1.8) * lens;
                            if (photon > 0) {
                                const R sharp = R(60.0) + R((double)photon) * R(30.0);
                                const R bright = G::exp_(-R((double)photon) * R(1.0)) * R(1.2);
                                ring = ring + G::exp_(-dpr * sharp) * bright * lens;
                            }
                        }
                        R ergo = R(0.0);
                        if (absA > R(0.1) && !hitHorizon) {
                            const R rF = len3(p);
                            const R cT = p.y / gl_max(rF, R(0.001));
                            const R r_ergo = M + G::sqrt_(gl_max(R(0.0), M * M - a * a * cT * cT));
                            ergo = G::exp_(-G::abs_(rF - r_ergo) * R(20.0)) * R(0.35) * absA;
                        }
                        if (hitHorizon) { bg[0] = bg[1] = bg[2] = R(0); }
                        const R w = R(1.0) - alpha;
                        const R ec[3] = {R(0.3), R(0.35), R(0.9)};
                        R fin[3];
                        for (int k = 0; k < 3; k++) fin[k] = bg[k] * w + col[k] + ring * w + (ec[k] * ergo) * w;
                        if (U.show_kerr_shadow > 0.5f) {
                            // ---- Kerr shadow guide: distance to the critical-curve polyline (:279-322) ----
                            const Vec3<R> cam_dir = unit3(ro);
                            const Vec3<R> sky_right = unit3(cross3(Vec3<R>{R(0), R(1), R(0)}, cam_dir));
                            const Vec3<R> sky_up = cross3(cam_dir, sky_right);
                            const R lro = len3(ro);
                            Vec3<R> iv = cross3(cam_dir, rd);
                            iv = {iv.x * lro, iv.y * lro, iv.z * lro};
                            const R al = -dot3(iv, sky_up), be = dot3(iv, sky_right);
                            R minDist = R(1e10);
                            const int count = (int)U.shadow_count;
                            auto seg = [&](int i1, int i2) {
                                const R p1x = R(U.shadow_curve[2 * i1]), p1y = R(U.shadow_curve[2 * i1 + 1]);
                                const R bax = R(U.shadow_curve[2 * i2]) - p1x, bay = R(U.shadow_curve[2 * i2 + 1]) - p1y;
                                const R pax = al - p1x, pay = be - p1y;
                                const R h = gl_clamp((pax * bax + pay * bay) / (bax * bax + bay * bay), R(0.0), R(1.0));
                                const R dx = pax - bax * h, dy = pay - bay * h;
                                minDist = gl_min(minDist, G::sqrt_(dx * dx + dy * dy));
                            };
                            for (int j = 0; j < 63 && j < count - 1; j++) seg(j, j + 1);
                            if (count > 2) seg(min(count, 64) - 1, 0);
                            const R th = M * R(0.045);
                            if (minDist < th) {
                                const R edge = gl_smoothstep(th, th * R(0.5), minDist);
                                fin[0] = gl_mix(fin[0], R(0), edge); fin[1] = gl_mix(fin[1], R(1), edge); fin[2] = gl_mix(fin[2], R(0), edge);
                            }
                        }
                        if (!(feat & GVT_GLSL_LINEAR_OUTPUT)) {
                            for (int k = 0; k < 3; k++) {   // ACES (Narkowicz) then gamma 1/2.2
                                const R x = fin[k];
                                const R t = gl_clamp((x * (R(2.51) * x + R(0.03))) / (x * (R(2.43) * x + R(0.59)) + R(0.14)), R(0.0), R(1.0));
                                fin[k] = G::pow_(gl_max(t, R(0.0)), R(0.4545));
                            }
                        }
                        out[0] = fin[0]; out[1] = fin[1]; out[2] = fin[2];
                    }
                }
            }
            float4 px_out = make_float4((float)out[0], (float)out[1], (float)out[2], 1.0f);
            // NaN guard: GLSL leaves a NaN fragment undefined; here it becomes black instead of poisoning TAA / bloom
            if (!(px_out.x == px_out.x) || !(px_out.y == px_out.y) || !(px_out.z == px_out.z)) px_out = make_float4(0.f, 0.f, 0.f, 1.0f);
            const size_t o = (size_t)py * P.width + px;
            const bool f16 = P.frame_f16 != 0u;
            if (P.frame) frag_store_px(P.frame, o, px_out, f16);
            if (P.host_frame) frag_store_px(P.host_frame, o, px_out, f16);
#pragma unroll 1
            for (uint32_t q = 0; q < P.n_peer; q++) frag_store_px(P.peer_frame[q], o, px_out, f16);
            if (P.dbg_steps) P.dbg_steps[o] = steps;
            if (P.dbg_hit) P.dbg_hit[o] = hitHorizon ? 1u : 0u;
        }
        acc_steps += __reduce_add_sync(0xffffffffu, valid ? steps : 0u);
        acc_hit += __popc(__ballot_sync(0xffffffffu, valid && hitHorizon));
        acc_other += __popc(__ballot_sync(0xffffffffu, valid && !hitHorizon));
    }
    if (lane == 0) {
        atomicAdd(&P.counters->steps_committed, acc_steps);
        atomicAdd(&P.counters->steps_executed, acc_steps);
        atomicAdd(&P.counters->rhs_evals, 2ull * acc_steps);
        if (acc_hit) atomicAdd(&P.counters->n_horizon, (unsigned long long)acc_hit);
        if (acc_other) atomicAdd(&P.counters->n_escape, (unsigned long long)acc_other);
    }
}

static cudaError_t launch_impl(const GlslParams& p, int precision, int sm_count, cudaStream_t stream) {
    if (p.width == 0 || (p.stripe.s ? p.n_lattice_rows == 0 : (p.y1 <= p.y0 || p.ys == 0))) return cudaSuccess;
    const size_t smem = 65536;
#ifdef GVT_FRAGMENT_FAST
    (void)precision;
    auto kern = k_fragment_glsl<float>;
#else
    auto kern = precision == 1 ? k_fragment_glsl<float> : k_fragment_glsl<double>;
#endif
    // the shared-memory opt-in and the occupancy query are per kernel AND per device, not per frame (~10 us each on the host)
    static PerDeviceInt resident[2];
    const int ki = precision == 1 ? 1 : 0;
    int* cached = resident[ki].slot();
    int per_sm = cached ? *cached : 0;
    if (!per_sm) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int n = 1;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, GVT_FRAG_THREADS, smem);
        if (e != cudaSuccess) return e;
        per_sm = n > 0 ? n : 1;
        if (cached) *cached = per_sm;
    }
    const uint32_t n_rows = p.stripe.s ? p.n_lattice_rows : (p.y1 - p.y0 + p.ys - 1u) / p.ys;
    const uint32_t tiles = ((p.width + 7u) / 8u) * ((n_rows + 3u) / 4u);
    const uint32_t wpc = GVT_FRAG_THREADS / 32;
    uint32_t ctas = (tiles + wpc - 1u) / wpc;
    const uint32_t max_ctas = (uint32_t)sm_count * (uint32_t)per_sm;
    if (ctas > max_ctas) ctas = max_ctas;   // persistent CTAs: as many as are resident at once
    kern<<<ctas, GVT_FRAG_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace fragment_fast / fragment_precise

#ifdef GVT_FRAGMENT_FAST
cudaError_t launch_fragment_glsl_fast(const GlslParams& p, int sm_count, cudaStream_t stream) {
    return fragment_fast::launch_impl(p, 1, sm_count, stream);
}
#else
cudaError_t launch_fragment_glsl(const GlslParams& p, int precision, int sm_count, cudaStream_t stream) {
    return fragment_precise::launch_impl(p, precision, sm_count, stream);
}
#endif

}  // namespace gvt
