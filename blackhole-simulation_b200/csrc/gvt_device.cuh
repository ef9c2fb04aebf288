// gvt_device.cuh — device-side Kerr geodesic mathematics for sm_100a, templated on the scalar type.
//
// This is NOT a transcription of the reference's Rust: the Hamiltonian is refactored so that one RHS
// evaluation costs two reciprocals, one sincos and ~50 FMA-shaped flops instead of the 16 divisions / 3 trig
// calls / 129 flops of the as-written form (metric/kerr.rs:412-499 + geodesic/hamiltonian.rs:13-35):
//
//   Kerr-Schild form used by the reference:   H = -pt^2/2 + N / (2 Sigma)
//       N = 2Mr (2 pt pr - pt^2) + Delta pr^2 + pth^2 + pph^2 / sin^2 + 2 a pr pph
//       dH/dr  = (N_r/2 - r N/Sigma) / Sigma,          N_r/2  = M (2 pt pr - pt^2) + (r - M) pr^2
//       dH/dth = (N_th/2 + a^2 s c N/Sigma) / Sigma,   N_th/2 = - s c pph^2 / sin^4
//   with p_t and p_phi constants of the motion (dp_t = dp_phi = 0, hamiltonian.rs:30-33), so everything
//   built from them is hoisted out of the step loop.
//
// Results agree with the reference's formulas to rounding; parity against the f64 oracle is tested to 1e-6
// relative per RGBA component (tests/test_gpu_parity.py).
#pragma once
#include <cuda_runtime.h>

#include <stdint.h>

namespace gvt {

// --------------------------------------------------------------------------------------------------
// scalar traits
// --------------------------------------------------------------------------------------------------
// ---- lean reciprocal: MUFU.RCP64H seed (~2^-21) + one cubic Newton step on the FP64 FMA pipe; no IEEE special-case
// slow path (callers guarantee a normal, non-zero argument: Sigma >= r^2 > 0, sin^2 >= 1e-12). <= 1 ulp.
__device__ __forceinline__ double rcp_nr(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    // one cubically convergent step: y (1 + e + e^2), e = 1 - x y ~ 2^-21  ->  e^3 ~ 2^-63
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
}
__device__ __forceinline__ float rcp_nr(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));   // MUFU.RCP: 1 ulp, no refinement needed in f32
    return y;
}

// ---- what one Kerr-Schild RHS needs from theta: a with a^2 = sin^2 and |a| = |sin|, and sc = sin cos.
// Branch-free Cody-Waite reduction by the magic-number rint (valid for |theta| < 2^30 pi/2 -- geodesic polar
// angles stay within a few multiples of pi), fdlibm minimax kernels on |r| <= pi/4, quadrant handled by two
// selects and one sign flip: for odd k sin <-> cos and the product changes sign. ~19 FP64 ops, no I2F/F2I, no
// Payne-Hanek slow path (CUDA's sincos() costs ~2x that in issue slots).
// ---- call-free sqrt / x^0.4 for the march. CUDA's sqrt(), pow() and IEEE division compile to a fast path plus a
// CALL to a slow-path subroutine; a CALL anywhere inside the step loop makes ptxas keep every loop-invariant FP64
// constant in ordinary registers instead of uniform registers, and a DFMA with three distinct register operands
// costs 3 cycles instead of 2 on B200 (register-file port limit; scripts/ubench/issue_model.cu). Arguments here are
// finite and either zero or comfortably normal.
__device__ __forceinline__ double sqrt_nr(double x) {
    if (!(x > 1e-290)) return 0.0;                    // 0 (double root), or negative/NaN never reach here meaningfully
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));          // MUFU.RSQ64H seed
    double e = fma(-(x * y), y, 1.0);                                  // 1 - x y^2
    y = fma(y, e * fma(0.375, e, 0.5), y);                             // cubic: y (1 + e/2 + 3e^2/8)
    e = fma(-(x * y), y, 1.0);
    y = fma(0.5 * y, e, y);
    const double s = x * y;                                            // sqrt estimate, then one correction step
    return fma(0.5 * y, fma(-s, s, x), s);
}
__device__ __forceinline__ float sqrt_nr(float x) { return sqrtf(x); }    // MUFU.SQRT path, no call in f32
// t^0.4 = (t^2)^(1/5): MUFU.LG2/EX2 seed in f32, then Newton on y^5 = t^2 in the working precision.
__device__ __forceinline__ double pow04(double t) {
    if (!(t > 1e-30)) return 0.0;          // below this the LUT coordinate clamps to texel 0 anyway
    double y = (double)exp2f(0.4f * __log2f((float)t));
    const double c = t * t;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double y2 = y * y, y5 = y2 * y2 * y;
        y = y * fma(0.2, c * rcp_nr(y5), 0.8);
    }
    return y;
}
__device__ __forceinline__ float pow04(float t) { return (t > 0.0f) ? exp2f(0.4f * log2f(t)) : 0.0f; }
// x^(-1/5) for the RKF45 step controller (integrator.rs:90): division-free Newton on y^-5 = x, y <- y (6 - x y^5)/5
__device__ __forceinline__ double pow_m02(double x) {
    double y = (double)exp2f(-0.2f * __log2f((float)x));
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double y2 = y * y, y5 = y2 * y2 * y;
        y = y * fma(-0.2 * x, y5, 1.2);
    }
    return y;
}
__device__ __forceinline__ float pow_m02(float x) { return exp2f(-0.2f * log2f(x)); }
// x^(-1/4) (integrator.rs:96) = 1 / sqrt(sqrt(x))
__device__ __forceinline__ double pow_m025(double x) { return rcp_nr(sqrt_nr(sqrt_nr(x))); }
__device__ __forceinline__ float pow_m025(float x) { return rsqrtf(sqrtf(x)); }

// minimax kernels of fdlibm's __kernel_sin / __kernel_cos. The kernels receive the table inside their
// __grid_constant__ parameter block (constant bank 0), so the coefficients reach DFMA/FFMA through uniform
// registers instead of per-iteration literal materialisation.
struct TrigTable {
    double d[16];
    float f[16];
};
#define GVT_TRIG_TABLE_INIT                                                                                   \
    {{0.63661977236758138, 6755399441055744.0, 1.5707963267948966, 6.123233995736766e-17,                     \
      -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,                   \
      2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,                    \
      4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,                    \
      -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11},                  \
     {0.636619772f, 12582912.0f, 1.57079601e+00f, 3.13916473e-07f, 5.39030253e-15f, -1.6666654611e-1f,         \
      8.3321608736e-3f, -1.9515295891e-4f, 4.166664568298827e-2f, -1.388731625493765e-3f,                      \
      2.443315711809948e-5f, 0.f, 0.f, 0.f, 0.f, 0.f}}
// d: [0] 2/pi  [1] 1.5*2^52  [2] pi/2 hi  [3] pi/2 lo  [4..9] S1..S6  [10..15] C1..C6
// f: [0] 2/pi  [1] 1.5*2^23  [2..4] pi/2 in three parts  [5..7] S1..S3  [8..10] C1..C3

__constant__ TrigTable c_trig = GVT_TRIG_TABLE_INIT;
__device__ __forceinline__ void trig_pair(const TrigTable& T, double x, double& a, double& sc) {
#ifdef GVT_TRIG_NAMED_CONST
    const double* K = c_trig.d;
#else
    const double* K = T.d;
#endif
    const double kd_m = fma(x, K[0], K[1]);
    const int k = __double2loint(kd_m);
    const double kd = kd_m - K[1];
    // one-term Cody-Waite: pi/2 as a single double. The dropped tail (6.1e-17 per quadrant) moves the reduced argument
    // by < ulp(theta)/2 for the |k| <= a few that geodesic polar angles reach -- below the rounding theta itself carries
    // -- and saves one FP64 instruction per evaluation (3 per step; measured -1.3 % frame time, identical parity).
    const double r = fma(-kd, K[2], x);
    const double z = r * r;
    double ps = K[9];
    ps = fma(ps, z, K[8]);
    ps = fma(ps, z, K[7]);
    ps = fma(ps, z, K[6]);
    ps = fma(ps, z, K[5]);
    ps = fma(ps, z, K[4]);
    const double sr = fma(r * z, ps, r);
    double pc = K[15];
    pc = fma(pc, z, K[14]);
    pc = fma(pc, z, K[13]);
    pc = fma(pc, z, K[12]);
    pc = fma(pc, z, K[11]);
    pc = fma(pc, z, K[10]);
    const double cr = fma(z, fma(z, pc, -0.5), 1.0);
    const bool odd = (k & 1) != 0;
    a = odd ? cr : sr;
    const double p = sr * cr;
    // odd quadrant: sin cos changes sign -- flip the sign bit with one integer op
    sc = __hiloint2double(__double2hiint(p) ^ (k << 31), __double2loint(p));
}
// Full (sin, cos) of theta, quadrant and all: the base the rotated evaluations below start from. Same reduction and
// kernels as trig_pair; quadrant k mod 4 -> sin = {sr, cr, -sr, -cr}, cos = {cr, -sr, -cr, sr}.
__device__ __forceinline__ void trig_full(const TrigTable& T, double x, double& s, double& c) {
    const double* K = T.d;
    const double kd_m = fma(x, K[0], K[1]);
    const int k = __double2loint(kd_m);
    const double kd = kd_m - K[1];
    const double r = fma(-kd, K[2], x);
    const double z = r * r;
    double ps = K[9];
    ps = fma(ps, z, K[8]); ps = fma(ps, z, K[7]); ps = fma(ps, z, K[6]); ps = fma(ps, z, K[5]); ps = fma(ps, z, K[4]);
    const double sr = fma(r * z, ps, r);
    double pc = K[15];
    pc = fma(pc, z, K[14]); pc = fma(pc, z, K[13]); pc = fma(pc, z, K[12]); pc = fma(pc, z, K[11]); pc = fma(pc, z, K[10]);
    const double cr = fma(z, fma(z, pc, -0.5), 1.0);
    const bool odd = (k & 1) != 0;
    const double ss = odd ? cr : sr, cc = odd ? sr : cr;
    // sign bits: sin flips for k mod 4 in {2, 3}, cos for k mod 4 in {1, 2}
    s = __hiloint2double(__double2hiint(ss) ^ ((k & 2) << 30), __double2loint(ss));
    c = __hiloint2double(__double2hiint(cc) ^ (((k + 1) & 2) << 30), __double2loint(cc));
}
// (sin, cos)(theta0 + d) from (s0, c0) = (sin, cos)(theta0) for |d| <= 1/16: the angle-addition theorem with short
// series for sin d and cos d (first neglected terms d^11/11!, d^10/10! < 3e-19 relative). 14 FP64 instructions and no
// quadrant logic, against 19 + 7 for a full evaluation: the implicit midpoint's two predictor-shifted angles differ from
// the step's own by d = (h/2) p_theta / Sigma, tiny far from the hole. The coefficients are the first terms of the
// fdlibm kernels already in the table (they differ from the Taylor ones by < 1e-13 relative, i.e. < 1e-20 here).
__device__ __forceinline__ void trig_rot(const TrigTable& T, double s0, double c0, double d, double& s, double& c) {
    const double* K = T.d;
    const double z = d * d;
    double ps = fma(K[7], z, K[6]);
    ps = fma(ps, z, K[5]); ps = fma(ps, z, K[4]);
    const double sd = fma(d * z, ps, d);
    double pc = fma(K[12], z, K[11]);
    pc = fma(pc, z, K[10]);
    const double cd = fma(z, fma(z, pc, -0.5), 1.0);
    s = fma(c0, sd, s0 * cd);
    c = fma(-s0, sd, c0 * cd);
}
// The same for |d| <= 2^-8: two series terms less each (first neglected: d^7/7! and d^6/6!, < 5e-18 relative), 10 FP64
// instructions. What the far zone of the f64 kernel uses for all three rotations of a step.
__device__ __forceinline__ void trig_rot_small(const TrigTable& T, double s0, double c0, double d, double& s, double& c) {
    const double* K = T.d;
    const double z = d * d;
    const double sd = fma(d * z, fma(K[5], z, K[4]), d);
    const double cd = fma(z, fma(z, K[10], -0.5), 1.0);
    s = fma(c0, sd, s0 * cd);
    c = fma(-s0, sd, c0 * cd);
}
__device__ __forceinline__ void trig_pair(const TrigTable& T, float x, float& a, float& sc) {
    const float* K = T.f;
    const float kf_m = fmaf(x, K[0], K[1]);
    const int k = __float_as_int(kf_m);
    const float kf = kf_m - K[1];
    float r = fmaf(-kf, K[2], x);          // |k| is a handful for geodesic polar angles: two parts suffice
    r = fmaf(-kf, K[3], r);               // (|k| * 5.4e-15 is far below f32 resolution)
    const float z = r * r;
    float ps = K[7];
    ps = fmaf(ps, z, K[6]);
    ps = fmaf(ps, z, K[5]);
    const float sr = fmaf(r * z, ps, r);
    float pc = K[10];
    pc = fmaf(pc, z, K[9]);
    pc = fmaf(pc, z, K[8]);
    const float cr = fmaf(z, fmaf(z, pc, -0.5f), 1.0f);
    const bool odd = (k & 1) != 0;
    a = odd ? cr : sr;
    sc = __int_as_float(__float_as_int(sr * cr) ^ (k << 31));
}

template <class R> struct Num;
template <> struct Num<double> {
    static __device__ __forceinline__ double rcp(double x) { return rcp_nr(x); }
    static __device__ __forceinline__ double rcp_ieee(double x) { return 1.0 / x; }   // may be 0/inf (BL poles, horizon)
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
    static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
    static __device__ __forceinline__ double max_(double a, double b) { return fmax(a, b); }
    static __device__ __forceinline__ double min_(double a, double b) { return fmin(a, b); }
    static __device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
    static __device__ __forceinline__ void sincos_(double x, double* s, double* c) { sincos(x, s, c); }
    static __device__ __forceinline__ double pow_(double a, double b) { return pow(a, b); }
    static __device__ __forceinline__ double floor_(double a) { return floor(a); }
};
template <> struct Num<float> {
    static __device__ __forceinline__ float rcp(float x) { return rcp_nr(x); }
    static __device__ __forceinline__ float rcp_ieee(float x) { return 1.0f / x; }
    static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float abs_(float x) { return fabsf(x); }
    static __device__ __forceinline__ float max_(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ float min_(float a, float b) { return fminf(a, b); }
    static __device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ void sincos_(float x, float* s, float* c) { sincosf(x, s, c); }
    static __device__ __forceinline__ float pow_(float a, float b) { return powf(a, b); }
    static __device__ __forceinline__ float floor_(float a) { return floorf(a); }
};

// plain compare-selects (no NaN canonicalisation like fmin/fmax): inputs here are never NaN unless the ray
// already is, and then the result is garbage either way
template <class R> __device__ __forceinline__ R clampR(R x, R lo, R hi) {
    x = (x < lo) ? lo : x;
    return (x > hi) ? hi : x;
}
template <class R> __device__ __forceinline__ R floorAt(R x, R lo) { return (x > lo) ? x : lo; }

// x < c / x > c against a POSITIVE threshold c without touching the FP64 pipe: IEEE doubles are sign-magnitude, so for
// c > 0 the signed 64-bit compare of the bit patterns orders every non-NaN x (negatives and -0 sort below c) exactly
// as the floating-point compare does. DSETP issues on the FP64 pipe (2 cycles per warp, the pipe this kernel is bound
// by); the two ISETPs this compiles to run on the ALU pipe, which idles. NaN: a positive-sign NaN sorts above
// everything, a negative-sign one below (the FP compare would be false both ways) -- only ever seen on rays that
// have already blown up, and the termination test below treats both as "not alive" exactly like the FP form.
#ifdef GVT_NO_INTCMP   // A/B switch: plain FP64 compares (DSETP)
__device__ __forceinline__ bool lt_pos(double x, double c) { return !(x >= c); }
__device__ __forceinline__ bool gt_pos(double x, double c) { return !(x <= c); }
#else
__device__ __forceinline__ bool lt_pos(double x, double c) { return __double_as_longlong(x) < __double_as_longlong(c); }
__device__ __forceinline__ bool gt_pos(double x, double c) { return __double_as_longlong(x) > __double_as_longlong(c); }
#endif
__device__ __forceinline__ bool abs_lt_pos(double x, double c) {   // |x| < c  (hi/lo words: a 64-bit AND would come back as a DADD |x|)
#ifdef GVT_NO_INTCMP
    return fabs(x) < c;
#else
    const uint32_t xh = (uint32_t)__double2hiint(x) & 0x7fffffffu, ch = (uint32_t)__double2hiint(c);
    return xh < ch || (xh == ch && (uint32_t)__double2loint(x) < (uint32_t)__double2loint(c));
#endif
}
__device__ __forceinline__ bool abs_lt_pos(float x, float c) { return fabsf(x) < c; }
__device__ __forceinline__ bool lt_pos(float x, float c) { return !(x >= c); }  // FP32 compares are cheap; NaN counts as
__device__ __forceinline__ bool gt_pos(float x, float c) { return !(x <= c); }  // "beyond" here too
template <class R> __device__ __forceinline__ R clampPos(R x, R lo, R hi) {     // clampR for 0 < lo < hi
    x = lt_pos(x, lo) ? lo : x;
    return gt_pos(x, hi) ? hi : x;
}

// --------------------------------------------------------------------------------------------------
// Per-ray state. t (x[0]) is carried only when WITH_T (the parity hook); p_t and p_phi are constants.
// --------------------------------------------------------------------------------------------------
template <class R>
struct Ray {
    R t, r, th, ph;  // x^mu
    R pr, pth;       // the two evolving momenta
};
template <class R>
struct Deriv {
    R dt, dr, dth, dph, dpr, dpth;
};

// Constants of the hole + of the ray (p_t, p_phi and products), hoisted out of the loop.
template <class R>
struct HoleRay {
    const TrigTable* trig;
    R M, a, a2, twoM;
    R pt, pph;
    R pt2, pph2, two_pt, two_a_pph, a_pph;
    R twoM_pt, twoM_pt2, M_two_pt, M_pt2;     // 2M pt, 2M pt^2, 2M pt (= M * 2pt), M pt^2
    __device__ __forceinline__ void set_hole(R M_, R a_) { M = M_; a = a_; a2 = a_ * a_; twoM = R(2) * M_; }
    __device__ __forceinline__ void set_ray(R pt_, R pph_) {
        pt = pt_; pph = pph_; pt2 = pt_ * pt_; pph2 = pph_ * pph_; two_pt = R(2) * pt_;
        a_pph = a * pph_; two_a_pph = R(2) * a_pph;
        twoM_pt = twoM * pt_; twoM_pt2 = twoM * pt2; M_two_pt = M * two_pt; M_pt2 = M * pt2;
    }
};

// One Hamiltonian RHS in the reference's Kerr-Schild form (kerr.rs:412-499, hamiltonian.rs:13-35), returned
// UNSCALED: the true derivatives are (dr, dth, dph, dpr, dpth) * isig, dt is already scaled. The implicit-midpoint
// stepper folds isig into its step factor (one multiply instead of four).
// sin^2 is clamped at 1e-12 (kerr.rs:417,448); dH/dtheta is zeroed when |sin| < 1e-10 (kerr.rs:494-496).
// Inputs: a with a^2 = sin^2(theta), |a| = |sin(theta)|; sc = sin(theta) cos(theta)  (see trig_pair).
template <class R>
struct DerivU {
    R dr, dth, dph, dpr, dpth, dt, isig;
};
// POLAR = false drops the polar clamp: only for rays that provably stay away from the axis (see kPolarSafePph).
template <class R, bool WITH_T, bool WITH_PHI, bool POLAR = true>
__device__ __forceinline__ DerivU<R> rhs_ks_u(const HoleRay<R>& c, R r, R a, R sc, R pr, R pth) {
    using N = Num<R>;
    R sin2 = a * a;
    // Within 1e-6 rad of the polar axis (rare): clamp sin^2 and zero dH/dtheta below |sin| = 1e-10 (both of its
    // terms carry sc). f64: nested rare path behind an integer compare (ptxas predicates it). f32: two plain selects.
    if (!POLAR) {
    } else if (sizeof(R) == 8) {
        if (lt_pos(sin2, R(1e-12))) {
            sin2 = R(1e-12);
            if (abs_lt_pos(a, R(1e-10))) sc = R(0);
        }
    } else {
        if (sin2 < R(1e-20)) sc = R(0);
        sin2 = floorAt<R>(sin2, R(1e-12));
    }
    const R r2a2 = N::fma_(r, r, c.a2);
    const R sigma = N::fma_(-c.a2, sin2, r2a2);      // r^2 + a^2 cos^2, cos^2 = 1 - sin^2 (kerr.rs:418,449)
    const R delta = N::fma_(-c.twoM, r, r2a2);
    // one reciprocal serves both 1/Sigma and 1/sin^2 (Sigma sin^2 >= 1e-12 r^2: far from under/overflow)
    const R t = N::rcp(sigma * sin2);
    const R isig = t * sin2;
    const R w = t * sigma;                          // 1/sin^2
    const R A1 = c.pph2 * w;                        // pph^2 / sin^2
    const R pr2 = pr * pr;
    // dr/dlambda * Sigma = Delta pr + D0,  D0 = 2Mr pt + a pph.  Grouping N by powers of pr reuses it:
    //   N = 2Mr (2 pt pr - pt^2) + Delta pr^2 + pth^2 + A1 + 2 a pph pr = pr (dr + D0) + pth^2 + A1 - 2M pt^2 r
    // (4 FP64 ops after dr instead of 6: every instruction here is a slot on the pipe that bounds the kernel)
    const R D0 = N::fma_(c.twoM_pt, r, c.a_pph);
    DerivU<R> d;
    d.isig = isig;
    d.dr = N::fma_(delta, pr, D0);
    R Nn = N::fma_(-c.twoM_pt2, r, A1);
    Nn = N::fma_(pth, pth, Nn);
    Nn = N::fma_(pr, d.dr + D0, Nn);
    const R q = Nn * isig;
    // N_r / 2 = M (2 pt pr - pt^2) + (r - M) pr^2
    const R halfNr = N::fma_(r - c.M, pr2, N::fma_(c.M_two_pt, pr, -c.M_pt2));
    d.dth = pth;
    d.dpr = N::fma_(r, q, -halfNr);                          // -(N_r/2 - r N/Sigma)
    d.dpth = sc * N::fma_(-c.a2, q, A1 * w);                 // -(sc (a^2 N/Sigma - pph^2/sin^4))
    d.dph = WITH_PHI ? N::fma_(c.pph, w, c.a * pr) : R(0);
    d.dt = WITH_T ? N::fma_((c.twoM * r) * isig, pr - c.pt, -c.pt) : R(0);
    return d;
}
template <class R, bool WITH_T>
__device__ __forceinline__ Deriv<R> rhs_ks(const HoleRay<R>& c, R r, R a, R sc, R pr, R pth) {
    const DerivU<R> u = rhs_ks_u<R, WITH_T, true>(c, r, a, sc, pr, pth);
    Deriv<R> d;
    d.dr = u.dr * u.isig; d.dth = u.dth * u.isig; d.dph = u.dph * u.isig;
    d.dpr = u.dpr * u.isig; d.dpth = u.dpth * u.isig; d.dt = u.dt;
    return d;
}

// Boyer-Lindquist form (kerr.rs:266-372): H = (P/D + Q/Sigma)/2 with D = Delta Sigma,
//   P = -U pt^2 - 4 M r a pt pph - a^2 pph^2,  U = Sigma (r^2+a^2) + 2 M r a^2 sin^2,
//   Q = Delta pr^2 + pth^2 + pph^2/sin^2.   g^phph is gated to 0 when sin^2 < 1e-9 (kerr.rs:282-286).
template <class R, bool WITH_T>
__device__ __forceinline__ Deriv<R> rhs_bl(const HoleRay<R>& c, R r, R s, R cth, R pr, R pth) {
    using N = Num<R>;
    const R sin2 = s * s;
    const R cos2 = cth * cth;
    const R r2 = r * r;
    const R sigma = N::fma_(c.a2, cos2, r2);
    const R delta = N::fma_(-c.twoM, r, r2 + c.a2);
    const R ra2 = r2 + c.a2;
    const R D = delta * sigma;
    const R iD = N::rcp_ieee(D);
    const R isig = N::rcp_ieee(sigma);
    const R w = N::rcp_ieee(sin2);
    const R twoMr = c.twoM * r;
    const R U = N::fma_(twoMr * c.a2, sin2, sigma * ra2);
    const R cross = R(2) * twoMr * c.a * c.pt * c.pph;   // 4 M r a pt pph
    const R P = -(U * c.pt2) - cross - c.a2 * c.pph2;
    const R pr2 = pr * pr;
    const R A1 = c.pph2 * w;
    const R Q = N::fma_(delta, pr2, N::fma_(pth, pth, A1));
    const R sc = s * cth;
    // radial derivatives
    const R dDelta = R(2) * r - c.twoM;
    const R Ur = R(2) * r * ra2 + sigma * (R(2) * r) + c.twoM * c.a2 * sin2;
    const R Dr = dDelta * sigma + delta * (R(2) * r);
    const R Pr = -(Ur * c.pt2) - R(2) * c.twoM * c.a * c.pt * c.pph;
    const R Qr = dDelta * pr2;
    const R dHdr = R(0.5) * ((Pr * D - P * Dr) * (iD * iD) + (Qr * sigma - Q * (R(2) * r)) * (isig * isig));
    // polar derivatives
    const R Sth = R(-2) * c.a2 * sc;
    const R Uth = Sth * ra2 + twoMr * c.a2 * (R(2) * sc);
    const R Dth = delta * Sth;
    const R Pth = -(Uth * c.pt2);
    const R Qth = R(-2) * sc * A1 * w;
    const R dHdth = R(0.5) * ((Pth * D - P * Dth) * (iD * iD) + (Qth * sigma - Q * Sth) * (isig * isig));
    const R g_tph = -(twoMr * c.a) * iD;
    const R g_phph = (sin2 < R(1e-9)) ? R(0) : (delta - c.a2 * sin2) * iD * w;
    Deriv<R> d;
    d.dr = delta * isig * pr;
    d.dth = pth * isig;
    d.dph = g_tph * c.pt + g_phph * c.pph;
    d.dpr = -dHdr;
    d.dpth = -dHdth;
    if (WITH_T) d.dt = -(U * iD) * c.pt + g_tph * c.pph; else d.dt = R(0);
    return d;
}

template <class R, int COORDS, bool WITH_T>
__device__ __forceinline__ Deriv<R> rhs_at(const HoleRay<R>& c, R r, R th, R pr, R pth) {
    if (COORDS == 1) {
        R a, sc;
        trig_pair(*c.trig, th, a, sc);
        return rhs_ks<R, WITH_T>(c, r, a, sc, pr, pth);
    }
    R s, cth;
    Num<R>::sincos_(th, &s, &cth);
    return rhs_bl<R, WITH_T>(c, r, s, cth, pr, pth);
}

// The quadratic of invariants/renormalization.rs:13-45 and H of invariants/mod.rs:25-37 share the metric
// evaluation; `want` selects what to produce.
// KNOWN_SIN: `th` carries sin(theta) itself (the rotated zone of the f64 kernel has it at hand), no evaluation here.
template <class R, int COORDS, bool KNOWN_SIN = false>
__device__ __forceinline__ void null_quadratic(const HoleRay<R>& c, R r, R th, R pth, R& A, R& B, R& C) {
    using N = Num<R>;
    const R r2 = r * r;
    const R delta = N::fma_(-c.twoM, r, r2 + c.a2);
    if (COORDS == 1) {
        R a = th, sc;
        if (!KNOWN_SIN) trig_pair(*c.trig, th, a, sc);
        const R sin2 = floorAt<R>(a * a, R(1e-12));
        const R cos2 = R(1) - sin2;
        const R sigma = N::fma_(c.a2, cos2, r2);
        const R isig = N::rcp(sigma);
        const R twoMr = c.twoM * r;
        A = delta * isig;                                              // g^rr
        B = R(2) * isig * N::fma_(twoMr, c.pt, c.a_pph);                // 2 (g^tr pt + g^rph pph)
        // g^tt pt^2 + g^thth pth^2 + g^phph pph^2
        C = N::fma_(isig, N::fma_(-twoMr, c.pt2, N::fma_(pth, pth, c.pph2 * N::rcp(sin2))), -c.pt2);
    } else {
        R s, cth;
        N::sincos_(th, &s, &cth);
        const R sin2 = s * s;
        const R cos2 = cth * cth;
        const R sigma = N::fma_(c.a2, cos2, r2);
        const R isig = N::rcp_ieee(sigma);
        const R D = delta * sigma;
        const R iD = N::rcp_ieee(D);
        const R twoMr = c.twoM * r;
        const R U = N::fma_(twoMr * c.a2, sin2, sigma * (r2 + c.a2));
        const R g_tt = -(U * iD);
        const R g_tph = -(twoMr * c.a) * iD;
        const R g_phph = (sin2 < R(1e-9)) ? R(0) : (delta - c.a2 * sin2) * iD / sin2;
        A = delta * isig;
        B = R(0);                                                       // g^tr = g^rph = 0 in BL
        C = g_tt * c.pt2 + isig * pth * pth + g_phph * c.pph2 + R(2) * g_tph * c.pt * c.pph;
    }
}

// invariants/renormalization.rs:13-45
template <class R, int COORDS, bool KNOWN_SIN = false>
__device__ __forceinline__ R renormalize_pr(const HoleRay<R>& c, R r, R th, R pr, R pth) {
    using N = Num<R>;
    R A, B, C;
    null_quadratic<R, COORDS, KNOWN_SIN>(c, r, th, pth, A, B, C);
    if (N::abs_(A) > R(1e-12)) {
        const R disc = N::fma_(B, B, R(-4) * A * C);
        if (disc >= R(0)) {
            const R sq = (COORDS == 1) ? sqrt_nr(disc) : N::sqrt_(disc);
            const R inv2a = (COORDS == 1) ? N::rcp(R(2) * A) : N::rcp_ieee(R(2) * A);
            const R sol1 = (-B + sq) * inv2a;
            const R sol2 = (-B - sq) * inv2a;
            return (N::abs_(sol1 - pr) < N::abs_(sol2 - pr)) ? sol1 : sol2;
        }
    }
    return pr;
}

// The same root for the far, off-axis zone of the f64 trace kernel, from sin(theta) = s (at hand there), with the quadratic
// multiplied through by Sigma: Delta pr^2 + 2 D0 pr + C' = 0, D0 = 2 M r pt + a pph (the radial-velocity numerator of the
// RHS), C' = -2 M pt^2 r + pth^2 + pph^2/sin^2 - pt^2 Sigma. One reciprocal (1/(Delta sin^2) serves 1/Delta and 1/sin^2)
// instead of three, no trigonometric evaluation: 33 FP64 instructions instead of 81. Callers guarantee what the reference's
// guards test (|g^rr| > 1e-12, sin^2 far above the 1e-12 clamp): r beyond the disk, ray off the polar axis.
template <class R>
__device__ __forceinline__ R renormalize_pr_far(const HoleRay<R>& c, R r, R s, R pr, R pth) {
    using N = Num<R>;
    const R r2a2 = N::fma_(r, r, c.a2), sin2 = s * s;
    const R sigma = N::fma_(-c.a2, sin2, r2a2), delta = N::fma_(-c.twoM, r, r2a2);
    const R t = N::rcp(delta * sin2);
    const R inv_delta = t * sin2, w = t * delta;
    const R D0 = N::fma_(c.twoM_pt, r, c.a_pph);
    R Cp = N::fma_(-c.twoM_pt2, r, N::fma_(pth, pth, c.pph2 * w));
    Cp = N::fma_(-c.pt2, sigma, Cp);
    const R disc = N::fma_(D0, D0, -(delta * Cp));
    if (disc >= R(0)) {
        const R sq = sqrt_nr(disc);
        const R sol1 = (sq - D0) * inv_delta, sol2 = -(sq + D0) * inv_delta;
        return (N::abs_(sol1 - pr) < N::abs_(sol2 - pr)) ? sol1 : sol2;
    }
    return pr;
}

// invariants/mod.rs:25-37:  H = (A pr^2 + B pr + C)/2
template <class R, int COORDS>
__device__ __forceinline__ R hamiltonian_of(const HoleRay<R>& c, R r, R th, R pr, R pth) {
    R A, B, C;
    null_quadratic<R, COORDS>(c, r, th, pth, A, B, C);
    return R(0.5) * Num<R>::fma_(Num<R>::fma_(A, pr, B), pr, C);
}

// --------------------------------------------------------------------------------------------------
// Steppers
// --------------------------------------------------------------------------------------------------
// geodesic/integrator.rs:209-226 implicit midpoint: 2 fixed-point iterations + final evaluation.
// s_mid = 0.5 (s + (s + d h)) = s + d h/2.
template <class R, int COORDS, bool WITH_T, bool POLAR = true>
__device__ __forceinline__ void step_symplectic(const HoleRay<R>& c, Ray<R>& y, R h) {
    using N = Num<R>;
    const R hh = R(0.5) * h;
    if (COORDS == 1) {
        R a, sc;
        trig_pair(*c.trig, y.th, a, sc);
        DerivU<R> d = rhs_ks_u<R, false, false, POLAR>(c, y.r, a, sc, y.pr, y.pth);
        R f = hh * d.isig;
        R mr = N::fma_(d.dr, f, y.r), mth = N::fma_(d.dth, f, y.th);
        R mpr = N::fma_(d.dpr, f, y.pr), mpth = N::fma_(d.dpth, f, y.pth);
        trig_pair(*c.trig, mth, a, sc);
        d = rhs_ks_u<R, false, false, POLAR>(c, mr, a, sc, mpr, mpth);
        f = hh * d.isig;
        mr = N::fma_(d.dr, f, y.r); mth = N::fma_(d.dth, f, y.th);
        mpr = N::fma_(d.dpr, f, y.pr); mpth = N::fma_(d.dpth, f, y.pth);
        trig_pair(*c.trig, mth, a, sc);
        d = rhs_ks_u<R, WITH_T, WITH_T, POLAR>(c, mr, a, sc, mpr, mpth);
        f = h * d.isig;
        y.r = N::fma_(d.dr, f, y.r);
        y.th = N::fma_(d.dth, f, y.th);
        y.pr = N::fma_(d.dpr, f, y.pr);
        y.pth = N::fma_(d.dpth, f, y.pth);
        if (WITH_T) { y.ph = N::fma_(d.dph, f, y.ph); y.t = N::fma_(d.dt, h, y.t); }
        return;
    }
    Deriv<R> d = rhs_at<R, COORDS, false>(c, y.r, y.th, y.pr, y.pth);
    R mr = N::fma_(d.dr, hh, y.r), mth = N::fma_(d.dth, hh, y.th);
    R mpr = N::fma_(d.dpr, hh, y.pr), mpth = N::fma_(d.dpth, hh, y.pth);
    d = rhs_at<R, COORDS, false>(c, mr, mth, mpr, mpth);
    mr = N::fma_(d.dr, hh, y.r); mth = N::fma_(d.dth, hh, y.th);
    mpr = N::fma_(d.dpr, hh, y.pr); mpth = N::fma_(d.dpth, hh, y.pth);
    d = rhs_at<R, COORDS, WITH_T>(c, mr, mth, mpr, mpth);
    y.r = N::fma_(d.dr, h, y.r);
    y.th = N::fma_(d.dth, h, y.th);
    y.ph = N::fma_(d.dph, h, y.ph);
    y.pr = N::fma_(d.dpr, h, y.pr);
    y.pth = N::fma_(d.dpth, h, y.pth);
    if (WITH_T) y.t = N::fma_(d.dt, h, y.t);
}

// The implicit-midpoint step with NO full trigonometric evaluation: (s0, c0) = (sin, cos) of the step's own theta come in
// from the caller, the two predictor-shifted angles are rotations of them (trig_rot), and so is the pair of the NEW theta
// the step hands back -- theta moves by h p_theta / Sigma, as small as the shifts. The caller anchors (s0, c0) with one
// trig_full per 8-step chunk, so rounding of the chained rotations (~1 ulp each) never accumulates beyond a chunk.
// Callers guarantee |h p_theta / Sigma| <= 2^-8 for the whole chunk (trig_rot_small) and a ray off the polar axis
// (k_trace_tile's zone 3 of the f64 kernel); f64 only. theta itself is updated exactly as in step_symplectic.
template <bool WITH_T, class RS>
__device__ __forceinline__ void step_symplectic_rot(const HoleRay<RS>& c, Ray<RS>& y, RS h, double& s0, double& c0) {
    using N = Num<RS>;
    const RS hh = RS(0.5) * h;
    double s1, c1;
    DerivU<RS> d = rhs_ks_u<RS, false, false, false>(c, y.r, RS(s0), RS(s0 * c0), y.pr, y.pth);
    RS f = hh * d.isig;
    RS mr = N::fma_(d.dr, f, y.r), mpr = N::fma_(d.dpr, f, y.pr), mpth = N::fma_(d.dpth, f, y.pth);
    trig_rot_small(*c.trig, s0, c0, (double)(d.dth * f), s1, c1);
    d = rhs_ks_u<RS, false, false, false>(c, mr, RS(s1), RS(s1 * c1), mpr, mpth);
    f = hh * d.isig;
    mr = N::fma_(d.dr, f, y.r); mpr = N::fma_(d.dpr, f, y.pr); mpth = N::fma_(d.dpth, f, y.pth);
    trig_rot_small(*c.trig, s0, c0, (double)(d.dth * f), s1, c1);
    d = rhs_ks_u<RS, WITH_T, WITH_T, false>(c, mr, RS(s1), RS(s1 * c1), mpr, mpth);
    f = h * d.isig;
    y.r = N::fma_(d.dr, f, y.r);
    y.th = N::fma_(d.dth, f, y.th);
    y.pr = N::fma_(d.dpr, f, y.pr);
    y.pth = N::fma_(d.dpth, f, y.pth);
    if (WITH_T) { y.ph = N::fma_(d.dph, f, y.ph); y.t = N::fma_(d.dt, h, y.t); }
    trig_rot_small(*c.trig, s0, c0, (double)(d.dth * f), s1, c1);
    s0 = s1; c0 = c1;
}

// GVT_PRECISION_MIXED: the same implicit-midpoint step with its two fixed-point (predictor) evaluations in f32 and the
// final evaluation -- the only one whose value reaches the state directly -- in f64 on the f64 state.
// Why this is allowed: the predictors only locate the midpoint. An error e in a predictor's derivative moves the
// midpoint by (h/2) e and the final derivative by (h/2) J e, J = d(rhs)/d(state): far from the hole (r >~ 35 M:
// |J| ~ M/r^2, h <= 1) that factor is < 5e-4 for the second predictor and its square for the first, so f32's 6e-8
// reaches the state as <~ 3e-11 per step, on the ray's final outbound leg where nothing amplifies it afterwards.
// Measured with the CPU oracle on the headline frame (oracle/experiments/mixed_probe.cpp): worst RGBA component
// 1.3e-9 relative against all-f64 with r_switch = 35 M, against 1.6e-4 when the f32 predictors are used everywhere.
// `frozen` lanes (finished rays, h = 0) must not let an f32 overflow on their parked state turn 0 * inf into NaN.
// The predictors' trigonometry is MUFU.SIN / MUFU.COS (sin.approx.f32, |error| < 4e-7 on [-pi, pi]): the same damping
// argument covers it (measured on the headline frame: worst RGBA component 7e-10 with MUFU, 4e-10 with polynomial f32 trig).
// Callers guarantee a ray off the polar axis (kPolarSafePph), so neither precision carries the polar clamp here; hf is
// the predictors' copy of the step (0 for frozen lanes: their increments are then exactly 0, the f32 derivatives of a
// parked state being finite).
template <bool WITH_T, class RS>   // RS = double (a template parameter only so that the f32 instantiations of the caller compile)
__device__ __forceinline__ void step_symplectic_mixed(const HoleRay<RS>& c, const HoleRay<float>& cf, Ray<RS>& y, RS h, float hf) {
    const float rf = (float)y.r, thf = (float)y.th, prf = (float)y.pr, pthf = (float)y.pth;
    const float hh = 0.5f * hf;
    float sn = __sinf(thf), cs = __cosf(thf);
    DerivU<float> d = rhs_ks_u<float, false, false, false>(cf, rf, sn, sn * cs, prf, pthf);
    float f = hh * d.isig;
    const float mr = fmaf(d.dr, f, rf), mth = fmaf(d.dth, f, thf), mpr = fmaf(d.dpr, f, prf), mpth = fmaf(d.dpth, f, pthf);
    sn = __sinf(mth); cs = __cosf(mth);
    d = rhs_ks_u<float, false, false, false>(cf, mr, sn, sn * cs, mpr, mpth);
    f = hh * d.isig;
    // midpoint = f64 state + f32 increment; final evaluation and state update in f64
    using N = Num<RS>;
    RS a64, sc64;
    trig_pair(*c.trig, y.th + (RS)(d.dth * f), a64, sc64);
    const DerivU<RS> D = rhs_ks_u<RS, WITH_T, WITH_T, false>(c, y.r + (RS)(d.dr * f), a64, sc64, y.pr + (RS)(d.dpr * f), y.pth + (RS)(d.dpth * f));
    const RS F = h * D.isig;
    y.r = N::fma_(D.dr, F, y.r);
    y.th = N::fma_(D.dth, F, y.th);
    y.pr = N::fma_(D.dpr, F, y.pr);
    y.pth = N::fma_(D.dpth, F, y.pth);
    if (WITH_T) { y.ph = N::fma_(D.dph, F, y.ph); y.t = N::fma_(D.dt, h, y.t); }
}

// geodesic/integrator.rs:193-203 classic RK4
template <class R, int COORDS, bool WITH_T>
__device__ __forceinline__ void step_rk4(const HoleRay<R>& c, Ray<R>& y, R h) {
    using N = Num<R>;
    const R hh = R(0.5) * h;
    Deriv<R> k1 = rhs_at<R, COORDS, WITH_T>(c, y.r, y.th, y.pr, y.pth);
    Deriv<R> k2 = rhs_at<R, COORDS, WITH_T>(c, N::fma_(k1.dr, hh, y.r), N::fma_(k1.dth, hh, y.th),
                                            N::fma_(k1.dpr, hh, y.pr), N::fma_(k1.dpth, hh, y.pth));
    Deriv<R> k3 = rhs_at<R, COORDS, WITH_T>(c, N::fma_(k2.dr, hh, y.r), N::fma_(k2.dth, hh, y.th),
                                            N::fma_(k2.dpr, hh, y.pr), N::fma_(k2.dpth, hh, y.pth));
    Deriv<R> k4 = rhs_at<R, COORDS, WITH_T>(c, N::fma_(k3.dr, h, y.r), N::fma_(k3.dth, h, y.th),
                                            N::fma_(k3.dpr, h, y.pr), N::fma_(k3.dpth, h, y.pth));
    const R h6 = h / R(6);
    y.r += h6 * (k1.dr + R(2) * k2.dr + R(2) * k3.dr + k4.dr);
    y.th += h6 * (k1.dth + R(2) * k2.dth + R(2) * k3.dth + k4.dth);
    y.ph += h6 * (k1.dph + R(2) * k2.dph + R(2) * k3.dph + k4.dph);
    y.pr += h6 * (k1.dpr + R(2) * k2.dpr + R(2) * k3.dpr + k4.dpr);
    y.pth += h6 * (k1.dpth + R(2) * k2.dpth + R(2) * k3.dpth + k4.dpth);
    if (WITH_T) y.t += h6 * (k1.dt + R(2) * k2.dt + R(2) * k3.dt + k4.dt);
}

// geodesic/integrator.rs:113-190 Fehlberg 4(5) attempt: returns the 5th-order state in `out` and the error
// estimate = max over the POSITION components (t, r, theta, phi) of |h * sum (b5-b4)_j k_j|.
// The t-row of the error needs dt at every stage, so dt is always evaluated here (it is 3 flops).
template <class R, int COORDS>
__device__ __forceinline__ R rkf45_attempt(const HoleRay<R>& c, const Ray<R>& y, R h, Ray<R>& out) {
    using N = Num<R>;
#define GVT_STAGE(K, EXPR_R, EXPR_TH, EXPR_PR, EXPR_PTH) \
    const Deriv<R> K = rhs_at<R, COORDS, true>(c, y.r + h * (EXPR_R), y.th + h * (EXPR_TH), y.pr + h * (EXPR_PR), y.pth + h * (EXPR_PTH))
    const Deriv<R> k1 = rhs_at<R, COORDS, true>(c, y.r, y.th, y.pr, y.pth);
    const R a21 = R(1.0 / 4.0);
    GVT_STAGE(k2, a21 * k1.dr, a21 * k1.dth, a21 * k1.dpr, a21 * k1.dpth);
    const R a31 = R(3.0 / 32.0), a32 = R(9.0 / 32.0);
    GVT_STAGE(k3, a31 * k1.dr + a32 * k2.dr, a31 * k1.dth + a32 * k2.dth, a31 * k1.dpr + a32 * k2.dpr,
              a31 * k1.dpth + a32 * k2.dpth);
    const R a41 = R(1932.0 / 2197.0), a42 = R(-7200.0 / 2197.0), a43 = R(7296.0 / 2197.0);
    GVT_STAGE(k4, a41 * k1.dr + a42 * k2.dr + a43 * k3.dr, a41 * k1.dth + a42 * k2.dth + a43 * k3.dth,
              a41 * k1.dpr + a42 * k2.dpr + a43 * k3.dpr, a41 * k1.dpth + a42 * k2.dpth + a43 * k3.dpth);
    const R a51 = R(439.0 / 216.0), a52 = R(-8.0), a53 = R(3680.0 / 513.0), a54 = R(-845.0 / 4104.0);
    GVT_STAGE(k5, a51 * k1.dr + a52 * k2.dr + a53 * k3.dr + a54 * k4.dr,
              a51 * k1.dth + a52 * k2.dth + a53 * k3.dth + a54 * k4.dth,
              a51 * k1.dpr + a52 * k2.dpr + a53 * k3.dpr + a54 * k4.dpr,
              a51 * k1.dpth + a52 * k2.dpth + a53 * k3.dpth + a54 * k4.dpth);
    const R a61 = R(-8.0 / 27.0), a62 = R(2.0), a63 = R(-3544.0 / 2565.0), a64 = R(1859.0 / 4104.0),
            a65 = R(-11.0 / 40.0);
    GVT_STAGE(k6, a61 * k1.dr + a62 * k2.dr + a63 * k3.dr + a64 * k4.dr + a65 * k5.dr,
              a61 * k1.dth + a62 * k2.dth + a63 * k3.dth + a64 * k4.dth + a65 * k5.dth,
              a61 * k1.dpr + a62 * k2.dpr + a63 * k3.dpr + a64 * k4.dpr + a65 * k5.dpr,
              a61 * k1.dpth + a62 * k2.dpth + a63 * k3.dpth + a64 * k4.dpth + a65 * k5.dpth);
#undef GVT_STAGE
    const R b1 = R(16.0 / 135.0), b3 = R(6656.0 / 12825.0), b4 = R(28561.0 / 56430.0), b5 = R(-9.0 / 50.0),
            b6 = R(2.0 / 55.0);
#define GVT_B5(F) (b1 * k1.F + b3 * k3.F + b4 * k4.F + b5 * k5.F + b6 * k6.F)
    out.t = y.t + h * GVT_B5(dt);
    out.r = y.r + h * GVT_B5(dr);
    out.th = y.th + h * GVT_B5(dth);
    out.ph = y.ph + h * GVT_B5(dph);
    out.pr = y.pr + h * GVT_B5(dpr);
    out.pth = y.pth + h * GVT_B5(dpth);
#undef GVT_B5
    const R e1 = R(16.0 / 135.0 - 25.0 / 216.0), e3 = R(6656.0 / 12825.0 - 1408.0 / 2565.0),
            e4 = R(28561.0 / 56430.0 - 2197.0 / 4104.0), e5 = R(-9.0 / 50.0 + 1.0 / 5.0), e6 = R(2.0 / 55.0);
#define GVT_E(F) N::abs_(h * (e1 * k1.F + e3 * k3.F + e4 * k4.F + e5 * k5.F + e6 * k6.F))
    R err = N::max_(N::max_(GVT_E(dt), GVT_E(dr)), N::max_(GVT_E(dth), GVT_E(dph)));
#undef GVT_E
    return err;
}

// geodesic/integrator.rs:53-108 AdaptiveStepper::step (safety 0.9, min 1e-5, max 10). Updates y, returns the
// next step; *evals counts RHS evaluations (6 per attempt).
template <class R, int COORDS>
__device__ __forceinline__ R adaptive_step(const HoleRay<R>& c, Ray<R>& y, R h_try, R tol, uint32_t& evals) {
    using N = Num<R>;
    const R max_step = R(10), min_step = R(1e-5), safety = R(0.9);
    R h = clampR<R>(h_try, -max_step, max_step);
    const R inv_tol = Num<R>::rcp_ieee(tol);
    for (;;) {
        Ray<R> ny;
        const R err = rkf45_attempt<R, COORDS>(c, y, h, ny);
        evals += 6;
        const R ratio = (err == R(0)) ? R(0) : err * inv_tol;
        if (ratio <= R(1)) {
            y = ny;
            const R growth = (ratio < R(1e-4)) ? R(5) : safety * pow_m02(ratio);
            return clampR<R>(h * N::min_(growth, R(5)), -max_step, max_step);
        }
        const R shrink = safety * pow_m025(ratio);
        h *= N::max_(shrink, R(0.1));
        if (N::abs_(h) < min_step) {
            const R hs = (h < R(0)) ? -min_step : min_step;
            (void)rkf45_attempt<R, COORDS>(c, y, hs, ny);
            evals += 6;
            y = ny;
            return hs;
        }
        // A NaN step size would make the reference loop forever (|NaN| < min is false); a kernel must not.
        if (!(h == h)) { y = ny; return min_step; }
    }
}

// The same controller for a converged warp (the trace kernel): every lane executes every attempt, lanes that are
// finished or already accepted step with h = 0 (their state does not move), and the retry loop runs while any lane
// still has a rejected step. In SIMT a rejection on one lane costs the whole warp another pass anyway; keeping the
// attempt out of divergent control flow lets ptxas feed the Butcher coefficients from uniform registers.
template <class R, int COORDS>
__device__ __forceinline__ void adaptive_step_warp(const HoleRay<R>& c, Ray<R>& y, R& h, R tol, bool active,
                                                   uint32_t& evals) {
    using N = Num<R>;
    const R max_step = R(10), min_step = R(1e-5), safety = R(0.9);
    const R inv_tol = Num<R>::rcp_ieee(tol);
    R ht = clampR<R>(h, -max_step, max_step);
    bool pending = active, forced = false;
    while (__any_sync(0xffffffffu, pending)) {
        Ray<R> ny;
        const R err = rkf45_attempt<R, COORDS>(c, y, pending ? ht : R(0), ny);
        if (pending) {
            evals += 6;
            const R ratio = (err == R(0)) ? R(0) : err * inv_tol;
            if (forced || ratio <= R(1)) {
                y = ny;
                const R growth = (ratio < R(1e-4)) ? R(5) : safety * pow_m02(ratio);
                h = forced ? ht : clampR<R>(ht * N::min_(growth, R(5)), -max_step, max_step);
                pending = false;
            } else {
                ht *= N::max_(safety * pow_m025(ratio), R(0.1));
                if (N::abs_(ht) < min_step) { ht = (ht < R(0)) ? -min_step : min_step; forced = true; }
                else if (!(ht == ht)) { ht = min_step; forced = true; }   // NaN step: take the forced minimum step
            }
        }
    }
}

// physics/redshift.rs:65-95 kerr_g_factor (sm = sqrt(mass), evaluated once on the host)
template <class R>
__device__ __forceinline__ R g_factor(R r, R mass, R sm, R spin, R lambda) {
    using N = Num<R>;
    const R a = spin * mass;
    const R r2 = r * r;
    const R omega = sm * N::rcp(N::fma_(r, sqrt_nr(r), a * sm));     // sqrt(M) / (r^1.5 + a sqrt(M))
    const R twoM_r = R(2) * mass * N::rcp(r);                         // 2 M r / Sigma at the equator (Sigma = r^2)
    const R g_tt = -(R(1) - twoM_r);
    const R g_tphi = -(twoM_r * a);
    const R g_phiphi = r2 + a * a + twoM_r * a * a;
    const R ut_denom = -g_tt - R(2) * omega * g_tphi - omega * omega * g_phiphi;
    if (ut_denom <= R(0)) return R(0);
    const R factor = R(1) - lambda * omega;
    if (N::abs_(factor) < R(1e-30)) return R(0);
    return sqrt_nr(ut_denom) * N::rcp(factor);                        // 1 / (u^t (1 - lambda Omega))
}

// --------------------------------------------------------------------------------------------------
// GLSL-semantics path: Cartesian Velocity-Verlet on the production fragment shader's pseudo-Kerr
// acceleration field (src/shaders/blackhole/chunks/metric.ts:96-149, fragment.glsl.ts:129-221).
// --------------------------------------------------------------------------------------------------
template <class R>
struct Vec3 {
    R x, y, z;
};
template <class R> __device__ __forceinline__ R dot3(const Vec3<R>& a, const Vec3<R>& b) {
    return Num<R>::fma_(a.x, b.x, Num<R>::fma_(a.y, b.y, a.z * b.z));
}

// kerr_geodesic_accel: Darwin-type radial pull scaled by r^2/Sigma with the spin-coupled L_eff^2, plus the
// gravito-magnetic term (0,1,0) x v. Returns the acceleration and the ZAMO angular velocity.
template <class R>
__device__ __forceinline__ void glsl_accel(const Vec3<R>& p, const Vec3<R>& v, R M, R a, Vec3<R>& acc, R& omega) {
    using N = Num<R>;
    const R a2 = a * a;
    const R rho2 = dot3(p, p);
    const R diff = rho2 - a2;
    const R disc = N::fma_(diff, diff, R(4) * a2 * p.y * p.y);
    const R r2 = R(0.5) * (diff + sqrt_nr(N::max_(R(0), disc)));
    const R r2c = N::max_(R(1e-8), r2);
    const R r_k = sqrt_nr(r2c);
    const R sigma = N::fma_(a2, p.y * p.y * N::rcp(r2c), r2);
    const R Lx = p.y * v.z - p.z * v.y, Ly = p.z * v.x - p.x * v.z, Lz = p.x * v.y - p.y * v.x;
    const R Ly_eff = Ly - a;
    const R L2_eff = N::fma_(Ly_eff, Ly_eff, N::fma_(Lx, Lx, Lz * Lz));          // Ly_eff^2 + (L.L - Ly^2)
    const R r2_inv = N::rcp(r_k * r_k);
    const R sigma_ratio = r2 * N::rcp(N::max_(R(1e-8), sigma));
    const R f = M * r2_inv * sigma_ratio * N::fma_(R(3) * N::max_(R(0), L2_eff), r2_inv, R(1));
    const R k = -f * N::rcp(sqrt_nr(rho2));                                       // r_hat = -p/|p|
    omega = R(2) * M * a * N::rcp(N::max_(R(1e-8), r_k * (r2 + a2)));
    acc.x = N::fma_(p.x, k, v.z * omega);
    acc.y = p.y * k;
    acc.z = N::fma_(p.z, k, -v.x * omega);
}

template <class R> __device__ __forceinline__ R smoothstep_glsl(R e0, R e1, R x) {
    const R t = clampR<R>((x - e0) * Num<R>::rcp_ieee(e1 - e0), R(0), R(1));
    return t * t * (R(3) - R(2) * t);
}

// One iteration's step size (fragment.glsl.ts:141-162).
template <class R>
__device__ __forceinline__ R glsl_step_size(R r, R py_abs, R rh, R rph) {
    using N = Num<R>;
    const R MIN_STEP = R(0.01), MAX_STEP = R(1.2);
    const R distFactor = N::fma_(r, R(0.05), R(1));
    R dt = clampR<R>((r - rh) * R(0.1) * distFactor, MIN_STEP, MAX_STEP * distFactor);
    if (r > R(30)) {
        dt = N::max_(dt, N::fma_(r - R(30), R(0.08), MIN_STEP));
        dt = N::min_(dt, MAX_STEP * R(2.5));
    }
    dt = N::min_(dt, N::fma_(N::abs_(r - rph), R(0.15), MIN_STEP));
    const R t = clampR<R>((py_abs - R(0.2)) * R(-5), R(0), R(1));                 // smoothstep(0.2, 0.0, |y|)
    const R hRef = t * t * (R(3) - R(2) * t);
    return dt * N::fma_(hRef, R(-0.7), R(1));
}

}  // namespace gvt
