// gvt_kernels.cu — hand-written sm_100a kernels of the Kerr geodesic hot path.
//
//   k_trace_tile      fused per-pixel kernel: camera -> (x,p) (compute.wgsl.ts:159-187), null renormalisation,
//                     march (geodesic/mod.rs:180-253 with step_symplectic / RK4 / adaptive RKF45), thin-disk
//                     crossing shade through the GR g-factor + Planckian redshift LUT, float4 RGBA store.
//                     Persistent CTAs; each warp pulls 8x4-pixel tiles from an atomic queue; the camera block +
//                     disk LUT and the spectral LUT are staged into shared memory by TMA bulk copies
//                     (cp.async.bulk + mbarrier); the ray state lives in registers; the LUT barrier is waited on
//                     after the first tile's ray generation, so the copy overlaps that set-up work.
//   k_integrate_rays  geodesic::integrate over a batch of explicit initial states (the PhysicsEngine seam).
//   k_taa_resolve     YCoCg variance-clip TAA resolve (ataa.wgsl.ts:28-83), 3x3 moments by warp shuffles.
//   k_fma_peak        FFMA / DFMA micro-benchmark: the FP32 / FP64 roofline denominators.
#include "gvt_device.cuh"
#include "gvt_internal.h"
#include <cuda_fp16.h>

namespace gvt {

// --------------------------------------------------------------------------------------------------
// TMA bulk copy + mbarrier (PTX, sm_90+/sm_100a). SASS: UBLKCP + SYNCS.
// --------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// --------------------------------------------------------------------------------------------------
// Frame-buffer pixel formats. The frame chain (cur / frame / hist) is RGBA32F -- the parity format -- or RGBA16F, the
// reference's own texture format (rendering/reprojection.ts:120-140, webgpu/renderer.ts:161-180): half the HBM bytes of
// the TAA and bloom passes, no conversion pass for hosts that want RGBA16F.
// --------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint2 pack_half4(const float4& v) {
    const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&lo); o.y = *reinterpret_cast<const uint32_t*>(&hi);
    return o;
}
__device__ __forceinline__ float4 unpack_half4(const uint2& v) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void store_px(float4* base, size_t idx, const float4& v, bool f16) {
    if (f16) reinterpret_cast<uint2*>(base)[idx] = pack_half4(v);
    else base[idx] = v;
}
__device__ __forceinline__ float4 load_px(const float4* base, size_t idx, bool f16) {
    return f16 ? unpack_half4(reinterpret_cast<const uint2*>(base)[idx]) : base[idx];
}

// --------------------------------------------------------------------------------------------------
// LUT sampling (the two joints the reference never wired in f64; the sampling rule is specified in DESIGN.md)
// --------------------------------------------------------------------------------------------------
template <class R>
__device__ __forceinline__ R sample_tdisk(const float* tdisk, uint32_t n, R rin, R scale, R r) {
    using N = Num<R>;
    R t = clampR<R>((r - rin) * scale, R(0), R((double)(n - 1)));
    const R fl = N::floor_(t);
    const uint32_t i0 = (uint32_t)fl;
    const uint32_t i1 = min(i0 + 1u, n - 1u);
    const R f = t - fl;
    const R a = R(tdisk[i0]), b = R(tdisk[i1]);
    return N::fma_(b - a, f, a);
}

// bilinear, clamp-to-edge, texel-centre (GL LINEAR): x = u W - 0.5 (rendering/spectral.ts:52-54)
template <class R>
__device__ __forceinline__ void sample_spectrum(const float4* lut, uint32_t W, uint32_t H, R u, R v, R rgb[3]) {
    using N = Num<R>;
    R x = clampR<R>(N::fma_(u, R((double)W), R(-0.5)), R(0), R((double)(W - 1)));
    R y = clampR<R>(N::fma_(v, R((double)H), R(-0.5)), R(0), R((double)(H - 1)));
    const R xf = N::floor_(x), yf = N::floor_(y);
    const uint32_t x0 = (uint32_t)xf, y0 = (uint32_t)yf;
    const uint32_t x1 = min(x0 + 1u, W - 1u), y1 = min(y0 + 1u, H - 1u);
    const R fx = x - xf, fy = y - yf;
    const float4 t00 = lut[(size_t)y0 * W + x0], t10 = lut[(size_t)y0 * W + x1];
    const float4 t01 = lut[(size_t)y1 * W + x0], t11 = lut[(size_t)y1 * W + x1];
    const float a00[3] = {t00.x, t00.y, t00.z}, a10[3] = {t10.x, t10.y, t10.z};
    const float a01[3] = {t01.x, t01.y, t01.z}, a11[3] = {t11.x, t11.y, t11.z};
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const R top = N::fma_(R(a10[c]) - R(a00[c]), fx, R(a00[c]));
        const R bot = N::fma_(R(a11[c]) - R(a01[c]), fx, R(a01[c]));
        rgb[c] = N::fma_(bot - top, fy, top);
    }
}

// --------------------------------------------------------------------------------------------------
// The fused trace kernel
// --------------------------------------------------------------------------------------------------
#ifndef GVT_MAXT_F64
#define GVT_MAXT_F64 512   // threads per persistent CTA (one CTA per SM) for the fixed-step f64 instantiations
#endif
#ifndef GVT_MAXT_RKF
#define GVT_MAXT_RKF 512
#endif
#ifndef GVT_MAXT_F32
#define GVT_MAXT_F32 512
#endif
#ifndef GVT_UNROLL_NEAR
#define GVT_UNROLL_NEAR 4          // step-loop unroll of RK4
#endif
#ifndef GVT_UNROLL_SYMP
#define GVT_UNROLL_SYMP 1          // step-loop unroll of the implicit-midpoint zones (measured at 4K x 512, one loop: unroll 1 / 2 /
#endif                             // 4 / 8 -> 51.60 / 51.72 / 51.59 / 53.93 ms; with one loop per zone the small body is kinder to the I-cache)
#ifndef GVT_UNROLL_NEAR_MIXED
#define GVT_UNROLL_NEAR_MIXED 1    // the same in the GVT_PRECISION_MIXED instantiation: its two step loops share the instruction
                                   // cache (measured at 4K x 512: unroll 4/2 -> 51.0 ms with no_instruction stalls, 1/1 -> 44.4 ms)
#endif
#ifndef GVT_UNROLL_FAR
#define GVT_UNROLL_FAR 1           // unroll of the f32-predictor step loop
#endif
constexpr int kUnrollSymp = GVT_UNROLL_SYMP, kUnrollNear = GVT_UNROLL_NEAR, kUnrollNearMixed = GVT_UNROLL_NEAR_MIXED, kUnrollFar = GVT_UNROLL_FAR;

// Compile-time flavour of one march step (the step loop is instantiated once per flavour a kernel needs and the warp
// picks one per 8-step chunk): FAR = f32 predictors (GVT_PRECISION_MIXED), HCONST = the step rule has saturated and
// neither termination radius is within reach, so h is a constant and the radius tests are dropped, POLAR = the ray
// may come within reach of the polar clamp (kerr.rs:417,448,494) and the RHS carries it.
template <bool FAR, bool HCONST, bool POLAR, bool ROT = false, bool NODISK = false> struct StepKind {
    static constexpr bool far = FAR, hconst = HCONST, polar = POLAR, rot = ROT;   // ROT: shifted angles by rotation (trig_rot)
    static constexpr bool nodisk = NODISK;   // the whole chunk stays beyond the disk's outer edge: no equatorial-crossing test
};
// A ray whose conserved L_z = p_phi satisfies L^2 > kPolarSafe (Q + a^2 + L^2) cannot approach the axis: the polar
// potential Theta = Q + a^2 cos^2 - L^2 cot^2 >= 0 gives sin^2(theta) >= L^2 / (Q + a^2 + L^2) > 1e-4 along the whole
// geodesic, eight orders of magnitude above the 1e-12 clamp, so the clamp code is dead for it. Tiles holding any other
// ray (a few pixel columns around the image of the spin axis) run the generic step.
constexpr double kPolarSafe = 1e-4;

// Which side of the equatorial plane: sign of (theta - pi/2) as -1 / 0 / +1. The subtraction is sign-exact in IEEE
// arithmetic, so comparing the bit patterns (pi/2 > 0: sign-magnitude order, see lt_pos) gives the same answer with
// two integer compares on the idle ALU pipe instead of a DADD on the FP64 pipe.
__device__ __forceinline__ int hiword(double x) { return __double2hiint(x); }
__device__ __forceinline__ int hiword(float x) { return __float_as_int(x); }
__device__ __forceinline__ int equator_side(double th) {
    const long long b = __double_as_longlong(th), h = __double_as_longlong(1.5707963267948966);
    return (int)(b > h) - (int)(b < h);
}
__device__ __forceinline__ int equator_side(float th) { return (int)(th > 1.5707963267948966f) - (int)(th < 1.5707963267948966f); }
// one warp = one 8x4 pixel tile (compute.wgsl.ts:147 uses 8x8 groups)

template <class R, int METHOD, bool BUDGET, bool DEBUG, int MAXT, bool WGSL_RULE, bool MIXED = false>
__global__ void __launch_bounds__(MAXT, 1) k_trace_tile(const __grid_constant__ FrameParams P) {
    static_assert(!MIXED || (sizeof(R) == 8 && METHOD == 2), "GVT_PRECISION_MIXED: f64 state, implicit-midpoint stepper");
    using N = Num<R>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars[2];
    FrameBlock* fb = reinterpret_cast<FrameBlock*>(smem_raw);
    float4* lut_s = reinterpret_cast<float4*>(smem_raw + ((sizeof(FrameBlock) + 127) / 128) * 128);
    const uint32_t lut_bytes = P.spec_w * P.spec_h * (uint32_t)sizeof(float4);

    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bars[0], (uint32_t)sizeof(FrameBlock));
        tma_bulk_g2s(fb, P.block, (uint32_t)sizeof(FrameBlock), &bars[0]);
        if (P.lut_in_smem) {
            mbar_expect_tx(&bars[1], lut_bytes);
            // one bulk copy per <=64 KB chunk keeps each request well inside the tx-count range
            for (uint32_t off = 0; off < lut_bytes; off += 65536u) {
                const uint32_t n = min(65536u, lut_bytes - off);
                tma_bulk_g2s(reinterpret_cast<unsigned char*>(lut_s) + off,
                             reinterpret_cast<const unsigned char*>(P.spectrum) + off, n, &bars[1]);
            }
        }
    }
    mbar_wait(&bars[0], 0);  // camera block + disk LUT are needed for ray generation
    bool lut_ready = !P.lut_in_smem;
    const float4* lut = P.lut_in_smem ? lut_s : P.spectrum;

    const uint32_t lane = threadIdx.x & 31u;
    // One warp = one tile of 32 pixels: 8x4 (compute.wgsl.ts:147 uses 8x8 groups), or 16x2 / 32x1 when the launch's row
    // count is not a multiple of 4 -- a rank's 270-row block of a 2160-row frame would otherwise end in a half-empty
    // tile row (0.7 % of the launch).
    const uint32_t tw_log2 = P.tile_w_log2, TW = 1u << tw_log2, TH = 32u >> tw_log2;
    const uint32_t lx = lane & (TW - 1u), ly = lane >> tw_log2;
    const uint32_t tiles_x = (P.nx + TW - 1u) >> tw_log2, tiles_y = (P.ny + TH - 1u) / TH;
    const uint32_t n_tiles = tiles_x * tiles_y;

    HoleRay<R> hc;
    hc.trig = &P.trig;
    hc.set_hole(R(P.M), R(P.a));
    // f64: the derived hole constants straight from the parameter block (host-computed, the same IEEE products): a value
    // COMPUTED in the kernel lives in vector registers, one READ from the constant bank reaches the step loops' DFMAs as a
    // uniform-register operand -- a two-register DFMA (2 issue cycles) where a three-register one takes 3
    if (sizeof(R) == 8) { hc.a2 = R(P.a2); hc.twoM = R(P.twoM); }
    const R r_term = R(P.r_term), escape_r = R(P.escape_r), rh = R(P.rh);
    const R half_pi = R(1.5707963267948966);
    const R hs_bias = R(-0.15) * rh;              // step rule compute.wgsl.ts:213 as one FMA: 0.15 r - 0.15 r+
    HoleRay<float> hcf;                           // GVT_PRECISION_MIXED: the predictor evaluations' f32 constants
    if (MIXED) { hcf.trig = &P.trig; hcf.M = P.f32_M; hcf.a = P.f32_a; hcf.a2 = P.f32_a2; hcf.twoM = P.f32_twoM; }

    // per-warp census accumulators live in shared memory (lane 0 owns its warp's row): seven values that are touched
    // once per tile must not hold ten registers across the march
    __shared__ unsigned long long s_acc[MAXT / 32][3];
    __shared__ uint32_t s_term[MAXT / 32][4];
    const uint32_t wid = threadIdx.x >> 5;
    if ((threadIdx.x & 31u) == 0) {
        s_acc[wid][0] = s_acc[wid][1] = s_acc[wid][2] = 0ull;
        s_term[wid][0] = s_term[wid][1] = s_term[wid][2] = s_term[wid][3] = 0u;
    }

    // Nothing that is only needed before or after a tile's march may stay in registers across it: the step loops sit at the
    // 128-register cap, and every extra live value there evicts FP64 constants from the uniform registers (measured on the
    // SASS: +21 LDC per step for one more live loop bound). So the tile number is parked in shared memory and the pixel
    // coordinates are recomputed from it in the epilogue (`locate`), and the timeline diagnostics live there too.
    __shared__ uint32_t s_tile[MAXT / 32];
    __shared__ unsigned long long s_tstart[MAXT / 32];
    __shared__ uint32_t s_ntiles[MAXT / 32];
    if (P.timeline && lane == 0) {
        unsigned long long t0;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        s_tstart[wid] = t0; s_ntiles[wid] = 0u;
    }
    // Work distribution. Natural termination: one global atomic queue of warp-tiles (rows differ widely in cost). Budget
    // accounting: every tile costs nearly the same, so each CTA owns an equal share of the tiles -- dealt round-robin
    // (tile = CTA + k * grid), because the zones of the march do make tiles near the hole and on the polar columns a few
    // per cent dearer and a contiguous share would concentrate them -- and hands them to its warps through a
    // shared-memory counter. All SMs then finish together, where the global queue ends with the SMs that drew the last
    // tiles still busy (per-warp timelines: scripts/timeline_probe.py).
    __shared__ uint32_t s_next_tile;
    if (BUDGET) {
        if (threadIdx.x == 0) s_next_tile = 0u;
        __syncthreads();
    }
    for (;;) {
        uint32_t tile = 0;
        if (lane == 0) tile = BUDGET ? blockIdx.x + atomicAdd(&s_next_tile, 1u) * gridDim.x : atomicAdd(&P.counters->tile_counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= n_tiles) break;
        if (P.timeline && lane == 0) s_ntiles[wid]++;
        // tile -> this lane's pixel: here for ray generation and AGAIN in the epilogue, from the tile number parked in shared memory
        auto locate = [&](uint32_t tl, uint32_t& li, uint32_t& lj, uint32_t& px, uint32_t& py) -> bool {
            const uint32_t ti = tl % tiles_x, tj = tl / tiles_x;
            li = ti * TW + lx; lj = tj * TH + ly;  // lattice coordinates
            bool ok = (li < P.nx) && (lj < P.ny);
            px = P.x0 + min(li, P.nx - 1) * P.xs;
            py = P.y0 + min(lj, P.ny - 1) * P.ys;
            if (P.stripe.s) {   // GVT_FLAG_ROW_INTERLEAVE: this rank's stripes (+ TAA halo rows); rows off the frame are skipped
                uint32_t row;
                ok = stripe_row(P.stripe, min(lj, P.ny - 1), P.height, row) && ok;
                py = min(row, P.height - 1);
            }
            return ok;
        };
        if (lane == 0) s_tile[wid] = tile;
        uint32_t px, py;
        { uint32_t li_, lj_; (void)locate(tile, li_, lj_, px, py); }

        // ---- camera -> (x, p): compute.wgsl.ts:159-187 ----
        Ray<R> y;
        Vec3<R> wdir;          // world-space ray direction (the Cartesian march of METHOD 3 starts from it)
        bool polar_ray = true, rot_ray = false;
        float rot_rsmall = 0.0f;   // f64 zone 3 (rotated trigonometry) opens to this ray beyond this radius
        {
            const R ndcx = N::fma_(N::fma_(R((double)px), R(fb->inv_width), R(fb->jx)), R(2), R(-1));
            const R ndcy = N::fma_(N::fma_(R((double)py), R(fb->inv_height), R(fb->jy)), R(2), R(-1));
            const R cx = ndcx, cy = -ndcy;  // clip = (ndc.x, -ndc.y, 1, 1)
            R vt[4];
#pragma unroll
            for (int row = 0; row < 4; row++)
                vt[row] = R(fb->inv_proj[row]) * cx + R(fb->inv_proj[4 + row]) * cy + R(fb->inv_proj[8 + row]) +
                          R(fb->inv_proj[12 + row]);
            const R iw = N::rcp(vt[3]);
            R vx = vt[0] * iw, vy = vt[1] * iw, vz = vt[2] * iw;
            const R ivn = N::rcp(sqrt_nr(vx * vx + vy * vy + vz * vz));
            vx *= ivn; vy *= ivn; vz *= ivn;
            R w[3];
#pragma unroll
            for (int row = 0; row < 3; row++)
                w[row] = R(fb->inv_view[row]) * vx + R(fb->inv_view[4 + row]) * vy + R(fb->inv_view[8 + row]) * vz;
            const R iwn = N::rcp(sqrt_nr(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]));
            const R dx = w[0] * iwn, dy = w[1] * iwn, dz = w[2] * iwn;
            wdir.x = dx; wdir.y = dy; wdir.z = dz;
            const R st = R(fb->st), ct = R(fb->ct), sp = R(fb->sp), cp = R(fb->cp), r0 = R(fb->r0);
            const R pr_far = dx * (st * cp) + dy * ct + dz * (st * sp);
            const R inv_r0 = N::rcp(r0);
            const R pth_far = (dx * (ct * cp) - dy * st + dz * (ct * sp)) * inv_r0;
            const R pph_far = (dz * cp - dx * sp) * N::rcp(r0 * R(fb->safe_st));
            y.t = R(0); y.r = r0; y.th = R(fb->theta0); y.ph = R(fb->phi0);
            y.pr = pr_far;
            y.pth = pth_far * r0 * r0;
            hc.set_ray(R(-1), pph_far * r0 * r0 * st * st);
            // p_t = -1 for every camera ray (E = 1): its products with the hole constants are the constants themselves up to
            // sign -- exact identities (scalings by 1, 2), so nothing changes numerically -- and named this way they stay
            // constant-bank operands instead of per-thread register values (see set_hole above)
            hc.twoM_pt = -hc.twoM; hc.twoM_pt2 = hc.twoM; hc.M_two_pt = -hc.twoM; hc.M_pt2 = hc.M;
            if (MIXED) hcf.set_ray(-1.0f, (float)hc.pph);
            // Carter constant of the ray at the camera, for the polar-safety test (kPolarSafe)
            const R ct2 = ct * ct;
            const R Q = N::fma_(y.pth, y.pth, ct2 * N::fma_(hc.pph2, N::rcp(R(fb->safe_st) * R(fb->safe_st)), -hc.a2));
            polar_ray = !(hc.pph2 > R(kPolarSafe) * (Q + hc.a2 + hc.pph2));
            // Zone 3 of the f64 kernel (rotated trigonometry) is closed to rays on which the reference scheme itself goes unstable
            // out there: near its polar turning
            // point the theta-oscillator has stiffness k = 3 (Q + a^2 + L^2)^2 / (L^2 Sigma^2), and the two-iteration
            // midpoint blows up for h^2 k >~ 4 -- those rays end as garbage in the oracle too, and only the generic
            // arithmetic reproduces the oracle's garbage step for step (P.rot_stab carries h^2 / r_min^4 with a 32x margin).
            const R qal = Q + hc.a2 + hc.pph2;
            rot_ray = (Q + hc.a2) <= R(P.rot_q_max) && hc.pph2 >= R(P.rot_stab) * qal * qal;
            // ... and only from the radius on where a whole step turns theta by at most 2^-8 (trig_rot_small):
            // |h p_theta / Sigma| <= h B / r_min^2 with B^2 = Q + a^2 >= p_theta^2 (Carter) and r_min = r - travel, i.e.
            // r > travel + sqrt(256 h B). One float per ray, rounded up.
            if (!MIXED && sizeof(R) == 8 && METHOD == 2) {
                const double B = sqrt_nr(fmax((double)(Q + hc.a2), 0.0));
                rot_rsmall = __double2float_ru(P.rot_travel + sqrt_nr(P.rot_k * B));
            }
        }

        // The spectral-LUT copy was issued before ray generation; by now it has had a tile's worth of set-up time to
        // land. (The wait must not sit inside the step loop: an mbarrier try_wait spin loop there stops ptxas from
        // keeping the loop's FP64 constants in uniform registers, which costs a cycle on every three-register DFMA.)
        if (!lut_ready) { mbar_wait(&bars[1], 0); lut_ready = true; }
        R col[3] = {R(0), R(0), R(0)};
        R alpha = R(0), max_drift = R(0), h = R(P.h0);
        uint32_t steps = 0, term = 3u, rhs_evals = 0;
        Vec3<R> gp = {R(0), R(0), R(0)}, gv = {R(0), R(0), R(0)};   // METHOD 3 state (for the parity hook)
        uint32_t photon = 0;
        bool hit = false;
        if (METHOD == 3) {
            // ---- GLSL-semantics march: fragment.glsl.ts:89-221 (deterministic subset) ----
            const R Mh = R(P.M), ah = R(P.a), rph = R(fb->rph), max_dist = R(P.escape_r);
            Vec3<R> ro = {R(fb->cam_pos[0]), R(fb->cam_pos[1]), R(fb->cam_pos[2])};
            const R ron = N::sqrt_(dot3(ro, ro));
            if (ron < rh * R(1.5)) { const R k = rh * R(1.5) / ron; ro.x *= k; ro.y *= k; ro.z *= k; }   // :94-97
            gp = ro; gv = wdir;
            {
                const R cx = ro.y * wdir.z - ro.z * wdir.y, cy = ro.z * wdir.x - ro.x * wdir.z, cz = ro.x * wdir.y - ro.y * wdir.x;
                hit = N::sqrt_(cx * cx + cy * cy + cz * cz) < rh * R(0.9);                                 // :124-126
            }
            R prevY = gp.y;
            const uint32_t nmax = min(P.max_steps, 500u);                                                   // :115
            bool done = false;
            for (uint32_t it0 = 0; it0 < nmax; it0 += 8u) {
                if (__all_sync(0xffffffffu, done)) break;
                const uint32_t it1 = min(it0 + 8u, nmax);
                for (uint32_t it = it0; it < it1; it++) {
                    const R r = N::sqrt_(dot3(gp, gp));
                    if (!done) {
                        if (r < rh * R(1.15)) { hit = true; term = 1u; done = true; }                       // :135-138
                        else if (r > max_dist) { term = 2u; done = true; }
                    }
                    if (!done) {
                        const Vec3<R> pp = gp;
                        const R cdt = glsl_step_size<R>(r, N::abs_(gp.y), rh, rph);
                        Vec3<R> acc; R omega;
                        glsl_accel<R>(gp, gv, Mh, ah, acc, omega);
                        {   // v.xz *= rot(omega dt), rot(t) = mat2(c, -s, s, c)  (chunks/common.ts:46-49)
                            R sn, cs;
                            N::sincos_(omega * cdt, &sn, &cs);
                            const R nx = gv.x * cs - gv.z * sn, nz = gv.x * sn + gv.z * cs;
                            gv.x = nx; gv.z = nz;
                        }
                        const R hdt2 = R(0.5) * cdt * cdt;
                        gp.x = N::fma_(acc.x, hdt2, N::fma_(gv.x, cdt, gp.x));
                        gp.y = N::fma_(acc.y, hdt2, N::fma_(gv.y, cdt, gp.y));
                        gp.z = N::fma_(acc.z, hdt2, N::fma_(gv.z, cdt, gp.z));
                        const R r_new = N::sqrt_(dot3(gp, gp));
                        if (alpha < R(0.95)) {
                            Vec3<R> acc2; R om2;
                            glsl_accel<R>(gp, gv, Mh, ah, acc2, om2);
                            const R hdt = R(0.5) * cdt;
                            gv.x = N::fma_(acc.x + acc2.x, hdt, gv.x);
                            gv.y = N::fma_(acc.y + acc2.y, hdt, gv.y);
                            gv.z = N::fma_(acc.z + acc2.z, hdt, gv.z);
                        }
                        const R ivn = N::rcp_ieee(N::sqrt_(dot3(gv, gv)));
                        gv.x *= ivn; gv.y *= ivn; gv.z *= ivn;
                        if (prevY * gp.y < R(0) && r_new < rph * R(2) && r_new > rh) photon = min(photon + 1u, 3u);
                        prevY = gp.y;
                        steps++; rhs_evals += 2;
                        if (pp.y * gp.y < R(0)) {                                                            // chunks/disk.ts:22-30
                            const R t = N::abs_(pp.y) / N::max_(R(0.0001), N::abs_(pp.y) + N::abs_(gp.y));
                            const Vec3<R> sp = {N::fma_(gp.x - pp.x, t, pp.x), N::fma_(gp.y - pp.y, t, pp.y), N::fma_(gp.z - pp.z, t, pp.z)};
                            const R r_c = N::sqrt_(dot3(sp, sp));
                            if (r_c > R(P.r_in) && r_c < R(P.r_out)) {
                                const R lambda = gp.z * gv.x - gp.x * gv.z;                                  // L_photon, chunks/disk.ts:92
                                const R g = g_factor<R>(r_c, R(P.M), R(P.sqrtM), R(P.spin), lambda);
                                const R tn = sample_tdisk<R>(fb->tdisk, P.tdisk_n, R(P.tdisk_rin), R(P.tdisk_scale), r_c);
                                R rgb[3];
                                sample_spectrum<R>(lut, P.spec_w, P.spec_h, pow04(tn), (g - R(0.05)) * R(1.0 / 4.95), rgb);
                                const R opacity = R(0.6) * tn * g;
                                const R wgt = (R(1) - alpha) * opacity;
                                col[0] = N::fma_(rgb[0], wgt, col[0]);
                                col[1] = N::fma_(rgb[1], wgt, col[1]);
                                col[2] = N::fma_(rgb[2], wgt, col[2]);
                                alpha += opacity;
                            }
                        }
                        if (alpha > R(0.99)) { term = 4u; done = true; }
                    }
                }
            }
        } else {
        // ---- march: geodesic/mod.rs:180-253 ----
        y.pr = renormalize_pr<R, 1>(hc, y.r, y.th, y.pr, y.pth);  // mod.rs:200
        uint32_t renorm_in = 0;          // steps until the next renormalisation (steps % interval == 0, mod.rs:229)
        bool done = false, by_radius = false;
        int side_prev = equator_side(y.th);   // side of the equatorial plane before the step
        // The warp-uniform early exit (all 32 rays finished) is voted once per chunk of 8 steps, outside the inner
        // loop: a vote.sync inside the step loop, like any warp-synchronising instruction there, stops ptxas from
        // keeping the loop's FP64 constants in uniform registers (3-register DFMAs cost 3 cycles instead of 2).
        // Fixed-step methods unroll the inner loop x4 to amortise its bookkeeping; the adaptive stepper's body is far
        // too large for that (it spills when unrolled).
        constexpr uint32_t CHUNK = 8;
        // One march step, in the flavour `kind` (StepKind).
        bool parked = false;               // HCONST chunks: this ray left the chunk's radius window and waits for the replay below
        uint32_t park_it = 0;
        double rot_s = 0.0, rot_c = 0.0;   // zone 3 of the f64 kernel: (sin, cos)(theta), anchored per chunk and chained by rotation
        uint32_t rot_pth_hi = 0u;          //   and the chunk's bound on the high word of |p_theta| (beyond it the ray is parked)
        auto march_step = [&](auto kind, uint32_t it) {
            using K = decltype(kind);
            const R th0 = y.th, r_prev = y.r;
            R hs = R(0);
            if (METHOD == 0) {
                // mod.rs:204,255-265: alive iff 1.001 r+ <= r <= escape radius (which of the two ended it is resolved
                // after the loop from the frozen state)
                if (!done && (lt_pos(y.r, r_term) || gt_pos(y.r, escape_r))) { done = true; by_radius = true; }
                adaptive_step_warp<R, 1>(hc, y, h, R(P.tol), !done, rhs_evals);
            } else {
                // Hot path of the radius tests (f64): ONE unsigned range test on the high word of r says "strictly inside
                // (lo, escape_r)", lo = 1.001 r+ or, in HCONST chunks, the radius where the step rule saturates.
                bool inside;
                if (sizeof(R) == 8) inside = ((uint32_t)hiword(y.r) - P.alive_lo[K::hconst ? 1 : 0]) < P.alive_span[K::hconst ? 1 : 0];
                else inside = K::hconst ? (y.r > R(P.r_sat) && y.r < escape_r) : false;
                if (K::hconst) {
                    // A live ray outside the window broke the chunk's travel bound -- only a numerically blown-up one can.
                    // It is PARKED (frozen like a finished ray, three predicated instructions) and the generic loop replays
                    // it from this very step once the chunk is over, so termination and step size stay exact for every ray
                    // in every zone. (ptxas would predicate the exact tests and the generic step rule into this loop
                    // otherwise -- ~17 issue slots per step for something that happens to ~10 rays per 4K frame.)
                    // (the rotated-trigonometry zone also parks a ray whose p_theta has left the range its |d theta| <= 1/16
                    // bound was derived from: a ray that blew up elsewhere and wandered in)
                    if (K::rot && sizeof(R) == 8) inside = inside && (((uint32_t)hiword(y.pth) & 0x7fffffffu) < rot_pth_hi);
                    if (!inside && !done) { done = true; parked = true; park_it = it; }
                    hs = R(P.h_const);
                } else {
                    if (!inside && !done && (lt_pos(y.r, r_term) || gt_pos(y.r, escape_r))) { done = true; by_radius = true; }
                    hs = WGSL_RULE ? clampPos<R>(N::fma_(y.r, R(0.15), hs_bias), R(0.05), R(1.0)) : R(P.h0);
                }
                // Fixed-step methods run the step for every lane, finished rays with h = 0 on their frozen state (nothing to
                // select back afterwards): in budget accounting that IS the definition; with natural termination the lanes
                // would idle under predication anyway, and keeping the step out of a divergent region lets ptxas feed its
                // constants from uniform registers. The adaptive stepper does the same per attempt (adaptive_step_warp).
                if (done) hs = R(0);
                if (METHOD == 1) { step_rk4<R, 1, DEBUG>(hc, y, hs); rhs_evals += (BUDGET || !done) ? 4u : 0u; }
                else {
                    if constexpr (MIXED && K::far) step_symplectic_mixed<DEBUG>(hc, hcf, y, hs, done ? 0.0f : P.f32_hconst);
                    else if constexpr (K::rot && sizeof(R) == 8) step_symplectic_rot<DEBUG>(hc, y, hs, rot_s, rot_c);
                    else step_symplectic<R, 1, DEBUG, K::polar>(hc, y, hs);
                    rhs_evals += (BUDGET || !done) ? 3u : 0u;
                }
            }
            if (!done) {
                if (renorm_in == 0u) {
                    // (the rotated zone holds sin(theta) of the new state already: no trigonometric evaluation here either)
                    if constexpr (K::rot && sizeof(R) == 8) y.pr = renormalize_pr_far<R>(hc, y.r, R(rot_s), y.pr, y.pth);
                    else y.pr = renormalize_pr<R, 1>(hc, y.r, y.th, y.pr, y.pth);
                    renorm_in = P.renorm_interval;
                }
                renorm_in--;
                steps++;
                if (DEBUG) max_drift = N::max_(max_drift, N::abs_(hamiltonian_of<R, 1>(hc, y.r, y.th, y.pr, y.pth)));
                // ---- thin-disk crossing (compute.wgsl.ts:216-254 with a18 + a19 + a20):
                //      (th0 - pi/2)(th1 - pi/2) <= 0  <=>  the sides differ or the ray sits exactly on the plane
                // (a chunk that provably stays beyond the disk's outer edge compiles the whole block out: K::nodisk)
                const int side = K::nodisk ? 0 : equator_side(y.th);
                const bool crossed = !K::nodisk && ((side != side_prev) || (side == 0));
                if (!K::nodisk) side_prev = side;
                if (crossed) {
                    const R dth = y.th - th0;
                    const R f = (dth == R(0)) ? R(0) : (half_pi - th0) * N::rcp(dth);
                    const R r_c = N::fma_(f, y.r - r_prev, r_prev);
                    if (r_c > R(P.r_in) && r_c < R(P.r_out)) {
                        const R lambda = hc.pph * N::rcp(-hc.pt);
                        const R g = g_factor<R>(r_c, R(P.M), R(P.sqrtM), R(P.spin), lambda);
                        const R tn = sample_tdisk<R>(fb->tdisk, P.tdisk_n, R(P.tdisk_rin), R(P.tdisk_scale), r_c);
                        const R u = pow04(tn);
                        const R v = (g - R(0.05)) * R(1.0 / 4.95);
                        R rgb[3];
                        sample_spectrum<R>(lut, P.spec_w, P.spec_h, u, v, rgb);
                        const R opacity = R(0.6) * tn * g;
                        const R wgt = (R(1) - alpha) * opacity;
                        col[0] = N::fma_(rgb[0], wgt, col[0]);
                        col[1] = N::fma_(rgb[1], wgt, col[1]);
                        col[2] = N::fma_(rgb[2], wgt, col[2]);
                        alpha += opacity;
                        if (alpha > R(0.99)) { term = 4u; done = true; }
                    }
                }
            }
        };
        // The implicit-midpoint march is specialised per 8-step chunk (one warp vote per chunk, like the early exit, so
        // the step loops themselves stay free of warp-synchronising instructions). Every instruction of the loop body is
        // an issue slot -- FP64 instructions take two -- and the kernel is bound by exactly that (profiles/),
        // so what a zone does not need is compiled out of its loop rather than predicated off:
        //   zone 0  generic: step rule + exact radius tests; with the polar clamp iff the tile holds a polar-risk ray
        //   zone 1  r_hconst < r < r_escape_guard for every live ray: h = h_const, no polar clamp
        //   zone 2  (GVT_PRECISION_MIXED) additionally r > r_far and p_r > 0: f32 predictors
        //   zone 3  additionally beyond the disk's outer edge (r > r_rot): no crossing test; f64 rot tiles: rotated trigonometry
        const bool polar_tile = METHOD != 2 || __any_sync(0xffffffffu, polar_ray);
        const bool rot_tile = !MIXED && sizeof(R) == 8 && __all_sync(0xffffffffu, rot_ray);
        for (uint32_t it0 = 0; it0 < P.max_steps; it0 += CHUNK) {
            if (!BUDGET) { if (__all_sync(0xffffffffu, done)) break; }
            const uint32_t it1 = min(it0 + CHUNK, P.max_steps);
            uint32_t zone = 0u;
            if (METHOD == 2 && !polar_tile) {
                const bool z1 = y.r > R(P.r_hconst) && y.r < R(P.r_escape_guard);
                // zone 2: f32 predictors for rays on their way out (MIXED only). zone 3: beyond the disk's outer edge by a
                // chunk's travel -- no equatorial-crossing test; f64: with rotated trigonometry (rot tiles only), MIXED: with
                // the f32 predictors (outbound rays), f32: the saturated-step loop without the disk block
                const bool out = !MIXED || y.pr > R(0);
                const bool far3 = y.r > R(P.r_rot) && out && (MIXED || sizeof(R) == 4 || __double2float_rd((double)y.r) > rot_rsmall);
                const uint32_t lane_zone = done ? 3u : (!z1 ? 0u : (far3 ? 3u : ((MIXED && y.r > R(P.r_far) && out) ? 2u : 1u)));
                zone = __reduce_min_sync(0xffffffffu, lane_zone);
                if (!MIXED && sizeof(R) == 8 && !rot_tile) zone = min(zone, 1u);
            }
            if (MIXED && zone == 3u) {
#pragma unroll(kUnrollFar)
                for (uint32_t it = it0; it < it1; it++) march_step(StepKind<true, true, false, false, true>{}, it);
                side_prev = equator_side(y.th);        // the invariant the crossing test of the other zones relies on
            } else if (MIXED && zone == 2u) {
#pragma unroll(kUnrollFar)
                for (uint32_t it = it0; it < it1; it++) march_step(StepKind<true, true, false>{}, it);
            } else if (METHOD == 2 && !MIXED && sizeof(R) == 8 && zone == 3u) {
                if constexpr (sizeof(R) == 8) {
                    trig_full(P.trig, (double)y.th, rot_s, rot_c);
                    // the |p_theta| the ray's radius gate was derived from, B = (r_small - travel)^2 / (256 h): a ray beyond it
                    // blew up elsewhere and wandered in -- it is parked and replayed by the generic loop
                    const float t = rot_rsmall - P.f32_rot_travel;
                    rot_pth_hi = (uint32_t)__double2hiint((double)(t * t * P.f32_rot_inv_k * 1.001f)) + 1u;
                }
#pragma unroll(kUnrollSymp)
                for (uint32_t it = it0; it < it1; it++) march_step(StepKind<false, true, false, true, true>{}, it);
                side_prev = equator_side(y.th);
            } else if (METHOD == 2 && sizeof(R) == 4 && zone == 3u) {
#pragma unroll(kUnrollSymp)
                for (uint32_t it = it0; it < it1; it++) march_step(StepKind<false, true, false, false, true>{}, it);
                side_prev = equator_side(y.th);
            } else if (METHOD == 2 && zone >= 1u) {
#pragma unroll(MIXED ? kUnrollNearMixed : kUnrollSymp)
                for (uint32_t it = it0; it < it1; it++) march_step(StepKind<false, true, false>{}, it);
            } else if (METHOD == 2 && !polar_tile) {
#pragma unroll(MIXED ? kUnrollNearMixed : kUnrollSymp)
                for (uint32_t it = it0; it < it1; it++) march_step(StepKind<false, false, false>{}, it);
            } else {
#pragma unroll(METHOD == 1 ? kUnrollNear : 1)
                for (uint32_t it = it0; it < it1; it++) march_step(StepKind<false, false, true>{}, it);
            }
            // Replay of parked rays (see HCONST above): the generic step, everyone else frozen, each parked ray from the
            // step it was parked at. Rare enough that its cost is the one vote per chunk.
            if (METHOD == 2 && zone >= 1u && __any_sync(0xffffffffu, parked)) {
                const bool mine = parked, done_saved = done;
                const uint32_t rhs_saved = rhs_evals;
                bool fin = false;                                  // a replayed ray's own "done"
                const uint32_t first = __reduce_min_sync(0xffffffffu, mine ? park_it : it1);
#pragma unroll 1
                for (uint32_t it = first; it < it1; it++) {
                    const bool act = mine && it >= park_it && !fin;
                    done = !act;
                    march_step(StepKind<false, false, true>{}, it);
                    if (act) fin = done;
                }
                done = mine ? fin : done_saved;
                if (BUDGET) rhs_evals = rhs_saved;                 // budget accounting counted these steps in the zone loop already
                parked = false;
            }
        }
        if (by_radius) term = (y.r < r_term) ? 1u : 2u;   // Horizon is tested first (mod.rs:257-263)
        else if (!done) term = 3u;
        }   // METHOD != 3

        // ---- epilogue: coalesced float4 store + census ----
        __syncwarp();
        uint32_t li, lj;
        const bool valid = locate(*reinterpret_cast<volatile uint32_t*>(&s_tile[wid]), li, lj, px, py);
        if (valid) {
            // NaN guard (failure detection): a ray that went non-finite must not poison the TAA history
            float4 px_out = make_float4((float)col[0], (float)col[1], (float)col[2], 1.0f);
            if (!(px_out.x == px_out.x) || !(px_out.y == px_out.y) || !(px_out.z == px_out.z)) px_out = make_float4(0.f, 0.f, 0.f, 1.0f);
            const size_t o = (size_t)py * P.width + px;
            const bool f16 = P.frame_f16 != 0u;
            if (P.frame) store_px(P.frame, o, px_out, f16);
            if (P.host_frame) store_px(P.host_frame, o, px_out, f16);             // posted PCIe write, off the critical path
            for (uint32_t q = 0; q < P.n_peer; q++)                               // fused gather: NVLink peer stores
                store_px(P.peer_frame[q], o, px_out, f16);
            if (DEBUG) {
                const size_t k = (size_t)lj * P.nx + li;
                if (P.dbg_xp && METHOD == 3) {
                    double* o = P.dbg_xp + 8 * k;
                    o[0] = (double)gp.x; o[1] = (double)gp.y; o[2] = (double)gp.z; o[3] = (double)gv.x; o[4] = (double)gv.y;
                    o[5] = (double)gv.z; o[6] = (double)photon; o[7] = hit ? 1.0 : 0.0;
                } else if (P.dbg_xp) {
                    double* o = P.dbg_xp + 8 * k;
                    o[0] = (double)y.t; o[1] = (double)y.r; o[2] = (double)y.th; o[3] = (double)y.ph;
                    o[4] = (double)hc.pt; o[5] = (double)y.pr; o[6] = (double)y.pth; o[7] = (double)hc.pph;
                }
                if (P.dbg_term) P.dbg_term[k] = term;
                if (P.dbg_steps) P.dbg_steps[k] = steps;
                if (P.dbg_drift) P.dbg_drift[k] = (double)max_drift;
                if (P.dbg_rgba) {
                    double* o = P.dbg_rgba + 4 * k;
                    o[0] = (double)col[0]; o[1] = (double)col[1]; o[2] = (double)col[2]; o[3] = 1.0;
                }
            }
        }
        const uint32_t vsteps = valid ? steps : 0u, vrhs = valid ? rhs_evals : 0u;
        {
            const uint32_t t_commit = __reduce_add_sync(0xffffffffu, vsteps), t_rhs = __reduce_add_sync(0xffffffffu, vrhs);
            const uint32_t n_valid = __popc(__ballot_sync(0xffffffffu, valid));
            uint32_t t_term[4];
#pragma unroll
            for (uint32_t c = 0; c < 4; c++) t_term[c] = __popc(__ballot_sync(0xffffffffu, valid && term == c + 1u));
            if (lane == 0) {
                s_acc[wid][0] += t_commit;
                s_acc[wid][1] += BUDGET ? (unsigned long long)n_valid * P.max_steps : (unsigned long long)t_commit;
                s_acc[wid][2] += t_rhs;
#pragma unroll
                for (uint32_t c = 0; c < 4; c++) s_term[wid][c] += t_term[c];
            }
        }
    }
    if (P.timeline && lane == 0) {
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end));
        unsigned long long* o = P.timeline + 3ull * (blockIdx.x * (blockDim.x >> 5) + wid);
        o[0] = s_tstart[wid]; o[1] = t_end; o[2] = s_ntiles[wid];
    }
    // a CTA must not exit while its bulk copy is still in flight
    if (P.lut_in_smem && !lut_ready) mbar_wait(&bars[1], 0);
    // census: one set of global atomics per CTA (2368 warps hitting the same seven addresses at the end of a launch
    // serialise in L2 for tens of microseconds, which shows at 1/8-frame launches)
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long a0 = 0, a1 = 0, a2 = 0;
        uint32_t t0 = 0, t1 = 0, t2 = 0, t3 = 0;
        for (uint32_t w = 0; w < blockDim.x / 32u; w++) {
            a0 += s_acc[w][0]; a1 += s_acc[w][1]; a2 += s_acc[w][2];
            t0 += s_term[w][0]; t1 += s_term[w][1]; t2 += s_term[w][2]; t3 += s_term[w][3];
        }
        atomicAdd(&P.counters->steps_committed, a0);
        atomicAdd(&P.counters->steps_executed, a1);
        atomicAdd(&P.counters->rhs_evals, a2);
        if (t0) atomicAdd(&P.counters->n_horizon, (unsigned long long)t0);
        if (t1) atomicAdd(&P.counters->n_escape, (unsigned long long)t1);
        if (t2) atomicAdd(&P.counters->n_maxsteps, (unsigned long long)t2);
        if (t3) atomicAdd(&P.counters->n_disk, (unsigned long long)t3);
    }
}

template <class R, int METHOD, bool BUDGET, bool DEBUG, int MAXT, bool MIXED = false>
static cudaError_t launch_trace_t(const FrameParams& p, int sm_count, cudaStream_t stream) {
    // the per-step rule of compute.wgsl.ts:213 is a compile-time switch for the fixed-step methods
    auto kern = (METHOD != 0 && p.step_rule == 1u) ? k_trace_tile<R, METHOD, BUDGET, DEBUG, MAXT, true, MIXED>
                                                    : k_trace_tile<R, METHOD, BUDGET, DEBUG, MAXT, false, MIXED>;
    size_t smem = ((sizeof(FrameBlock) + 127) / 128) * 128;
    if (p.lut_in_smem) smem += (size_t)p.spec_w * p.spec_h * sizeof(float4);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint32_t tw = 1u << p.tile_w_log2, th = 32u >> p.tile_w_log2;
    const uint32_t tiles = ((p.nx + tw - 1) / tw) * ((p.ny + th - 1) / th);
    const uint32_t warps_per_cta = MAXT / 32;
    uint32_t ctas = (tiles + warps_per_cta - 1) / warps_per_cta;
    if (ctas > (uint32_t)sm_count) ctas = (uint32_t)sm_count;  // persistent: one CTA per SM
    if (ctas == 0) ctas = 1;
    kern<<<ctas, MAXT, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_trace(const FrameParams& p_in, int method, int precision, bool budget, bool debug, int sm_count,
                         cudaStream_t stream) {
    FrameParams p = p_in;
    // the squarest warp tile that wastes no lanes on the launch's last tile row
    p.tile_w_log2 = (p.ny % 4u != 0u && p.ny % 2u == 0u && p.nx % 16u == 0u) ? 4u : 3u;
    const TrigTable tt = GVT_TRIG_TABLE_INIT;
    p.trig = tt;
#define GVT_DISPATCH(RT, MT)                                                                                   \
    do {                                                                                                         \
        if (method == 0) return debug ? launch_trace_t<RT, 0, false, true, GVT_MAXT_RKF>(p, sm_count, stream)             \
                                      : launch_trace_t<RT, 0, false, false, GVT_MAXT_RKF>(p, sm_count, stream);           \
        if (method == 3) return debug ? launch_trace_t<RT, 3, false, true, MT>(p, sm_count, stream)              \
                                      : launch_trace_t<RT, 3, false, false, MT>(p, sm_count, stream);            \
        if (method == 1 && budget) return debug ? launch_trace_t<RT, 1, true, true, MT>(p, sm_count, stream)    \
                                                : launch_trace_t<RT, 1, true, false, MT>(p, sm_count, stream);  \
        if (method == 1) return debug ? launch_trace_t<RT, 1, false, true, MT>(p, sm_count, stream)              \
                                      : launch_trace_t<RT, 1, false, false, MT>(p, sm_count, stream);            \
        if (budget) return debug ? launch_trace_t<RT, 2, true, true, MT>(p, sm_count, stream)                    \
                                 : launch_trace_t<RT, 2, true, false, MT>(p, sm_count, stream);                  \
        return debug ? launch_trace_t<RT, 2, false, true, MT>(p, sm_count, stream)                               \
                     : launch_trace_t<RT, 2, false, false, MT>(p, sm_count, stream);                             \
    } while (0)
    if (precision == 1) GVT_DISPATCH(float, GVT_MAXT_F32);
    if (precision == 3) {   // GVT_PRECISION_MIXED (implicit midpoint only; validated by the caller)
        if (budget) return debug ? launch_trace_t<double, 2, true, true, GVT_MAXT_F64, true>(p, sm_count, stream)
                                 : launch_trace_t<double, 2, true, false, GVT_MAXT_F64, true>(p, sm_count, stream);
        return debug ? launch_trace_t<double, 2, false, true, GVT_MAXT_F64, true>(p, sm_count, stream)
                     : launch_trace_t<double, 2, false, false, GVT_MAXT_F64, true>(p, sm_count, stream);
    }
    GVT_DISPATCH(double, GVT_MAXT_F64);
#undef GVT_DISPATCH
}

// --------------------------------------------------------------------------------------------------
// geodesic::integrate over explicit initial states (gravitas-wasm/src/lib.rs:422-464), f64.
// --------------------------------------------------------------------------------------------------
template <int COORDS, int METHOD>
__global__ void __launch_bounds__(128) k_integrate_rays(const __grid_constant__ RayBatchParams P) {
    using R = double;
    using N = Num<R>;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const double* in = P.in_xp + 8 * i;
    HoleRay<R> hc;
    hc.trig = &P.trig;
    hc.set_hole(P.M, P.a);
    hc.set_ray(in[4], in[7]);
    Ray<R> y;
    y.t = in[0]; y.r = in[1]; y.th = in[2]; y.ph = in[3]; y.pr = in[5]; y.pth = in[6];
    y.pr = renormalize_pr<R, COORDS>(hc, y.r, y.th, y.pr, y.pth);
    R h = P.h0, max_drift = 0.0;
    uint32_t steps = 0, term = 3u, evals = 0, renorm_in = 0;
    for (uint32_t it = 0; it < P.max_steps; it++) {
        if (y.r < P.r_term) { term = 1u; break; }
        if (y.r > P.escape_r) { term = 2u; break; }
        if (METHOD == 0) {
            h = adaptive_step<R, COORDS>(hc, y, h, P.tol, evals);
        } else {
            const R hs = (P.step_rule == 1u) ? clampR<R>((y.r - P.rh) * 0.15, 0.05, 1.0) : P.h0;
            if (METHOD == 1) { step_rk4<R, COORDS, true>(hc, y, hs); evals += 4; }
            else { step_symplectic<R, COORDS, true>(hc, y, hs); evals += 3; }
        }
        if (renorm_in == 0u) { y.pr = renormalize_pr<R, COORDS>(hc, y.r, y.th, y.pr, y.pth); renorm_in = P.renorm_interval; }
        renorm_in--;
        max_drift = N::max_(max_drift, N::abs_(hamiltonian_of<R, COORDS>(hc, y.r, y.th, y.pr, y.pth)));
        steps++;
    }
    double* o = P.out_xp + 8 * i;
    o[0] = y.t; o[1] = y.r; o[2] = y.th; o[3] = y.ph; o[4] = hc.pt; o[5] = y.pr; o[6] = y.pth; o[7] = hc.pph;
    if (P.term) P.term[i] = term;
    if (P.steps) P.steps[i] = steps;
    if (P.drift) P.drift[i] = max_drift;
    if (P.rhs) P.rhs[i] = evals;
}

cudaError_t launch_integrate_rays(const RayBatchParams& p_in, cudaStream_t stream) {
    if (p_in.n == 0) return cudaSuccess;
    RayBatchParams p = p_in;
    const TrigTable tt = GVT_TRIG_TABLE_INIT;
    p.trig = tt;
    const unsigned blocks = (unsigned)((p.n + 127) / 128);
#define GVT_RB(C, M) k_integrate_rays<C, M><<<blocks, 128, 0, stream>>>(p)
    if (p.coords == 1) { if (p.method == 0) GVT_RB(1, 0); else if (p.method == 1) GVT_RB(1, 1); else GVT_RB(1, 2); }
    else { if (p.method == 0) GVT_RB(0, 0); else if (p.method == 1) GVT_RB(0, 1); else GVT_RB(0, 2); }
#undef GVT_RB
    return cudaGetLastError();
}

// --------------------------------------------------------------------------------------------------
// TAA resolve (ataa.wgsl.ts:28-83; mode 1 = reprojection.glsl.ts:70-115). One warp owns a 30-pixel-wide, R-row
// strip unit and walks down it with a rolling 3-row window: every lane loads ONE pixel per row (lanes 0 and 31 are
// the horizontal halo), the vertical 3-row sums live in registers and the horizontal 3-tap sums come from warp
// shuffles. HBM-bound by design (16 B cur + 16 B history + 16 B store per pixel), so the per-pixel instruction
// count is what had to come down to reach that bound:
//  * the reprojection chain clip -> inv_proj -> normalise -> inv_view -> +12 d -> prev_view_proj is affine in
//    (cx, cy) up to ONE scalar s = sign(w) / |v|: pc = c0 + s (gA cx + gB cy + gC), with the 4x3 products formed
//    once per frame on the host in f64 (TaaParams). 3+3+3 FMAs, one MUFU.RSQ and one MUFU.RCP per pixel instead
//    of three mat4 x vec4 products and five IEEE divisions;
//  * everything that depends on the column only (cx terms, clamped column index) is hoisted out of the row loop,
//    everything that depends on the row only is warp-uniform;
//  * sigma uses MUFU.SQRT (sqrt.approx): the pass is f32 with a 1e-3 conditioning floor (tests/test_gpu_taa.py).
//  * every global load runs ahead of its use at no register cost: current-frame rows AND the four history taps go
//    through per-lane cp.async rings in shared memory, two row-iterations deep (45 KB per CTA).
// Work units are distributed grid-stride over a grid sized to the resident warps, and the rows per unit are chosen
// per launch so that every warp walks the same number of units (no partial last wave).
// Measured (4K, B200): 163 us (first version, 440 instr/pixel) -> 113 us (factored reprojection, register prefetch)
// -> 101 us with the tap ring (64 registers, 32 warps/SM); the WebGL2 variant (one tap) 87 us. The same strip walk as a
// bare 2-read/1-write copy runs 66 us (scripts/ubench/strip_copy.cu): the rest is the resolve's own arithmetic.
// --------------------------------------------------------------------------------------------------
constexpr int TAA_STRIP_W = 30;
#ifndef GVT_TAA_MINB
#define GVT_TAA_MINB 4
#endif
#ifndef GVT_TAA_UNROLL
#define GVT_TAA_UNROLL 1
#endif
constexpr int taa_unroll = GVT_TAA_UNROLL;
#ifndef GVT_TAA_DEPTH
#define GVT_TAA_DEPTH 2
#endif
constexpr int TAA_DEPTH = GVT_TAA_DEPTH;

struct YCC { float y, co, cg; };
__device__ __forceinline__ YCC rgb_to_ycocg(float r, float g, float b) {  // ataa.wgsl.ts:11-16
    YCC o;                                                                 // (the 1/4, 1/2 weights are exact scalings)
    const float t = r + b, hg = 0.5f * g;
    o.y = fmaf(0.25f, t, hg);
    o.co = 0.5f * (r - b);
    o.cg = fmaf(-0.25f, t, hg);
    return o;
}
struct YCC2 { float y, co, cg, yy, coco, cgcg; };   // a pixel's first and second moments
__device__ __forceinline__ YCC2 moments_of(const float4& p) {
    const YCC c = rgb_to_ycocg(p.x, p.y, p.z);
    return YCC2{c.y, c.co, c.cg, c.y * c.y, c.co * c.co, c.cg * c.cg};
}
__device__ __forceinline__ float sum3(float v) {
    return __shfl_up_sync(0xffffffffu, v, 1) + v + __shfl_down_sync(0xffffffffu, v, 1);
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// per-lane 16-B asynchronous global -> shared copy (LDGSTS): in flight without holding a register or a scoreboard
__device__ __forceinline__ void cp_async16_s(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t a, float x, float y) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 8-B flavour of the per-lane asynchronous copy, for RGBA16F frames (cp.async.cg is 16-B only)
__device__ __forceinline__ void cp_async8_s(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ uint2 lds64u(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
// F16: cur / hist / out are RGBA16F (8 B per pixel); the arithmetic stays f32.
template <int MODE, bool F16>
__global__ void __launch_bounds__(256, GVT_TAA_MINB) k_taa_resolve(const __grid_constant__ TaaParams P) {
    constexpr int NT = (MODE == 1) ? 1 : 4;                        // history taps per pixel
    constexpr uint32_t ES = F16 ? 8u : 16u;                        // bytes per pixel of the frame chain
    constexpr uint32_t LS = 32u * ES;                              // one ring row: 32 lanes
    auto cp_px = [](uint32_t dst, const void* src) { if (F16) cp_async8_s(dst, src); else cp_async16_s(dst, src); };
    auto ld_px = [](uint32_t a) { return F16 ? unpack_half4(lds64u(a)) : lds128(a); };
    // per-warp, per-lane rings: every lane reads back only what it copied itself, so no warp-level synchronisation
    extern __shared__ __align__(16) unsigned char taa_smem[];     // dynamic: TAA_DEPTH = 3 already exceeds the 48 KB static limit
    // this lane's ring entries, as shared-space byte addresses fixed for the whole kernel: slot s of the current-frame
    // ring is at my_cur + s * LS, tap t of slot s at my_tap + (s * NT + t) * LS, the fractions at my_frac + s * 256
    const uint32_t smem0 = smem_u32(taa_smem), w_ = threadIdx.x >> 5, l_ = threadIdx.x & 31;
    const uint32_t my_cur = smem0 + (w_ * TAA_DEPTH * 32u + l_) * ES;
    const uint32_t my_tap = smem0 + 8u * TAA_DEPTH * LS + (w_ * TAA_DEPTH * NT * 32u + l_) * ES;
    const uint32_t my_frac = smem0 + 8u * TAA_DEPTH * LS * (1u + NT) + (w_ * TAA_DEPTH * 32u + l_) * 8u;
    const int W = (int)P.width, H = (int)P.height;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int strips_x = (W + TAA_STRIP_W - 1) / TAA_STRIP_W;
    const int rows = (int)P.row1 - (int)P.row0;
    const bool striped = P.stripe.s != 0u;                        // GVT_FLAG_ROW_INTERLEAVE: one unit row-range per owned stripe
    const int TAA_ROWS = striped ? (int)P.stripe.s : (int)P.unit_rows;
    const int n_work = strips_x * (striped ? (int)P.n_stripes : (rows + TAA_ROWS - 1) / TAA_ROWS);
    const int n_warps = (int)(gridDim.x * (blockDim.x >> 5));
    const float nsig = (MODE == 1) ? 1.5f : 2.0f;                 // reprojection.glsl.ts:90-91 / ataa.wgsl.ts:51-52
    const float k9 = 1.0f / 9.0f;
    const float half_w = 0.5f * (float)W, half_h = 0.5f * (float)H;
    const float x_max = (float)(W - 1), y_max = (float)(H - 1);
    const float fb_gl = P.moving ? 0.0f : P.blend;

    for (int unit = (int)(blockIdx.x * (blockDim.x >> 5) + wid); unit < n_work; unit += n_warps) {
        const int sx = unit % strips_x, sy = unit / strips_x;
        const int x = sx * TAA_STRIP_W + lane - 1;                // lane 0 / 31 = halo columns
        const int xc = max(0, min(x, W - 1));                     // clamp(pos + d, 0, size-1), ataa.wgsl.ts:43
        const bool owner = lane >= 1 && lane <= TAA_STRIP_W && x < W;
        const int y_begin = striped ? (sy * (int)P.stripe.world + (int)P.stripe.rank) * TAA_ROWS : (int)P.row0 + sy * TAA_ROWS;
        const int y_end = min(y_begin + TAA_ROWS, striped ? H : (int)P.row1);
        const int n_rows = y_end - y_begin;
        if (n_rows <= 0) continue;
        const char* col = reinterpret_cast<const char*>(P.cur) + (size_t)xc * ES;
        const char* hist_b = reinterpret_cast<const char*>(P.hist);

        // column-only part of the reprojection (MODE 0)
        const float cx = fmaf((float)x + 0.5f, 2.0f / (float)W, -1.0f);
        const float vx0 = fmaf(P.vA[0], cx, P.vC[0]), vx1 = fmaf(P.vA[1], cx, P.vC[1]), vx2 = fmaf(P.vA[2], cx, P.vC[2]);
        const float vxw = fmaf(P.vA[3], cx, P.vC[3]);
        const float gx0 = fmaf(P.gA[0], cx, P.gC[0]), gx1 = fmaf(P.gA[1], cx, P.gC[1]), gx2 = fmaf(P.gA[2], cx, P.gC[2]);

        // Every global load of the walk is a 16-B cp.async into the rings, TAA_DEPTH rows ahead of its use and at no
        // register cost (a DRAM miss costs more than one row-iteration of this loop, and holding a row of taps in
        // registers ahead of time costs a quarter of the occupancy). Group k carries what row-iteration j = k - 2
        // consumes: current-frame row clamp(y_begin - 1 + k) -- the bottom row of j's 3x3 window -- and the history
        // taps (+ bilinear fractions) of row y_begin + j.
        auto issue = [&](int k) {
            const uint32_t slot = (uint32_t)k % TAA_DEPTH;
            if (k <= n_rows + 1) cp_px(my_cur + slot * LS, col + (size_t)min(max(y_begin - 1 + k, 0), H - 1) * W * ES);
            const int j = k - 2;
            if (j >= 0 && j < n_rows) {
                const int y = y_begin + j;
                const uint32_t tap = my_tap + slot * (NT * LS);
                if (MODE == 1) {
                    // WebGL2 resolve: history at the same texel (reprojection.glsl.ts:93-110)
                    cp_px(tap, hist_b + ((size_t)y * W + xc) * ES);
                } else {
                    // reprojection at depth 12 through prev_view_proj (ataa.wgsl.ts:54-69), factored form (header)
                    const float cy = -fmaf((float)y + 0.5f, 2.0f / (float)H, -1.0f);
                    const float v0 = fmaf(P.vB[0], cy, vx0), v1 = fmaf(P.vB[1], cy, vx1), v2 = fmaf(P.vB[2], cy, vx2);
                    const float vw = fmaf(P.vB[3], cy, vxw);
                    const float s = copysignf(rsqrtf(fmaf(v0, v0, fmaf(v1, v1, v2 * v2))), vw);
                    const float p0 = fmaf(s, fmaf(P.gB[0], cy, gx0), P.c0[0]);
                    const float p1 = fmaf(s, fmaf(P.gB[1], cy, gx1), P.c0[1]);
                    const float p3 = fmaf(s, fmaf(P.gB[2], cy, gx2), P.c0[2]);
                    const float ip3 = rcp_approx(p3);
                    // bilinear, clamp-to-edge history fetch (textureSampleLevel + linear sampler, ataa.wgsl.ts:72):
                    // pu W - 0.5 with pu = ndc.x/2 + 1/2,  pv H - 0.5 with pv = -ndc.y/2 + 1/2
                    const float hx = fminf(fmaxf(fmaf(p0 * ip3, half_w, half_w - 0.5f), 0.0f), x_max);
                    const float hy = fminf(fmaxf(fmaf(p1 * ip3, -half_h, half_h - 0.5f), 0.0f), y_max);
                    const float hxf = floorf(hx), hyf = floorf(hy);
                    const int x0 = (int)hxf, y0 = (int)hyf;
                    // one 64-bit address, the other three taps by small element offsets (0 or 1 column, 0 or W rows)
                    const char* t00 = hist_b + ((size_t)y0 * W + x0) * ES;
                    const size_t dx = (x0 + 1 < W) ? ES : 0, dy = (y0 + 1 < H) ? (size_t)W * ES : 0;
                    cp_px(tap, t00); cp_px(tap + LS, t00 + dx);
                    cp_px(tap + 2u * LS, t00 + dy); cp_px(tap + 3u * LS, t00 + dy + dx);
                    sts64(my_frac + slot * 256u, hx - hxf, hy - hyf);
                }
            }
            cp_async_commit();   // one group per k, empty past the end, so wait_group counts stay aligned
        };
#pragma unroll
        for (int k = 0; k < TAA_DEPTH; k++) issue(k);
        // rolling window of per-pixel YCoCg moments for rows y-1, y, y+1
        YCC2 a, b, c;
        cp_async_wait<TAA_DEPTH - 1>();
        a = moments_of(ld_px(my_cur));
        issue(TAA_DEPTH);
        cp_async_wait<TAA_DEPTH - 1>();
        b = moments_of(ld_px(my_cur + (1u % TAA_DEPTH) * LS));
        issue(TAA_DEPTH + 1);
#pragma unroll taa_unroll
        for (int y = y_begin; y < y_end; y++) {
            const int k = y - y_begin + 2;
            const uint32_t slot = (uint32_t)k % TAA_DEPTH;
            cp_async_wait<TAA_DEPTH - 1>();
            c = moments_of(ld_px(my_cur + slot * LS));
            // vertical sums (this lane's column), then horizontal 3-tap by shuffles
            const float m1y = sum3(a.y + b.y + c.y), m1o = sum3(a.co + b.co + c.co), m1g = sum3(a.cg + b.cg + c.cg);
            const float m2y = sum3(a.yy + b.yy + c.yy), m2o = sum3(a.coco + b.coco + c.coco);
            const float m2g = sum3(a.cgcg + b.cgcg + c.cgcg);
            if (owner) {
                const float mean[3] = {m1y * k9, m1o * k9, m1g * k9};
                const float m2[3] = {m2y * k9, m2o * k9, m2g * k9};
                float lo[3], hi[3], sd_y = 0.0f;
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    const float sd = sqrt_approx(fmaxf(fmaf(-mean[i], mean[i], m2[i]), 0.0f));
                    if (i == 0) sd_y = sd;
                    lo[i] = fmaf(-nsig, sd, mean[i]); hi[i] = fmaf(nsig, sd, mean[i]);
                }
                float hr, hg, hb, fb_;
                const uint32_t tap = my_tap + slot * (NT * LS);
                const float4 h00 = ld_px(tap);
                if (MODE == 1) {
                    hr = h00.x; hg = h00.y; hb = h00.z;
                    fb_ = fb_gl * (1.0f - fminf(fmaxf(sd_y * 4.0f, 0.0f), 0.55f));   // variance-guided weight
                } else {
                    const float4 h10 = ld_px(tap + LS), h01 = ld_px(tap + 2u * LS), h11 = ld_px(tap + 3u * LS);
                    const float2 fr = lds64(my_frac + slot * 256u);
                    const float tr = fmaf(h10.x - h00.x, fr.x, h00.x), br = fmaf(h11.x - h01.x, fr.x, h01.x);
                    const float tg = fmaf(h10.y - h00.y, fr.x, h00.y), bg = fmaf(h11.y - h01.y, fr.x, h01.y);
                    const float tb = fmaf(h10.z - h00.z, fr.x, h00.z), bb = fmaf(h11.z - h01.z, fr.x, h01.z);
                    hr = fmaf(br - tr, fr.y, tr); hg = fmaf(bg - tg, fr.y, tg); hb = fmaf(bb - tb, fr.y, tb);
                    fb_ = 0.92f;  // ataa.wgsl.ts:77
                }
                YCC hs = rgb_to_ycocg(hr, hg, hb);
                hs.y = fminf(fmaxf(hs.y, lo[0]), hi[0]);
                hs.co = fminf(fmaxf(hs.co, lo[1]), hi[1]);
                hs.cg = fminf(fmaxf(hs.cg, lo[2]), hi[2]);
                const float ry = fmaf(hs.y - b.y, fb_, b.y), ro = fmaf(hs.co - b.co, fb_, b.co);
                const float rg = fmaf(hs.cg - b.cg, fb_, b.cg);
                // YCoCgToRGB, ataa.wgsl.ts:18-26
                const float4 px_out = make_float4(ry + ro - rg, ry + rg, ry - ro - rg, 1.0f);
                const size_t o = (size_t)y * W + x;
                store_px(P.out, o, px_out, F16);
                if (P.host_out) store_px(P.host_out, o, px_out, F16);
#pragma unroll 1
                for (uint32_t q = 0; q < P.n_peer; q++) store_px(P.peer_out[q], o, px_out, F16);
            }
            a = b; b = c;
            issue(k + TAA_DEPTH);   // refills the slot this iteration has just finished reading
        }
        cp_async_wait<0>();         // nothing of this unit may land after the next unit starts reusing the slots
    }
}

cudaError_t launch_taa(const TaaParams& p_in, int sm_count, cudaStream_t stream) {
    TaaParams p = p_in;
    const int strips_x = ((int)p.width + TAA_STRIP_W - 1) / TAA_STRIP_W;
    const int rows = (int)p.row1 - (int)p.row0;
    const bool striped = p.stripe.s != 0u;
    if (striped ? p.n_stripes == 0u : rows <= 0) return cudaSuccess;
    const int wpb = 8;
    static PerDeviceInt resident_cache[4];   // CTAs per SM of each instantiation, per device (occupancy query, once)
    const int m = p.mode == 1u ? 1 : 0, f = p.frame_f16 ? 1 : 0;
    // rings per CTA: 8 warps x TAA_DEPTH x 32 lanes x (ES B current + NT x ES B taps + 8 B fractions)
    const size_t es = f ? 8 : 16;
    const size_t smem = (size_t)8 * TAA_DEPTH * 32 * (es + (m ? 1 : 4) * es + 8);
    void (*kern)(TaaParams) = m ? (f ? k_taa_resolve<1, true> : k_taa_resolve<1, false>) : (f ? k_taa_resolve<0, true> : k_taa_resolve<0, false>);
    int* cached = resident_cache[2 * m + f].slot();
    int resident = cached ? *cached : 0;
    if (!resident) {
        int n = 0;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, wpb * 32, smem);
        if (e != cudaSuccess) return e;
        resident = n > 0 ? n : 1;
        if (cached) *cached = resident;
    }
    // Rows per strip unit: every warp walks k = ceil(units / resident warps) units of (R + 2) loaded rows; pick the R
    // that minimises k (R + 2), i.e. no partial last wave and as little halo as the frame allows.
    const int max_warps = max(sm_count, 1) * resident * wpb;
    int best_r = 16;
    long best_cost = -1;
    for (int R = 4; R <= 64; R++) {
        const long units = (long)strips_x * ((rows + R - 1) / R);
        const long cost = ((units + max_warps - 1) / max_warps) * (R + 2);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_r = R; }
    }
    p.unit_rows = (uint32_t)best_r;
    const int n_work = striped ? strips_x * (int)p.n_stripes : strips_x * ((rows + best_r - 1) / best_r);
    const int blocks = min((n_work + wpb - 1) / wpb, max(sm_count, 1) * resident);
    kern<<<blocks, wpb * 32, smem, stream>>>(p);
    return cudaGetLastError();
}

// --------------------------------------------------------------------------------------------------
// The PRECISE build of the resolve: one thread per pixel, every operation an explicitly rounded IEEE f32 operation
// (__fadd_rn / __fmul_rn / __fdiv_rn / __fsqrt_rn: never contracted into FMAs, never approximated) in the order the
// shader text evaluates them -- ataa.wgsl.ts:28-83 with the UNFOLDED reprojection chain of :54-69 (inv_proj, normalise,
// inv_view, + 12 d, prev_view_proj as three mat4 x vec4 products), reprojection.glsl.ts:70-115 for MODE 1. It exists so
// that the resolve's SEMANTICS can be tested to 1e-6 against the numpy restatement (tests/test_gpu_taa.py), and so that
// what the production kernel's MUFU.SQRT / MUFU.RCP / MUFU.RSQ, host-folded matrices and shuffle-order sums cost in
// accuracy is a measured number (fast vs precise) instead of an argument. ~4x slower than k_taa_resolve; never on the
// frame path unless GVT_FLAG_TAA_PRECISE asks for it.
// --------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ycocg_rn(float r, float g, float b, float& y, float& co, float& cg) {   // ataa.wgsl.ts:11-16
    y = __fadd_rn(__fadd_rn(__fmul_rn(0.25f, r), __fmul_rn(0.5f, g)), __fmul_rn(0.25f, b));
    co = __fadd_rn(__fmul_rn(0.5f, r), -__fmul_rn(0.5f, b));
    cg = __fadd_rn(__fadd_rn(__fmul_rn(-0.25f, r), __fmul_rn(0.5f, g)), -__fmul_rn(0.25f, b));
}
// column-major mat4 x vec4, terms summed in column order (((m0 v0 + m1 v1) + m2 v2) + m3 v3)
__device__ __forceinline__ void mat_vec_rn(const float* m, const float v[4], float out[4]) {
#pragma unroll
    for (int r = 0; r < 4; r++)
        out[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[r], v[0]), __fmul_rn(m[4 + r], v[1])), __fmul_rn(m[8 + r], v[2])),
                           __fmul_rn(m[12 + r], v[3]));
}
template <int MODE>
__global__ void __launch_bounds__(256) k_taa_resolve_precise(const __grid_constant__ TaaParams P) {
    const int W = (int)P.width, H = (int)P.height;
    const int x = (int)(blockIdx.x * 32u + (threadIdx.x & 31u));
    const int y = (int)P.row0 + (int)(blockIdx.y * 8u + (threadIdx.x >> 5));
    if (x >= W || y >= (int)P.row1) return;
    const bool f16 = P.frame_f16 != 0u;              // RGBA16F frames: texels widen exactly, the result is rounded once on store
    float m1[3] = {0.f, 0.f, 0.f}, m2[3] = {0.f, 0.f, 0.f}, c0[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {                                          // ataa.wgsl.ts:36-48, clamp :43
            const float4 p = load_px(P.cur, (size_t)min(max(y + dy, 0), H - 1) * W + min(max(x + dx, 0), W - 1), f16);
            float s[3];
            ycocg_rn(p.x, p.y, p.z, s[0], s[1], s[2]);
#pragma unroll
            for (int i = 0; i < 3; i++) { m1[i] = __fadd_rn(m1[i], s[i]); m2[i] = __fadd_rn(m2[i], __fmul_rn(s[i], s[i])); }
            if (dy == 0 && dx == 0) { c0[0] = s[0]; c0[1] = s[1]; c0[2] = s[2]; }
        }
    const float nsig = (MODE == 1) ? 1.5f : 2.0f;
    float lo[3], hi[3], sd0 = 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float mean = __fdiv_rn(m1[i], 9.0f);
        const float sd = __fsqrt_rn(fmaxf(__fadd_rn(__fdiv_rn(m2[i], 9.0f), -__fmul_rn(mean, mean)), 0.0f));
        if (i == 0) sd0 = sd;
        lo[i] = __fadd_rn(mean, -__fmul_rn(nsig, sd)); hi[i] = __fadd_rn(mean, __fmul_rn(nsig, sd));
    }
    float hr, hg, hb, alpha;
    if (MODE == 1) {
        const float4 h = load_px(P.hist, (size_t)y * W + x, f16);                   // same texel (reprojection.glsl.ts:93-110)
        hr = h.x; hg = h.y; hb = h.z;
        const float wv = __fadd_rn(1.0f, -fminf(fmaxf(__fmul_rn(sd0, 4.0f), 0.0f), 0.55f));
        alpha = P.moving ? 0.0f : __fmul_rn(P.blend, wv);
    } else {
        // ataa.wgsl.ts:54-69
        const float u = __fdiv_rn(__fadd_rn((float)x, 0.5f), (float)W), v = __fdiv_rn(__fadd_rn((float)y, 0.5f), (float)H);
        const float clip[4] = {__fadd_rn(__fmul_rn(u, 2.0f), -1.0f), -__fadd_rn(__fmul_rn(v, 2.0f), -1.0f), 1.0f, 1.0f};
        float vt[4];
        mat_vec_rn(P.m_inv_proj, clip, vt);
        float vd[3] = {__fdiv_rn(vt[0], vt[3]), __fdiv_rn(vt[1], vt[3]), __fdiv_rn(vt[2], vt[3])};
        const float n = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(vd[0], vd[0]), __fmul_rn(vd[1], vd[1])), __fmul_rn(vd[2], vd[2])));
        const float d4[4] = {__fdiv_rn(vd[0], n), __fdiv_rn(vd[1], n), __fdiv_rn(vd[2], n), 0.0f};
        float wd[4];
        mat_vec_rn(P.m_inv_view, d4, wd);
        const float wp[4] = {__fadd_rn(P.cam_pos[0], __fmul_rn(wd[0], 12.0f)), __fadd_rn(P.cam_pos[1], __fmul_rn(wd[1], 12.0f)),
                             __fadd_rn(P.cam_pos[2], __fmul_rn(wd[2], 12.0f)), 1.0f};
        float pc[4];
        mat_vec_rn(P.m_prev_vp, wp, pc);
        const float pu = __fadd_rn(__fmul_rn(__fdiv_rn(pc[0], pc[3]), 0.5f), 0.5f);
        const float pv = __fadd_rn(__fmul_rn(__fdiv_rn(pc[1], pc[3]), -0.5f), 0.5f);
        // textureSampleLevel + linear sampler, clamp-to-edge (ataa.wgsl.ts:72)
        const float fx_ = fminf(fmaxf(__fadd_rn(__fmul_rn(pu, (float)W), -0.5f), 0.0f), (float)(W - 1));
        const float fy_ = fminf(fmaxf(__fadd_rn(__fmul_rn(pv, (float)H), -0.5f), 0.0f), (float)(H - 1));
        const int x0 = (int)floorf(fx_), y0 = (int)floorf(fy_);
        const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
        const float fx = __fadd_rn(fx_, -(float)x0), fy = __fadd_rn(fy_, -(float)y0);
        const float4 t00 = load_px(P.hist, (size_t)y0 * W + x0, f16), t10 = load_px(P.hist, (size_t)y0 * W + x1, f16);
        const float4 t01 = load_px(P.hist, (size_t)y1 * W + x0, f16), t11 = load_px(P.hist, (size_t)y1 * W + x1, f16);
        auto bil = [&](float a00, float a10, float a01, float a11) {
            const float top = __fadd_rn(a00, __fmul_rn(__fadd_rn(a10, -a00), fx));
            const float bot = __fadd_rn(a01, __fmul_rn(__fadd_rn(a11, -a01), fx));
            return __fadd_rn(top, __fmul_rn(__fadd_rn(bot, -top), fy));
        };
        hr = bil(t00.x, t10.x, t01.x, t11.x); hg = bil(t00.y, t10.y, t01.y, t11.y); hb = bil(t00.z, t10.z, t01.z, t11.z);
        alpha = 0.92f;                                                                // ataa.wgsl.ts:77
    }
    float hy, hco, hcg;
    ycocg_rn(hr, hg, hb, hy, hco, hcg);
    hy = fminf(fmaxf(hy, lo[0]), hi[0]); hco = fminf(fmaxf(hco, lo[1]), hi[1]); hcg = fminf(fmaxf(hcg, lo[2]), hi[2]);
    // mix(center, history, alpha) = center (1 - alpha) + history alpha
    const float ia = (MODE == 1) ? __fadd_rn(1.0f, -alpha) : 0.08f;                   // float32(1 - 0.92) = 0.08f
    const float ry = __fadd_rn(__fmul_rn(c0[0], ia), __fmul_rn(hy, alpha));
    const float ro = __fadd_rn(__fmul_rn(c0[1], ia), __fmul_rn(hco, alpha));
    const float rg = __fadd_rn(__fmul_rn(c0[2], ia), __fmul_rn(hcg, alpha));
    // YCoCgToRGB, ataa.wgsl.ts:18-26
    const float4 px_out = make_float4(__fadd_rn(__fadd_rn(ry, ro), -rg), __fadd_rn(ry, rg), __fadd_rn(__fadd_rn(ry, -ro), -rg), 1.0f);
    const size_t o = (size_t)y * W + x;
    store_px(P.out, o, px_out, f16);
    if (P.host_out) store_px(P.host_out, o, px_out, f16);
    for (uint32_t q = 0; q < P.n_peer; q++) store_px(P.peer_out[q], o, px_out, f16);
}

cudaError_t launch_taa_precise(const TaaParams& p, cudaStream_t stream) {
    const int rows = (int)p.row1 - (int)p.row0;
    if (rows <= 0 || p.stripe.s != 0u) return rows <= 0 ? cudaSuccess : cudaErrorNotSupported;
    const dim3 grid((p.width + 31u) / 32u, ((unsigned)rows + 7u) / 8u);
    if (p.mode == 1u) k_taa_resolve_precise<1><<<grid, 256, 0, stream>>>(p);
    else k_taa_resolve_precise<0><<<grid, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

// RGBA32F -> RGBA16F (reprojection.ts:120-140 / webgpu/renderer.ts:161-180 texture format)
__global__ void k_f32_to_f16(const float4* __restrict__ src, uint2* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = src[i];
        const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&lo);
        o.y = *reinterpret_cast<const uint32_t*>(&hi);
        dst[i] = o;
    }
}
cudaError_t launch_f32_to_f16(const float4* src, void* dst, size_t n_px, cudaStream_t stream) {
    k_f32_to_f16<<<148 * 8, 256, 0, stream>>>(src, reinterpret_cast<uint2*>(dst), n_px);
    return cudaGetLastError();
}

// RGBA16F frame chain -> RGBA32F hosts
__global__ void k_f16_to_f32(const uint2* __restrict__ src, float4* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = unpack_half4(src[i]);
}
cudaError_t launch_f16_to_f32(const void* src, float4* dst, size_t n_px, cudaStream_t stream) {
    k_f16_to_f32<<<148 * 8, 256, 0, stream>>>(reinterpret_cast<const uint2*>(src), dst, n_px);
    return cudaGetLastError();
}

// Display-ready 8-bit output: Reinhard (webgpu/renderer.ts:45-47) or ACES + gamma (bloom.glsl.ts:106-124, no bloom).
__device__ __forceinline__ float aces_gamma(float c) {
    const float t = fminf(fmaxf((c * (2.51f * c + 0.03f)) / (c * (2.43f * c + 0.59f) + 0.14f), 0.0f), 1.0f);
    return powf(t, 0.4545f);
}
__device__ __forceinline__ uint32_t unorm8(float v) { return (uint32_t)__float2int_rn(fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f); }
__global__ void k_tonemap_rgba8(const float4* __restrict__ src, uint32_t* __restrict__ dst, size_t n, int aces, bool src_f16) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = load_px(src, i, src_f16);
        float r, g, b;
        if (aces == 1) { r = aces_gamma(v.x); g = aces_gamma(v.y); b = aces_gamma(v.z); }
        else if (aces == 2) { r = v.x; g = v.y; b = v.z; }   // already display-referred: quantise only
        else { r = v.x / (v.x + 1.0f); g = v.y / (v.y + 1.0f); b = v.z / (v.z + 1.0f); }
        dst[i] = unorm8(r) | (unorm8(g) << 8) | (unorm8(b) << 16) | (unorm8(v.w) << 24);
    }
}
cudaError_t launch_tonemap_rgba8(const float4* src, void* dst, size_t n_px, int aces, cudaStream_t stream, bool src_f16) {
    k_tonemap_rgba8<<<148 * 8, 256, 0, stream>>>(src, reinterpret_cast<uint32_t*>(dst), n_px, aces, src_f16);
    return cudaGetLastError();
}

// --------------------------------------------------------------------------------------------------
// FMA-pipe peak micro-benchmark: 8 independent dependent chains per thread.
// --------------------------------------------------------------------------------------------------
template <class R>
__global__ void __launch_bounds__(256) k_fma_peak(float* sink, unsigned long long iters) {
    R a0 = R(threadIdx.x) * R(1e-3), a1 = a0 + R(1), a2 = a0 + R(2), a3 = a0 + R(3), a4 = a0 + R(4), a5 = a0 + R(5),
      a6 = a0 + R(6), a7 = a0 + R(7);
    const R x = R(0.999999), yv = R(1e-6);
    for (unsigned long long i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = Num<R>::fma_(a0, x, yv); a1 = Num<R>::fma_(a1, x, yv); a2 = Num<R>::fma_(a2, x, yv);
            a3 = Num<R>::fma_(a3, x, yv); a4 = Num<R>::fma_(a4, x, yv); a5 = Num<R>::fma_(a5, x, yv);
            a6 = Num<R>::fma_(a6, x, yv); a7 = Num<R>::fma_(a7, x, yv);
        }
    }
    const R s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == R(-12345.678)) sink[0] = (float)s;  // never true; keeps the chains live
}

cudaError_t launch_fma_peak(int precision, int sm_count, unsigned long long iters, float* sink, cudaStream_t stream,
                            double* flops_out) {
    const int blocks = sm_count * 8, threads = 256;
    if (precision == 1) k_fma_peak<float><<<blocks, threads, 0, stream>>>(sink, iters);
    else k_fma_peak<double><<<blocks, threads, 0, stream>>>(sink, iters);
    *flops_out = (double)blocks * threads * (double)iters * 64.0 * 2.0;
    return cudaGetLastError();
}

}  // namespace gvt
