// gvt_internal.h — host<->kernel contract inside libgravitas_b200.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gvt_device.cuh"
#include "../../include/gravitas_b200.h"

namespace gvt {

// Staged by TMA (cp.async.bulk) from global into shared memory once per CTA: the camera block (the f32
// uniforms of src/types/webgpu.ts:67-116 promoted to f64, plus the per-frame constants of
// compute.wgsl.ts:174-187 evaluated once on the host) and the 1-D disk temperature LUT (physics/disk.rs:175-201).
struct alignas(16) FrameBlock {
    double inv_proj[16];   // column-major
    double inv_view[16];   // column-major
    double r0, theta0, phi0, st, ct, sp, cp, safe_st;
    double inv_width, inv_height, jx, jy;   // 1/W, 1/H, jitter / resolution (compute.wgsl.ts:153-157)
    double cam_pos[3], rph;                 // camera position; prograde photon sphere as the GLSL computes it
    float tdisk[512];
};
static_assert(sizeof(FrameBlock) % 16 == 0, "TMA bulk copies move multiples of 16 bytes");

constexpr uint32_t kMaxSmemLutBytes = 192 * 1024;  // spectral LUT is smem-resident up to this size

struct Counters {  // device-side accumulators, one set per frame
    unsigned long long steps_committed, steps_executed, rhs_evals;
    unsigned long long n_horizon, n_escape, n_maxsteps, n_disk;
    unsigned int tile_counter;  // dynamic warp-tile queue head
    unsigned int _pad;
};

// GVT_FLAG_ROW_INTERLEAVE: which frame rows a rank produces. Stripes of `s` rows are dealt round-robin to the ranks
// (stripe j belongs to rank j % world); with TAA every stripe is traced with `halo` redundant rows on each side so the
// 3x3 resolve stays local. s = 0: off (contiguous rows). Lattice row lj of a launch -> frame row:
struct StripeMap {
    uint32_t s, halo, world, rank;
};
__host__ __device__ inline bool stripe_row(const StripeMap& m, uint32_t lj, uint32_t height, uint32_t& row) {
    const uint32_t sh = m.s + 2u * m.halo, t = lj / sh, o = lj - t * sh;
    const long long r = (long long)(t * m.world + m.rank) * (long long)m.s - (long long)m.halo + (long long)o;
    row = r < 0 ? 0u : (uint32_t)r;
    return r >= 0 && r < (long long)height;
}

// Kernel parameters (constant bank): the scalars of the step loop are direct c[][] operands.
struct FrameParams {
    TrigTable trig;                                           // set by the launcher (GVT_TRIG_TABLE_INIT)
    double M, a, spin, rh, r_term, escape_r, r_in, r_out;   // r_term = 1.001 * r+ (geodesic/mod.rs:257)
    double sqrtM;                                             // sqrt(M) for the Keplerian frequency (redshift.rs:72)
    double a2, twoM;                                          // a * a, 2 M: what HoleRay<double> would otherwise compute per thread
    double tol, h0;
    // Zone radii of the march (chunk-level, warp-uniform specialisations of the step loop; see k_trace_tile):
    double r_hconst;                                          // beyond it the step rule has saturated: h == h_const, and the
    double h_const;                                           //   horizon cannot be reached within one chunk
    double r_escape_guard;                                    // below it the escape radius cannot be reached within one chunk
    double r_sat;                                             // where the step rule saturates (r+ + 1/0.15; r_term for the constant rule)
    double rot_travel, rot_k;                                 // f64 zone 3: a ray enters beyond travel + sqrt(rot_k B), B^2 = Q + a^2; rot_k = 256 h
    float f32_rot_travel, f32_rot_inv_k;
    uint32_t alive_lo[2], alive_span[2];                      // hot-path "alive" window on the high word of r: [0] (r_term, escape_r),
                                                              //   [1] (r_sat, escape_r): inside iff (hi(r) - lo) < span (unsigned)
    double r_far;                                             // zone 2 of GVT_PRECISION_MIXED (f32 predictors) beyond this radius
    double r_rot;                                             // zone 3 (no disk test; f64: rotated trigonometry): beyond r_far AND the disk's outer edge by a chunk's travel
    double rot_q_max;                                         // f64 zone 3 is open to rays with Q + a^2 <= this (a switch: huge, or < 0 = closed)
    double rot_stab;                                          //   and L^2 >= rot_stab (Q + a^2 + L^2)^2 (stable polar turning point out there)
    float f32_M, f32_a, f32_a2, f32_twoM, f32_hconst;         // the predictors' hole constants and (float)h_const
    double tdisk_rin, tdisk_scale;                            // (n-1)/(rout-rin)
    uint32_t width, height;                                   // full frame
    uint32_t x0, xs, y0, y1, ys;                              // pixel lattice traced by this launch
    uint32_t nx, ny;                                          // lattice extent
    uint32_t tile_w_log2, _pad_tile;                          // warp tile = 2^tile_w_log2 x (32 >> tile_w_log2) pixels: 3 (8x4), 4 (16x2), 5 (32x1)
    uint32_t max_steps, renorm_interval, step_rule;
    uint32_t tdisk_n, spec_w, spec_h, lut_in_smem;
    const FrameBlock* block;                                  // global copy of the TMA-staged block
    const float4* spectrum;                                   // global spectral LUT (W*H float4)
    float4* frame;                                            // full-frame RGBA32F (may be null for debug launches)
    float4* host_frame;                                       // device alias of a page-locked HOST frame (or null): the
                                                              // epilogue stores the pixel there too, so no D2H copy follows
    float4* peer_frame[8];                                    // GVT_FLAG_PEER_STORE: the same frame on every other rank
    uint32_t n_peer, frame_f16;                               // frame_f16 != 0: frame / host_frame / peer_frame hold RGBA16F (8 B per pixel)
    StripeMap stripe;                                         // s != 0: lattice rows map to frame rows through stripe_row()
    Counters* counters;
    unsigned long long* timeline;                             // diagnostics (GVT_TIMELINE_DUMP): per warp {globaltimer at start, at end, tiles}, or null
    // parity-hook outputs (DEBUG instantiations only), dense over the lattice
    double* dbg_xp; uint32_t* dbg_term; uint32_t* dbg_steps; double* dbg_drift; double* dbg_rgba;
};

struct RayBatchParams {  // gvt_engine_integrate_rays
    TrigTable trig;
    double M, a, rh, r_term, escape_r, tol, h0;
    uint32_t max_steps, renorm_interval, step_rule, method, coords;
    uint64_t n;
    const double* in_xp; double* out_xp; uint32_t* term; uint32_t* steps; double* drift; uint32_t* rhs;
};

struct TaaParams {
    // ataa.wgsl.ts:54-69: reprojection of a point at depth 12 along the pixel's world ray through prev_view_proj,
    // factored on the host (fill_taa, f64): view-space ray v = vA cx + vB cy + vC (xyz and w of inv_proj * clip);
    // prev clip coords (x, y, w) = c0 + s (gA cx + gB cy + gC), s = sign(v.w) / |v.xyz|, where
    // g* = 12 * prev_view_proj[rows x,y,w][:, :3] * inv_view[:3, :3] * v*  and  c0 = prev_view_proj * (cam_pos, 1).
    float vA[4], vB[4], vC[4];
    float gA[4], gB[4], gC[4], c0[4];
    uint32_t width, height;
    uint32_t mode;         // 0: ataa.wgsl.ts (WebGPU); 1: reprojection.glsl.ts (WebGL2)
    float blend;           // mode 1: u_blendFactor
    uint32_t moving, _pad_taa;   // mode 1: u_cameraMoving
    uint32_t row0, row1;   // rows resolved by this launch (a rank's block); neighbours outside are still read
    const float4* cur; const float4* hist; float4* out;
    float4* host_out;      // device alias of a page-locked host frame, or null
    float4* peer_out[8];   // GVT_FLAG_PEER_STORE targets
    uint32_t n_peer;
    uint32_t frame_f16;    // cur / hist / out (and host_out, peer_out) are RGBA16F: 8 B per pixel, f32 arithmetic
    uint32_t unit_rows;    // rows per warp work unit (chosen by launch_taa)
    StripeMap stripe;      // s != 0: resolve this rank's n_stripes stripes (rows [(t world + rank) s, +s)) instead of [row0, row1)
    uint32_t n_stripes;
    // k_taa_resolve_precise only: the raw uniforms (types/webgpu.ts:67-116), column-major
    float m_inv_proj[16], m_inv_view[16], m_prev_vp[16], cam_pos[4];
};

// k_fragment_glsl (gvt_fragment.cu): the production WebGL2 fragment shader
struct GlslParams {
    GvtGlslUniforms u;            // chunks/common.ts:9-38 uniforms + feature bits, in the parameter bank
    uint32_t width, height;       // frame size (= u.resolution)
    uint32_t y0, y1, ys;          // rows shaded by this launch: y0, y0 + ys, ... < y1 (a rank's block)
    StripeMap stripe;             // s != 0: n_lattice_rows lattice rows map to frame rows through stripe_row() instead
    uint32_t n_lattice_rows;
    const uint8_t* noise_r;       // 256*256 red channel of u_noiseTex (global; TMA-staged into shared memory)
    const uint8_t* blue_r;        // 256*256 red channel of u_blueNoiseTex (one tap per pixel, stays in global)
    float4* frame; float4* host_frame; float4* peer_frame[8];
    uint32_t n_peer;
    uint32_t frame_f16;           // the three frame targets hold RGBA16F
    Counters* counters;
    uint32_t* dbg_steps; uint32_t* dbg_hit;   // parity hooks (full-frame arrays) or null
};

// Function attributes (the dynamic shared-memory opt-in) and occupancy are per DEVICE: a process may hold renderers on
// several GPUs (GvtDeviceConfig.device), so per-kernel launch state is cached per device ordinal, never in a bare static.
// Benign race between threads: both compute the same value.
constexpr int kMaxDevices = 64;
struct PerDeviceInt {
    int v[kMaxDevices] = {};
    int* slot() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) return nullptr;
        return &v[d];
    }
};

// launchers (gvt_kernels.cu)
cudaError_t launch_trace(const FrameParams& p, int method, int precision, bool budget, bool debug, int sm_count,
                         cudaStream_t stream);
cudaError_t launch_integrate_rays(const RayBatchParams& p, cudaStream_t stream);
cudaError_t launch_taa(const TaaParams& p, int sm_count, cudaStream_t stream);
cudaError_t launch_taa_precise(const TaaParams& p, cudaStream_t stream);   // IEEE build, one thread per pixel (row blocks only)
cudaError_t launch_fragment_glsl(const GlslParams& p, int precision, int sm_count, cudaStream_t stream);
cudaError_t launch_bloom(const float4* frame, int W, int H, uint2* half_tex, uint2* q1, uint2* q2, float4* display,
                         float threshold, float intensity, int blur_passes, int enabled, int sm_count, cudaStream_t stream,
                         int* launches, bool precise = false, bool scene_f16 = false);
cudaError_t launch_fragment_glsl_fast(const GlslParams& p, int sm_count, cudaStream_t stream);   // f32, MUFU maths
cudaError_t launch_f32_to_f16(const float4* src, void* dst, size_t n_px, cudaStream_t stream);
cudaError_t launch_f16_to_f32(const void* src, float4* dst, size_t n_px, cudaStream_t stream);
cudaError_t launch_tonemap_rgba8(const float4* src, void* dst, size_t n_px, int aces, cudaStream_t stream, bool src_f16 = false);
cudaError_t launch_fma_peak(int precision, int sm_count, unsigned long long iters, float* sink, cudaStream_t stream,
                            double* flops_out);

}  // namespace gvt
