/* gravitas_b200.h — C ABI of libgravitas_b200.so: the B200-native replacement for the reference's
 * per-pixel Kerr geodesic hot path, behind the reference's own two seams.
 *
 *   Seam A  `PhysicsEngine` (wasm-bindgen class, physics-engine/gravitas-wasm/src/lib.rs:42-465) and the
 *           SharedArrayBuffer f32-offset protocol (lib.rs:36-40, sab.rs:18-22, src/engine/physics-bridge.ts:5-11).
 *   Seam B  the src/rendering renderer / frame-buffer API: the WebGPU pipeline (src/rendering/webgpu/renderer.ts:280
 *           `render(camera, physics)`; uniform layouts src/types/webgpu.ts:25,42,67-116 and
 *           src/shaders/types.wgsl.ts:6-29) and the WebGL2 pipeline (src/rendering/webgl/renderer.ts:173
 *           `render(params, mouse)`: fragment shader -> TAA resolve -> bloom / final pass).
 *
 * Conventions: every entry point returns an int32 status (GVT_OK = 0, negative = error; the reference has no
 * error returns — Rust panics go to console_error_panic_hook, lib.rs:30-33 — so hosts may ignore it exactly as
 * they do today, or read gvt_last_error()). No exceptions cross the boundary. Handles are opaque. The caller
 * owns every buffer it passes. Thread-compatible, not thread-safe: one handle per thread (the worker model of
 * src/workers/physics.worker.ts). All per-pixel / per-ray arithmetic runs in hand-written sm_100a CUDA kernels; there
 * is no CPU fallback — creation fails with GVT_ERR_NO_DEVICE when no CUDA device is usable. (The engine's once-per-tick
 * closed forms, LUT generators and visualisation helpers are host maths, as they are one-off host calls in the reference.)
 */
#ifndef GRAVITAS_B200_H
#define GRAVITAS_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GVT_ABI_VERSION 1

/* status codes */
enum {
    GVT_OK = 0,
    GVT_ERR_INVALID = -1,     /* null pointer / bad argument */
    GVT_ERR_NO_DEVICE = -2,   /* no usable CUDA device (the product has no CPU path) */
    GVT_ERR_CUDA = -3,        /* CUDA runtime error, see gvt_last_error() */
    GVT_ERR_NCCL = -4,        /* NCCL error / libnccl not loadable for world_size > 1 */
    GVT_ERR_UNSUPPORTED = -5
};

/* SAB v2 layout: f32 ELEMENT indices (lib.rs:36-40; sab.rs:18-22; physics-bridge.ts:5-11) */
#define GVT_SAB_OFFSET_CONTROL 0
#define GVT_SAB_OFFSET_CAMERA 64
#define GVT_SAB_OFFSET_PHYSICS 128
#define GVT_SAB_OFFSET_TELEMETRY 256
#define GVT_SAB_OFFSET_LUTS 2048
#define GVT_SAB_INTERNAL_F32 2048 /* engine-owned buffer when no SAB is attached (lib.rs:67) */

/* geodesic/termination.rs:4-17 (#[repr(C)]) */
enum { GVT_TERM_NONE = 0, GVT_TERM_HORIZON = 1, GVT_TERM_ESCAPE = 2, GVT_TERM_MAXSTEPS = 3, GVT_TERM_DISK = 4 };
/* metric/kerr.rs:16-22 */
enum { GVT_COORDS_BOYER_LINDQUIST = 0, GVT_COORDS_KERR_SCHILD = 1 };
/* geodesic/integrator.rs:14-21 */
enum { GVT_METHOD_RKF45 = 0, GVT_METHOD_RK4 = 1, GVT_METHOD_SYMPLECTIC = 2,
       /* render path only: the production WebGL2 shader's Cartesian Velocity-Verlet march on its pseudo-Kerr acceleration
          (fragment.glsl.ts:129-221, chunks/metric.ts:96-149); max_steps is capped at 500 as there (:115); escape_radius
          plays MAX_DIST. Thin-disk light comes from the same LUT composite as the Hamiltonian path. */
       GVT_METHOD_VERLET_GLSL = 3 };
enum {
    GVT_PRECISION_F64 = 0,
    GVT_PRECISION_F32 = 1,
    GVT_PRECISION_F32_FAST = 2, /* gvt_render_fragment_glsl only: f32 with MUFU rcp/sqrt/exp2/log2/sin/cos, as a GLSL compiler builds the shader */
    GVT_PRECISION_MIXED = 3     /* gvt_render_frame / gvt_trace_states with GVT_METHOD_SYMPLECTIC only: f64 state and f64
                                 * final (corrector) evaluation of the implicit-midpoint step (integrator.rs:209-226); its
                                 * two fixed-point predictor evaluations run in f32 while a ray is on its way out beyond
                                 * 35 M, where their error reaches the state damped by h |J| / 2 < 5e-4.
                                 * RGBA stays within 1e-6 relative of the all-f64 scheme on every pixel (tested). */
};
enum {
    GVT_FORMAT_RGBA32F = 0,
    GVT_FORMAT_RGBA16F = 1,        /* reprojection.ts:120-140 / webgpu/renderer.ts:161-180 texture format (linear HDR) */
    GVT_FORMAT_RGBA8_REINHARD = 2, /* display-ready: the WebGPU blit's Reinhard map c/(1+c) (webgpu/renderer.ts:45-47), 8-bit unorm */
    GVT_FORMAT_RGBA8_ACES = 3,     /* display-ready: the WebGL final pass without bloom (bloom.glsl.ts:106-124): ACES
                                      (Narkowicz) then pow(., 0.4545), 8-bit unorm */
    GVT_FORMAT_RGBA8_UNORM = 4     /* clamp to [0,1] and quantise only: for frames that are already display-referred (the
                                    * fragment shader applies ACES + gamma itself unless GVT_GLSL_LINEAR_OUTPUT) */
};
/* step rule for the fixed-step methods: constant `initial_step` (geodesic/mod.rs:218-223) or the per-step rule
 * h = clamp(0.15 (r - r+), 0.05, 1.0) of src/shaders/compute.wgsl.ts:213 */
enum { GVT_STEP_CONSTANT = 0, GVT_STEP_WGSL = 1 };

enum {
    GVT_FLAG_JITTER = 1u << 0,      /* Halton(2,3) sub-pixel jitter on frame_index (compute.wgsl.ts:153-157) */
    GVT_FLAG_BUDGET = 1u << 1,      /* budget accounting: every pixel executes exactly max_steps step computations;
                                       terminated rays keep stepping with commits masked */
    /* bit 2 reserved (per-ray max |H| of geodesic/mod.rs:233-237 is reported by gvt_trace_states and
       gvt_engine_integrate_rays, not by the frame path) */
    GVT_FLAG_TAA = 1u << 3,         /* run the TAA resolve (ataa.wgsl.ts:28-83) after the trace */
    GVT_FLAG_NO_GATHER = 1u << 4,   /* multi-GPU: skip the all-gather (each rank keeps only its row block) */
    GVT_FLAG_TAA_WEBGL = 1u << 7,   /* with GVT_FLAG_TAA: the WebGL2 resolve of src/shaders/postprocess/reprojection.glsl.ts:70-115
                                       (mu +- 1.5 sigma clip, no reprojection, alpha = moving ? 0 : blend (1 - clamp(4 sigma_Y,
                                       0, 0.55))) instead of the WebGPU one (ataa.wgsl.ts:28-83) */
    GVT_FLAG_PEER_STORE = 1u << 6,  /* multi-GPU: fuse the gather into the producing kernel — every finished pixel is stored
                                       into every peer's frame over NVLink (peer frames imported with
                                       gvt_render_import_peer_frames) and a 4-byte ncclAllReduce closes the frame as the
                                       cross-GPU barrier; replaces the ncclAllGather */
    GVT_FLAG_ROW_INTERLEAVE = 1u << 8, /* multi-GPU, with GVT_FLAG_PEER_STORE: stripes of rows are dealt round-robin to the
                                       ranks instead of one contiguous block each, so that rows of very different cost
                                       (sky / disk / shadow under natural termination) spread evenly over the GPUs
                                       (SURVEY 8e). Without TAA a stripe is one row (rank k: rows k, k + world, ...); with
                                       TAA it is 16 rows traced with a one-row halo. The peer stores need no blocks. */
    GVT_FLAG_DEBUG_COUNTS = 1u << 9, /* gvt_render_fragment_glsl: also record per-pixel march steps and horizon flags for
                                       gvt_render_fragment_glsl_debug (8 B per pixel of extra stores; off by default) */
    GVT_FLAG_TAA_PRECISE = 1u << 10, /* with GVT_FLAG_TAA: run the validation build of the resolve (every operation an IEEE
                                       round-to-nearest f32 operation in shader order, unfolded reprojection) instead of
                                       the production kernel (MUFU sqrt/rcp/rsq, host-folded matrices). Row blocks only. */
    GVT_FLAG_D2H_OWN_ROWS = 1u << 5 /* multi-GPU: copy only this rank's row block into host_rgba (at its place in the
                                       full-size buffer). With one host frame shared by all ranks (POSIX shm registered
                                       through gvt_host_register) the ranks assemble the frame in parallel, one
                                       frame's worth of PCIe traffic in total instead of one per rank. */
};

/* src/types/webgpu.ts:67-116 / src/shaders/types.wgsl.ts:6-16 — 352 bytes, column-major mat4 as gl-matrix */
typedef struct GvtCamera {
    float view[16];
    float proj[16];
    float inv_view[16];
    float inv_proj[16];
    float prev_view_proj[16];
    float position[4];  /* xyz + pad */
    float direction[4]; /* xyz + pad */
} GvtCamera;

/* src/types/webgpu.ts:42-64 / types.wgsl.ts:19-29 — 32 bytes */
typedef struct GvtPhysicsParams {
    float mass;
    float spin; /* dimensionless a*, clamped to [-1,1] like Kerr::new (kerr.rs:49-55) */
    float resolution[2];
    float time;
    float dt;
    uint32_t frame_index;
    uint32_t _pad;
} GvtPhysicsParams;

/* What the WGSL bakes in as constants / overrides (compute.wgsl.ts:11-15) plus IntegrationOptions
 * (geodesic/integrator.rs:24-47; values used by the wasm seam: lib.rs:444-452). */
typedef struct GvtRenderParams {
    uint32_t struct_size;          /* sizeof(GvtRenderParams), for forward compatibility */
    uint32_t method;               /* GVT_METHOD_* */
    uint32_t precision;            /* GVT_PRECISION_* */
    uint32_t coords;               /* GVT_COORDS_* (render path: Kerr-Schild, lib.rs:64,454-455) */
    uint32_t step_rule;            /* GVT_STEP_* */
    uint32_t max_steps;            /* MAX_STEPS override (compute.wgsl.ts:13) / IntegrationOptions.max_steps */
    uint32_t renormalize_interval; /* integrator.rs:44 (10) */
    uint32_t flags;                /* GVT_FLAG_* */
    uint32_t output_format;        /* GVT_FORMAT_* of the host frame buffer */
    uint32_t _pad;
    double tolerance;              /* RKF45 local error tolerance (1e-8) */
    double initial_step;           /* h0 (0.01) or the constant step */
    double escape_radius;          /* 1000 (lib.rs:449) */
    double disk_r_out;             /* thin-disk outer edge, 50 M (physics/disk.rs:177); inner = ISCO prograde */
    float taa_blend;               /* GVT_FLAG_TAA_WEBGL: u_blendFactor (0.75 in webgl/renderer.ts:380-385) */
    uint32_t taa_camera_moving;    /* GVT_FLAG_TAA_WEBGL: u_cameraMoving */
} GvtRenderParams;

/* The production WebGL2 fragment shader's uniforms (src/shaders/blackhole/chunks/common.ts:9-38), with the values
 * webgl/renderer.ts:296-358 uploads, plus the shader manager's #defines (src/shaders/manager.ts:55-82) as bits.
 * Conventions GLSL leaves to the implementation, fixed here: u_noiseTex is sampled LINEAR + REPEAT with exact bilinear
 * weights, u_blueNoiseTex NEAREST + REPEAT (webgl-utils.ts:259-303), only their red channels are read (as the shader
 * does); normalize(v) = v / sqrt(dot(v, v)); row j of the output frame is gl_FragCoord.y = j + 0.5 (bottom-up). */
enum {
    GVT_GLSL_LENSING = 1,        /* ENABLE_LENSING */
    GVT_GLSL_DISK = 2,           /* ENABLE_DISK */
    GVT_GLSL_JETS = 4,           /* ENABLE_JETS (the manager emits it only together with the disk) */
    GVT_GLSL_STARS = 8,          /* ENABLE_STARS */
    GVT_GLSL_PHOTON_GLOW = 16,   /* ENABLE_PHOTON_GLOW */
    GVT_GLSL_DOPPLER = 32,       /* ENABLE_DOPPLER */
    GVT_GLSL_REDSHIFT = 64,      /* ENABLE_REDSHIFT */
    GVT_GLSL_LINEAR_OUTPUT = 128,/* ENABLE_LINEAR_OUTPUT: skip ACES + gamma (the bloom pipeline tone-maps later) */
    GVT_GLSL_QUALITY_LOW = 256   /* RAY_QUALITY_LOW / RAY_QUALITY_OFF: no march */
};
typedef struct GvtGlslUniforms {
    uint32_t struct_size;        /* sizeof(GvtGlslUniforms) */
    uint32_t features;           /* GVT_GLSL_* */
    float resolution[2];         /* u_resolution (pixels) */
    float time;                  /* u_time */
    float mass;                  /* u_mass */
    float spin;                  /* u_spin = a* x mass (renderer.ts:326) */
    float disk_density;          /* u_disk_density (default 4) */
    float disk_temp;             /* u_disk_temp = T x mass^-1/4 (renderer.ts:351-354) */
    float mouse[2];              /* u_mouse: azimuth/2pi, polar/pi */
    float zoom;                  /* u_zoom = 2 x params.zoom (renderer.ts:327) */
    float lensing_strength;      /* u_lensing_strength */
    float disk_size;             /* u_disk_size (default 50) */
    float disk_scale_height;     /* u_disk_scale_height (default 0.2) */
    int32_t max_ray_steps;       /* u_maxRaySteps (capped at 500 by the shader) */
    float debug;                 /* u_debug */
    float show_redshift;         /* u_show_redshift */
    float show_kerr_shadow;      /* u_show_kerr_shadow */
    float shadow_count;          /* u_shadowCount */
    float cam_pos[3];            /* u_camPos: (0,0,0) selects the mouse/zoom camera, as renderer.ts:314 does */
    float cam_quat[4];           /* u_camQuat xyzw */
    float shadow_curve[128];     /* u_shadowCurve: 64 (alpha, beta) pairs, SAB PHYSICS block [16..143] */
} GvtGlslUniforms;

typedef struct GvtDeviceConfig {
    uint32_t struct_size;
    int32_t device;        /* CUDA device ordinal */
    int32_t rank;          /* row-block shard index, 0..world_size-1 */
    int32_t world_size;    /* 1 = single GPU */
    uint8_t nccl_id[128];  /* ncclUniqueId from gvt_nccl_unique_id() on rank 0, broadcast by the host */
} GvtDeviceConfig;

typedef struct GvtFrameStats {
    double trace_ms;          /* device time of the fused trace kernel (CUDA events on the launch stream) */
    double taa_ms;            /* device time of the TAA resolve (0 when off) */
    double gather_ms;         /* device time of the all-gather (0 at world_size 1) */
    double total_ms;          /* first launch -> last device op of the frame, incl. copies when host buffers are used */
    uint64_t steps_committed; /* accepted geodesic steps summed over this rank's pixels (device-counted) */
    uint64_t steps_executed;  /* step computations executed (= pixels*max_steps in budget mode) */
    uint64_t rhs_evals;       /* Hamiltonian RHS evaluations */
    uint64_t n_horizon, n_escape, n_maxsteps, n_disk; /* termination census */
    uint64_t h2d_bytes, d2h_bytes;
    uint32_t kernel_launches; /* kernels of this library launched for the frame */
    uint32_t rows_begin, rows_end; /* this rank's row block */
} GvtFrameStats;

const char* gvt_last_error(void);
int32_t gvt_abi_version(void);
int32_t gvt_device_count(int32_t* out);
/* Identifies this build of the library: a hash of the kernel / API sources it was compiled from plus the launch
 * geometry of the trace kernel, e.g. "src=3f9a12c4e07b maxt_f64=512 maxt_f32=512". bench.py compares it with the
 * hash recorded in the committed ncu exports so that profile-derived figures (DRAM traffic, executed instruction mix)
 * are never reported for a kernel they were not measured on. Returns a pointer to a static string. */
const char* gvt_build_info(void);

/* ---- Seam A: PhysicsEngine (gravitas-wasm/src/lib.rs) -------------------------------------------------- */
typedef struct gvt_engine gvt_engine;

int32_t gvt_engine_create(double mass, double spin, gvt_engine** out);            /* lib.rs:59-72 `new` */
int32_t gvt_engine_destroy(gvt_engine* e);                                          /* wasm-bindgen `.free()` */
int32_t gvt_engine_update_params(gvt_engine* e, double mass, double spin);         /* lib.rs:78-83 */
int32_t gvt_engine_attach_sab(gvt_engine* e, float* sab);                           /* lib.rs:74-76 */
int32_t gvt_engine_get_sab_ptr(gvt_engine* e, const float** out);                   /* lib.rs:116-118 */
int32_t gvt_engine_get_sab_layout(gvt_engine* e, uint32_t out5[5]);                 /* lib.rs:411-419 */
int32_t gvt_engine_tick_sab(gvt_engine* e, double dt_override);                     /* lib.rs:308-409 */
int32_t gvt_engine_set_camera_state(gvt_engine* e, double px, double py, double pz,
                                    double lx, double ly, double lz);               /* lib.rs:120-122 */
int32_t gvt_engine_set_auto_spin(gvt_engine* e, int32_t enabled);                   /* lib.rs:124-126 */
int32_t gvt_engine_compute_horizon(gvt_engine* e, double* out);                     /* lib.rs:85-87 */
int32_t gvt_engine_compute_isco(gvt_engine* e, double* out);                        /* lib.rs:89-91 */
int32_t gvt_engine_compute_photon_sphere(gvt_engine* e, double* out);               /* lib.rs:93-95 */
int32_t gvt_engine_compute_dilation(gvt_engine* e, double r, double* out);          /* lib.rs:97-105 */
int32_t gvt_engine_compute_g_factor(gvt_engine* e, double r, double lambda, double* out); /* redshift.rs:65-95 */
/* Bardeen critical curve as [alpha0, beta0, alpha1, beta1, ...] f32 pairs (lib.rs:161-170; shadow.rs:81-183). The
 * curve has n_points points for a* = 0, 2 n_points otherwise; `capacity_pairs` bounds what is written, *n_pairs
 * receives the curve's full length (call with out = NULL to size the buffer). */
int32_t gvt_engine_compute_shadow_curve(gvt_engine* e, double theta_obs, uint32_t n_points, float* out_pairs,
                                        uint32_t capacity_pairs, uint32_t* n_pairs);
int32_t gvt_engine_compute_shadow_radius(gvt_engine* e, double* out);               /* lib.rs:173-175 */
int32_t gvt_engine_compute_shadow_shift(gvt_engine* e, double theta_obs, float out2[2]); /* lib.rs:179-196 */
int32_t gvt_engine_compute_disk_flux(gvt_engine* e, double r, double* out);         /* lib.rs:199-201 */
/* Spacetime-visualisation helpers of the class (lib.rs:139-305 over gravitas-core/src/spacetime/): one-off host maths on the
 * Boyer-Lindquist metric, as in the reference. Fields are (r, theta, value) f32 triples over n_radial x n_polar samples
 * (r linear in [r_min, r_max], theta in [0.1, pi - 0.1]); meshes are (x, y, z) f32 triples. n_radial / n_polar >= 2. */
int32_t gvt_engine_compute_kretschner(gvt_engine* e, double r, double theta, double* out);        /* lib.rs:213-215 */
int32_t gvt_engine_generate_curvature_field(gvt_engine* e, double r_min, double r_max, uint32_t n_radial, uint32_t n_polar,
                                            float* out3);                                          /* lib.rs:219-234 */
int32_t gvt_engine_compute_light_cone_tilt(gvt_engine* e, double r, double theta, double* out);   /* lib.rs:238-240 */
int32_t gvt_engine_generate_tilt_field(gvt_engine* e, double r_min, double r_max, uint32_t n_radial, uint32_t n_polar,
                                       float* out3);                                               /* lib.rs:244-263 */
int32_t gvt_engine_compute_frame_drag_omega(gvt_engine* e, double r, double theta, double* out);  /* lib.rs:267-269 */
int32_t gvt_engine_generate_frame_drag_field(gvt_engine* e, double r_min, double r_max, uint32_t n_radial, uint32_t n_polar,
                                             float* out3);                                         /* lib.rs:273-292 */
int32_t gvt_engine_compute_flamm_height(gvt_engine* e, double r, double* out);                    /* lib.rs:296-298 */
int32_t gvt_engine_compute_proper_distance(gvt_engine* e, double r1, double r2, uint32_t n_steps, double* out); /* :302-304 */
int32_t gvt_engine_generate_embedding_mesh(gvt_engine* e, double r_min, double r_max, uint32_t n_radial, uint32_t n_angular,
                                           float* out3);                                           /* lib.rs:139-150 */
int32_t gvt_engine_generate_ergosphere_mesh(gvt_engine* e, uint32_t n_polar, uint32_t n_azimuthal, float* out3); /* :153-157 */
int32_t gvt_engine_generate_disk_lut(gvt_engine* e, float* out512);                 /* lib.rs:107-110 */
/* the engine-owned copy the last generate_disk_lut left behind (lib.rs:112-114); NULL / 0 before the first call */
int32_t gvt_engine_get_disk_lut_ptr(gvt_engine* e, const float** out, uint32_t* n);
int32_t gvt_engine_generate_spectrum_lut(gvt_engine* e, uint32_t width, uint32_t height, double max_temp,
                                         float* out_rgba);                          /* lib.rs:128-136 */
/* lib.rs:422-464 integrate_ray_relativistic: RKF45, h0 0.01, escape 1000, renorm 10. Runs on the GPU.
 * term / steps_taken / max_drift may be NULL. */
int32_t gvt_engine_integrate_ray(gvt_engine* e, const double in8[8], uint64_t steps, double tolerance,
                                 int32_t use_kerr_schild, double out8[8], uint32_t* term, uint64_t* steps_taken,
                                 double* max_drift);
/* Batched form of the same call (n rays, xp[n][8]); geodesic::integrate (geodesic/mod.rs:180-253) with full
 * IntegrationOptions. Output arrays other than out_xp may be NULL. */
int32_t gvt_engine_integrate_rays(gvt_engine* e, const GvtRenderParams* opts, uint64_t n, const double* in_xp,
                                  double* out_xp, uint32_t* term, uint32_t* steps_taken, double* max_drift,
                                  uint32_t* rhs_evals);

/* ---- Seam B: renderer / frame buffer ------------------------------------------------------------------- */
typedef struct gvt_renderer gvt_renderer;

int32_t gvt_nccl_unique_id(uint8_t out128[128]);
int32_t gvt_render_params_default(GvtRenderParams* p); /* config-3 defaults: symplectic, f64, KS, WGSL rule, 512 */
int32_t gvt_render_create(const GvtDeviceConfig* cfg, gvt_renderer** out);          /* WebGPURenderer.init */
int32_t gvt_render_destroy(gvt_renderer* r);
/* (Re)build the spectral LUT (physics/spectrum.rs:76-102) and disk temperature LUT (physics/disk.rs:175-201)
 * for (mass, spin) on the host and upload them. Mirrors SpectralManager.initialize (rendering/spectral.ts:21-61). */
int32_t gvt_render_init_luts(gvt_renderer* r, double mass, double spin, uint32_t spec_w, uint32_t spec_h,
                             double max_temp);
/* Upload caller-provided LUTs instead (spectrum: w*h*4 f32; tdisk: n f32 over [rin, rout]). */
int32_t gvt_render_set_luts(gvt_renderer* r, const float* spectrum, uint32_t spec_w, uint32_t spec_h,
                            const float* tdisk, uint32_t tdisk_n, double tdisk_rin, double tdisk_rout);
int32_t gvt_render_resize(gvt_renderer* r, uint32_t width, uint32_t height);        /* renderer.ts:269-278 */
/* renderer.ts:280 `render(camera, physics)`: trace (+TAA) (+all-gather); the finished frame is copied into
 * host_rgba (width*height*4 f32 or f16) when it is non-NULL, else it stays in device memory. */
int32_t gvt_render_frame(gvt_renderer* r, const GvtCamera* cam, const GvtPhysicsParams* phys,
                         const GvtRenderParams* params, void* host_rgba, GvtFrameStats* stats);
/* The row block [row0, row1) of the same frame on a single-GPU renderer (what one rank of an N-GPU run traces): for hosts
 * that tile a frame themselves, and for measuring a rank's share of a frame on one device. The rest of the frame buffer is
 * left as it was; host_rgba (may be NULL) receives the WHOLE frame buffer. No TAA / interleave flags. */
int32_t gvt_render_rows(gvt_renderer* r, const GvtCamera* cam, const GvtPhysicsParams* phys, const GvtRenderParams* params,
                        uint32_t row0, uint32_t row1, void* host_rgba, GvtFrameStats* stats);
int32_t gvt_render_read_frame(gvt_renderer* r, uint32_t format, void* host_rgba);
/* Parity hook: per-pixel final state (x,p)[8] f64, termination, accepted steps, max|H|, f64 RGBA over the pixel
 * lattice x = x0 + i*xs, y = y0 + j*ys (y < y1). Any output may be NULL. */
int32_t gvt_trace_states(gvt_renderer* r, const GvtCamera* cam, const GvtPhysicsParams* phys,
                         const GvtRenderParams* params, uint32_t x0, uint32_t xs, uint32_t y0, uint32_t y1,
                         uint32_t ys, double* out_xp8, uint32_t* term, uint32_t* steps, double* max_drift,
                         double* rgba64);
/* TAA resolve on caller-provided frames (ataa.wgsl.ts:28-83): cur, hist, out are width*height*4 f32 host buffers. */
int32_t gvt_taa_resolve(gvt_renderer* r, const GvtCamera* cam, uint32_t width, uint32_t height, const float* cur,
                        const float* hist, float* out);
/* The WebGL2 variant (reprojection.glsl.ts:70-115) on caller-provided frames. */
int32_t gvt_taa_resolve_webgl(gvt_renderer* r, uint32_t width, uint32_t height, const float* cur, const float* hist,
                              float blend, int32_t camera_moving, float* out);
/* Both variants with two more knobs: precise != 0 runs the validation build (GVT_FLAG_TAA_PRECISE: IEEE f32 operations
 * in shader order; tested to 1e-6 against the numpy restatement of the shader text), ms_out (may be NULL) receives the
 * kernel's device time. webgl != 0: reprojection.glsl.ts (cam may be NULL), else ataa.wgsl.ts. */
int32_t gvt_taa_resolve_ex(gvt_renderer* r, const GvtCamera* cam, uint32_t width, uint32_t height, const void* cur,
                           const void* hist, void* out, uint32_t frame_format /* GVT_FORMAT_RGBA32F | RGBA16F: the three buffers */,
                           uint32_t webgl, float blend, int32_t camera_moving, int32_t precise, double* ms_out);
int32_t gvt_render_reset_history(gvt_renderer* r);
/* Format of the frame chain (trace output, finished frame, TAA history): GVT_FORMAT_RGBA32F (default; the parity format) or
 * GVT_FORMAT_RGBA16F, the reference's own texture format (rendering/reprojection.ts:120-140, webgpu/renderer.ts:161-180):
 * the producing kernels store half4, the TAA resolve and the bloom passes read / write 8 B per pixel (f32 arithmetic), and
 * an RGBA16F host frame is delivered without a conversion pass. Re-allocates the buffers and clears the history; with
 * peer stores, re-import the peers' frames afterwards. */
int32_t gvt_render_set_frame_format(gvt_renderer* r, uint32_t format);
/* Size of the frame buffers (0 x 0 before the first resize / frame): what a binding needs to validate caller buffers. */
int32_t gvt_render_get_size(gvt_renderer* r, uint32_t* width, uint32_t* height);

/* ---- Seam B, WebGL2 pipeline: WebGLRenderer.render(params, mouse) (src/rendering/webgl/renderer.ts:173-420) ----
 * The two 256x256 RGBA8 noise textures the reference fills with Math.random() at init (webgl-utils.ts:259-303);
 * here the caller supplies them, so a frame is a pure function of its inputs. */
int32_t gvt_render_set_noise_textures(gvt_renderer* r, const uint8_t* noise_rgba8, const uint8_t* blue_noise_rgba8,
                                      uint32_t size /* 256 */);
/* One frame of the production fragment shader (fragment.glsl.ts) into the renderer's frame buffer (RGBA32F; alpha 1)
 * and, if host_rgba is not NULL, into the caller's buffer in `output_format`. precision: GVT_PRECISION_F32 is the
 * shader's own arithmetic with IEEE division / 1-2 ulp libm (the parity build), GVT_PRECISION_F32_FAST the same source
 * on MUFU approximations (what a GL driver generates), GVT_PRECISION_F64 the rounding-insensitive reference. flags: GVT_FLAG_TAA | GVT_FLAG_TAA_WEBGL runs the WebGL2 TAA resolve after it
 * (taa_blend / taa_camera_moving as in GvtRenderParams), plus the multi-GPU flags of gvt_render_frame. */
int32_t gvt_render_fragment_glsl(gvt_renderer* r, const GvtGlslUniforms* u, uint32_t precision, uint32_t flags,
                                 uint32_t output_format, float taa_blend, uint32_t taa_camera_moving, void* host_rgba,
                                 GvtFrameStats* stats);
/* BloomManager (src/rendering/bloom.ts:22-39 config, :446-632 applyBloomToTexture / drawTextureToScreen) on the
 * renderer's finished linear-HDR frame: bright pass -> `blur_passes` x (horizontal, vertical) 9-tap Gaussian at quarter
 * resolution in RGBA16F -> scene + bloom * intensity -> ACES -> gamma (postprocess/bloom.glsl.ts). enabled = 0 is
 * drawTextureToScreen: the final ACES + gamma pass alone. The display-referred result stays on the device (it does
 * not replace the frame, which remains the TAA history) and is copied to host_out in `output_format`
 * (RGBA32F, RGBA16F or RGBA8_UNORM) when host_out is not NULL. *ms receives the device time. */
typedef struct GvtBloomConfig {
    uint32_t struct_size;
    uint32_t enabled;      /* features.bloom */
    float intensity;       /* 0.5 */
    float threshold;       /* 0.8 */
    uint32_t blur_passes;  /* 2 */
    uint32_t precise;      /* validation build: IEEE f32 operations in GLSL order, libm powf (tested to 1e-6 against the numpy
                              restatement of bloom.glsl.ts); 0 = production build (FMA contraction, MUFU rcp/lg2/ex2).
                              Read only when struct_size covers it. */
} GvtBloomConfig;
int32_t gvt_render_bloom(gvt_renderer* r, const GvtBloomConfig* cfg, uint32_t output_format, void* host_out, double* ms);
/* Parity hook: per-pixel step count and horizon flag of the last gvt_render_fragment_glsl frame rendered with
 * GVT_FLAG_DEBUG_COUNTS (width*height each). */
int32_t gvt_render_fragment_glsl_debug(gvt_renderer* r, uint32_t* steps, uint32_t* hit);
/* Peer-store gather (GVT_FLAG_PEER_STORE): each rank exports CUDA IPC handles of its two frame buffers; the host
 * exchanges them (any transport) and every rank imports every peer's pair. Call again after gvt_render_resize. All
 * ranks must issue the same sequence of gvt_render_frame calls (the buffers ping-pong in lockstep under TAA). */
#define GVT_MAX_PEERS 8
int32_t gvt_render_export_frames(gvt_renderer* r, uint8_t handles[2][64]);
int32_t gvt_render_import_peer_frames(gvt_renderer* r, int32_t peer_rank, const uint8_t handles[2][64]);
/* Pinned host memory for frame buffers (what an N-API external ArrayBuffer would wrap). */
int32_t gvt_host_alloc(size_t bytes, void** out);
int32_t gvt_host_free(void* p);
/* Page-lock caller-owned host memory (e.g. a shared-memory frame all ranks of a box write into). */
int32_t gvt_host_register(void* p, size_t bytes);
int32_t gvt_host_unregister(void* p);
/* Frame targets in DEVICE memory -- the hand-off without host memory. Wherever a frame entry point takes
 * a host buffer (gvt_render_frame, gvt_render_rows, gvt_render_fragment_glsl, gvt_render_read_frame, gvt_render_bloom) it
 * also takes a device pointer: when the requested format is the frame chain's own (and this rank produces every pixel it
 * delivers) the producing kernel stores each finished pixel straight into it, otherwise one device-to-device copy follows;
 * GvtFrameStats.d2h_bytes then counts no frame bytes. The reference keeps its frames on the GPU as RGBA16F textures
 * (src/rendering/reprojection.ts:120-140, webgpu/renderer.ts:161-180); this is the equivalent.
 *   gvt_external_import_fd          maps memory another API exported as a POSIX fd (VK_KHR_external_memory_fd,
 *                                   GL_EXT_memory_object_fd, or a CUDA VMM allocation: cuMemExportToShareableHandle) with
 *                                   cudaImportExternalMemory + cudaExternalMemoryGetMappedBuffer; on success the driver owns
 *                                   the fd. `dedicated` = the allocation is dedicated to one image (VkMemoryDedicatedAllocateInfo).
 *   gvt_external_semaphore_import_fd  the same for a binary (timeline = 0) or timeline semaphore
 *                                   (VK_KHR_external_semaphore_fd / GL_EXT_semaphore_fd).
 *   gvt_render_wait_external / gvt_render_signal_external  enqueue a wait / a signal on the renderer's stream: wait before a
 *                                   frame call (the presenter has finished sampling the target), signal after it. */
typedef struct gvt_external_buffer gvt_external_buffer;
typedef struct gvt_external_semaphore gvt_external_semaphore;
int32_t gvt_external_import_fd(int32_t device, int32_t fd, uint64_t bytes, int32_t dedicated, gvt_external_buffer** out, void** device_ptr);
int32_t gvt_external_release(gvt_external_buffer* b);
int32_t gvt_external_semaphore_import_fd(int32_t device, int32_t fd, int32_t timeline, gvt_external_semaphore** out);
int32_t gvt_external_semaphore_release(gvt_external_semaphore* s);
int32_t gvt_render_wait_external(gvt_renderer* r, gvt_external_semaphore* s, uint64_t value);
int32_t gvt_render_signal_external(gvt_renderer* r, gvt_external_semaphore* s, uint64_t value);
/* In-run FMA-pipe micro-benchmarks (dependent-chain FFMA / DFMA, all SMs): the FP32/FP64 roofline denominators
 * MEASURED_PEAKS.json does not carry. Returns TFLOP/s (2 flop per FMA). */
int32_t gvt_measure_fma_peak(gvt_renderer* r, int32_t precision, double* out_tflops, double* out_ms);
int32_t gvt_device_info(gvt_renderer* r, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, char name[256]);

#ifdef __cplusplus
}
#endif
#endif /* GRAVITAS_B200_H */
