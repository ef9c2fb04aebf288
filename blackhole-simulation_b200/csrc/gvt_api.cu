// gvt_api.cu — the C ABI of libgravitas_b200.so (include/gravitas_b200.h): PhysicsEngine seam + renderer seam.
// Host-side orchestration only: streams, events, pinned staging, NCCL (dlopen'd), LUT upload. All per-pixel /
// per-ray arithmetic is in gvt_kernels.cu.
#include "../../include/gravitas_b200.h"
#include "gvt_hostmath.h"
#include "gvt_internal.h"

#include <cuda_fp16.h>
#include <dlfcn.h>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

using namespace gvt;

// --------------------------------------------------------------------------------------------------
// errors
// --------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int32_t fail(int32_t code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) return fail(GVT_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

extern "C" const char* gvt_last_error(void) { return g_err.c_str(); }
extern "C" int32_t gvt_abi_version(void) { return GVT_ABI_VERSION; }
#ifndef GVT_SRC_HASH
#define GVT_SRC_HASH "unknown"
#endif
#ifndef GVT_MAXT_F64
#define GVT_MAXT_F64 512
#endif
#ifndef GVT_MAXT_F32
#define GVT_MAXT_F32 512
#endif
#define GVT_STR2(x) #x
#define GVT_STR(x) GVT_STR2(x)
extern "C" const char* gvt_build_info(void) {
    return "src=" GVT_SRC_HASH " maxt_f64=" GVT_STR(GVT_MAXT_F64) " maxt_f32=" GVT_STR(GVT_MAXT_F32);
}

extern "C" int32_t gvt_device_count(int32_t* out) {
    if (!out) return fail(GVT_ERR_INVALID, "null out");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *out = 0; return fail(GVT_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    *out = n;
    return GVT_OK;
}

namespace {
// scoped device allocation for the one-shot helpers (no leak on an early error return)
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes); }
    template <class T> T* as() const { return static_cast<T*>(p); }
};
}  // namespace

// --------------------------------------------------------------------------------------------------
// NCCL through dlopen (so a single-GPU host needs no libnccl at all)
// --------------------------------------------------------------------------------------------------
namespace {
struct NcclId { char internal[128]; };
typedef struct ncclComm* ncclComm_t_;
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(ncclComm_t_*, int, NcclId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t_) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t_, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t_, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*CommGetAsyncError)(ncclComm_t_, int*) = nullptr;   // optional (NCCL >= 2.4)
    bool load() {
        if (handle) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (handle) break; }
        if (!handle) return false;
        GetUniqueId = (int (*)(NcclId*))dlsym(handle, "ncclGetUniqueId");
        CommInitRank = (int (*)(ncclComm_t_*, int, NcclId, int))dlsym(handle, "ncclCommInitRank");
        CommDestroy = (int (*)(ncclComm_t_))dlsym(handle, "ncclCommDestroy");
        AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t_, cudaStream_t))dlsym(handle, "ncclAllGather");
        AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t_, cudaStream_t))dlsym(handle, "ncclAllReduce");
        GetErrorString = (const char* (*)(int))dlsym(handle, "ncclGetErrorString");
        CommGetAsyncError = (int (*)(ncclComm_t_, int*))dlsym(handle, "ncclCommGetAsyncError");
        return GetUniqueId && CommInitRank && CommDestroy && AllGather && AllReduce && GetErrorString;
    }
};
NcclApi g_nccl;
constexpr int kNcclFloat32 = 7;  // ncclDataType_t::ncclFloat32
constexpr int kNcclSum = 0;      // ncclRedOp_t::ncclSum
}  // namespace

extern "C" int32_t gvt_nccl_unique_id(uint8_t out128[128]) {
    if (!out128) return fail(GVT_ERR_INVALID, "null out");
    if (!g_nccl.load()) return fail(GVT_ERR_NCCL, "libnccl.so.2 not loadable: %s", dlerror());
    NcclId id;
    int rc = g_nccl.GetUniqueId(&id);
    if (rc != 0) return fail(GVT_ERR_NCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString(rc));
    memcpy(out128, id.internal, 128);
    return GVT_OK;
}

// --------------------------------------------------------------------------------------------------
// Seam A: PhysicsEngine
// --------------------------------------------------------------------------------------------------
struct gvt_engine {
    double mass, spin;
    std::vector<float> sab;     // engine-owned 2048 f32 (lib.rs:67)
    float* ext_sab = nullptr;   // lib.rs:74-76
    std::vector<float> lut_buffer;   // lib.rs:48,66: empty until generate_disk_lut() fills it
    host::CameraState camera, last_good;
    // device scratch for integrate_rays
    cudaStream_t stream = nullptr;
    double* d_in = nullptr; double* d_out = nullptr; double* d_drift = nullptr;
    uint32_t* d_term = nullptr; uint32_t* d_steps = nullptr; uint32_t* d_rhs = nullptr;
    uint64_t cap = 0;
    bool device_ready = false;
};

extern "C" int32_t gvt_engine_create(double mass, double spin, gvt_engine** out) {
    if (!out) return fail(GVT_ERR_INVALID, "null out");
    gvt_engine* e = new (std::nothrow) gvt_engine();
    if (!e) return fail(GVT_ERR_INVALID, "out of memory");
    e->mass = mass; e->spin = spin;
    e->sab.assign(GVT_SAB_INTERNAL_F32, 0.0f);
    *out = e;
    return GVT_OK;
}
static void engine_free_device(gvt_engine* e) {
    if (e->d_in) cudaFree(e->d_in);
    if (e->d_out) cudaFree(e->d_out);
    if (e->d_drift) cudaFree(e->d_drift);
    if (e->d_term) cudaFree(e->d_term);
    if (e->d_steps) cudaFree(e->d_steps);
    if (e->d_rhs) cudaFree(e->d_rhs);
    e->d_in = e->d_out = e->d_drift = nullptr; e->d_term = e->d_steps = e->d_rhs = nullptr; e->cap = 0;
}
extern "C" int32_t gvt_engine_destroy(gvt_engine* e) {
    if (!e) return GVT_OK;
    engine_free_device(e);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return GVT_OK;
}
extern "C" int32_t gvt_engine_update_params(gvt_engine* e, double mass, double spin) {
    if (!e) return fail(GVT_ERR_INVALID, "null engine");
    e->mass = mass; e->spin = spin;
    return GVT_OK;
}
extern "C" int32_t gvt_engine_attach_sab(gvt_engine* e, float* sab) {
    if (!e) return fail(GVT_ERR_INVALID, "null engine");
    e->ext_sab = sab;
    return GVT_OK;
}
extern "C" int32_t gvt_engine_get_sab_ptr(gvt_engine* e, const float** out) {
    if (!e || !out) return fail(GVT_ERR_INVALID, "null argument");
    *out = e->sab.data();
    return GVT_OK;
}
extern "C" int32_t gvt_engine_get_sab_layout(gvt_engine* e, uint32_t out5[5]) {
    if (!e || !out5) return fail(GVT_ERR_INVALID, "null argument");
    out5[0] = GVT_SAB_OFFSET_CONTROL; out5[1] = GVT_SAB_OFFSET_CAMERA; out5[2] = GVT_SAB_OFFSET_PHYSICS;
    out5[3] = GVT_SAB_OFFSET_TELEMETRY; out5[4] = GVT_SAB_OFFSET_LUTS;
    return GVT_OK;
}
extern "C" int32_t gvt_engine_set_camera_state(gvt_engine* e, double px, double py, double pz, double, double, double) {
    if (!e) return fail(GVT_ERR_INVALID, "null engine");
    e->camera.position = {px, py, pz};  // lib.rs:120-122 sets the position only
    return GVT_OK;
}
extern "C" int32_t gvt_engine_set_auto_spin(gvt_engine* e, int32_t enabled) {
    if (!e) return fail(GVT_ERR_INVALID, "null engine");
    e->camera.auto_spin = enabled != 0;
    return GVT_OK;
}
#define ENGINE_SCALAR(NAME, EXPR)                                              \
    extern "C" int32_t NAME(gvt_engine* e, double* out) {                      \
        if (!e || !out) return fail(GVT_ERR_INVALID, "null argument");         \
        host::Hole bh(e->mass, e->spin);                                       \
        *out = (EXPR);                                                         \
        return GVT_OK;                                                         \
    }
ENGINE_SCALAR(gvt_engine_compute_horizon, bh.horizon())
ENGINE_SCALAR(gvt_engine_compute_isco, bh.isco(true))
ENGINE_SCALAR(gvt_engine_compute_photon_sphere, bh.photon_sphere())
extern "C" int32_t gvt_engine_compute_dilation(gvt_engine* e, double r, double* out) {
    if (!e || !out) return fail(GVT_ERR_INVALID, "null argument");
    *out = host::Hole(e->mass, e->spin).dilation(r);
    return GVT_OK;
}
extern "C" int32_t gvt_engine_compute_g_factor(gvt_engine* e, double r, double lambda, double* out) {
    if (!e || !out) return fail(GVT_ERR_INVALID, "null argument");
    *out = host::g_factor(r, e->mass, e->spin, lambda);  // lib.rs:203-205 passes the raw (unclamped) spin
    return GVT_OK;
}
extern "C" int32_t gvt_engine_compute_shadow_curve(gvt_engine* e, double theta_obs, uint32_t n_points, float* out_pairs,
                                                   uint32_t capacity_pairs, uint32_t* n_pairs) {
    if (!e || !n_pairs) return fail(GVT_ERR_INVALID, "null argument");
    const auto curve = host::bardeen_shadow(host::Hole(e->mass, e->spin), theta_obs, n_points);
    *n_pairs = (uint32_t)curve.size();
    if (out_pairs)
        for (size_t i = 0; i < curve.size() && i < capacity_pairs; i++) {
            out_pairs[2 * i] = (float)curve[i].first; out_pairs[2 * i + 1] = (float)curve[i].second;
        }
    return GVT_OK;
}
extern "C" int32_t gvt_engine_compute_shadow_radius(gvt_engine* e, double* out) {
    if (!e || !out) return fail(GVT_ERR_INVALID, "null argument");
    *out = 3.0 * std::sqrt(3.0) * e->mass;   // shadow.rs:191-193
    return GVT_OK;
}
extern "C" int32_t gvt_engine_compute_shadow_shift(gvt_engine* e, double theta_obs, float out2[2]) {
    if (!e || !out2) return fail(GVT_ERR_INVALID, "null argument");
    const auto curve = host::bardeen_shadow(host::Hole(e->mass, e->spin), theta_obs, 32);
    double min_a = 0.0, max_a = 0.0;
    if (!curve.empty()) {
        min_a = max_a = curve[0].first;
        for (const auto& c : curve) { if (c.first < min_a) min_a = c.first; if (c.first > max_a) max_a = c.first; }
    }
    out2[0] = (float)min_a; out2[1] = (float)max_a;
    return GVT_OK;
}
extern "C" int32_t gvt_engine_compute_disk_flux(gvt_engine* e, double r, double* out) {
    if (!e || !out) return fail(GVT_ERR_INVALID, "null argument");
    *out = host::nt::flux(r, host::Hole(e->mass, e->spin), 1.0);   // page_thorne_flux(r, metric_bl, 1.0)
    return GVT_OK;
}
// ---- spacetime-visualisation helpers (lib.rs:139-305) ----
extern "C" int32_t gvt_engine_compute_kretschner(gvt_engine* e, double r, double theta, double* out) {
    if (!e || !out) return fail(GVT_ERR_INVALID, "null argument");
    *out = host::viz::kretschner(r, theta, e->mass, e->spin);
    return GVT_OK;
}
extern "C" int32_t gvt_engine_compute_light_cone_tilt(gvt_engine* e, double r, double theta, double* out) {
    if (!e || !out) return fail(GVT_ERR_INVALID, "null argument");
    *out = host::viz::light_cone_tilt(host::Hole(e->mass, e->spin), r, theta);
    return GVT_OK;
}
extern "C" int32_t gvt_engine_compute_frame_drag_omega(gvt_engine* e, double r, double theta, double* out) {
    if (!e || !out) return fail(GVT_ERR_INVALID, "null argument");
    *out = host::viz::frame_drag_omega(host::Hole(e->mass, e->spin), r, theta);
    return GVT_OK;
}
static int32_t check_field(gvt_engine* e, uint32_t n1, uint32_t n2, const float* out) {
    if (!e || !out) return fail(GVT_ERR_INVALID, "null argument");
    if (n1 < 2 || n2 < 2) return fail(GVT_ERR_INVALID, "sampling lattices need at least 2 points per axis (the reference divides by n - 1)");
    return GVT_OK;
}
extern "C" int32_t gvt_engine_generate_curvature_field(gvt_engine* e, double r_min, double r_max, uint32_t n_radial,
                                                       uint32_t n_polar, float* out3) {
    const int32_t rc = check_field(e, n_radial, n_polar, out3);
    if (rc != GVT_OK) return rc;
    const double m = e->mass, s = e->spin;
    host::viz::field(r_min, r_max, n_radial, n_polar, out3, [&](double r, double th) { return host::viz::kretschner(r, th, m, s); });
    return GVT_OK;
}
extern "C" int32_t gvt_engine_generate_tilt_field(gvt_engine* e, double r_min, double r_max, uint32_t n_radial, uint32_t n_polar,
                                                  float* out3) {
    const int32_t rc = check_field(e, n_radial, n_polar, out3);
    if (rc != GVT_OK) return rc;
    const host::Hole bh(e->mass, e->spin);
    host::viz::field(r_min, r_max, n_radial, n_polar, out3, [&](double r, double th) { return host::viz::light_cone_tilt(bh, r, th); });
    return GVT_OK;
}
extern "C" int32_t gvt_engine_generate_frame_drag_field(gvt_engine* e, double r_min, double r_max, uint32_t n_radial,
                                                        uint32_t n_polar, float* out3) {
    const int32_t rc = check_field(e, n_radial, n_polar, out3);
    if (rc != GVT_OK) return rc;
    const host::Hole bh(e->mass, e->spin);
    host::viz::field(r_min, r_max, n_radial, n_polar, out3, [&](double r, double th) { return host::viz::frame_drag_omega(bh, r, th); });
    return GVT_OK;
}
extern "C" int32_t gvt_engine_compute_flamm_height(gvt_engine* e, double r, double* out) {
    if (!e || !out) return fail(GVT_ERR_INVALID, "null argument");
    *out = host::viz::flamm_height(r, e->mass);
    return GVT_OK;
}
extern "C" int32_t gvt_engine_compute_proper_distance(gvt_engine* e, double r1, double r2, uint32_t n_steps, double* out) {
    if (!e || !out) return fail(GVT_ERR_INVALID, "null argument");
    *out = host::viz::proper_distance(host::Hole(e->mass, e->spin), r1, r2, n_steps);
    return GVT_OK;
}
extern "C" int32_t gvt_engine_generate_embedding_mesh(gvt_engine* e, double r_min, double r_max, uint32_t n_radial,
                                                      uint32_t n_angular, float* out3) {
    if (!e || !out3) return fail(GVT_ERR_INVALID, "null argument");
    if (n_radial < 2 || n_angular < 1) return fail(GVT_ERR_INVALID, "embedding mesh needs n_radial >= 2 and n_angular >= 1");
    host::viz::embedding_mesh(e->mass, e->spin, r_min, r_max, n_radial, n_angular, out3);
    return GVT_OK;
}
extern "C" int32_t gvt_engine_generate_ergosphere_mesh(gvt_engine* e, uint32_t n_polar, uint32_t n_azimuthal, float* out3) {
    if (!e || !out3) return fail(GVT_ERR_INVALID, "null argument");
    if (n_polar < 2 || n_azimuthal < 1) return fail(GVT_ERR_INVALID, "ergosphere mesh needs n_polar >= 2 and n_azimuthal >= 1");
    host::viz::ergosphere_mesh(host::Hole(e->mass, e->spin), n_polar, n_azimuthal, out3);
    return GVT_OK;
}
extern "C" int32_t gvt_engine_generate_disk_lut(gvt_engine* e, float* out512) {
    if (!e || !out512) return fail(GVT_ERR_INVALID, "null argument");
    host::disk_lut(host::Hole(e->mass, e->spin), 512, out512);  // lut_width 512 (lib.rs:65)
    e->lut_buffer.assign(out512, out512 + 512);                 // lib.rs:108: the engine keeps its own copy
    return GVT_OK;
}
extern "C" int32_t gvt_engine_get_disk_lut_ptr(gvt_engine* e, const float** out, uint32_t* n) {
    if (!e || !out) return fail(GVT_ERR_INVALID, "null argument");
    *out = e->lut_buffer.empty() ? nullptr : e->lut_buffer.data();   // lib.rs:112-114 (dangling-free: null while empty)
    if (n) *n = (uint32_t)e->lut_buffer.size();
    return GVT_OK;
}
extern "C" int32_t gvt_engine_generate_spectrum_lut(gvt_engine* e, uint32_t w, uint32_t h, double max_temp, float* out) {
    if (!e || !out || w == 0 || h == 0) return fail(GVT_ERR_INVALID, "bad argument");
    host::spectrum_lut(w, h, max_temp, out);
    return GVT_OK;
}

// lib.rs:308-409
extern "C" int32_t gvt_engine_tick_sab(gvt_engine* e, double dt_override) {
    if (!e) return fail(GVT_ERR_INVALID, "null engine");
    float* sab = e->ext_sab ? e->ext_sab : e->sab.data();
    const double mouse_dx = (double)sab[GVT_SAB_OFFSET_CONTROL + 1];
    const double mouse_dy = (double)sab[GVT_SAB_OFFSET_CONTROL + 2];
    const double zoom_delta = (double)sab[GVT_SAB_OFFSET_CONTROL + 3];
    const double dt = dt_override > 0.0 ? dt_override : (double)sab[GVT_SAB_OFFSET_CONTROL + 4];
    sab[GVT_SAB_OFFSET_CONTROL + 1] = 0.0f;
    sab[GVT_SAB_OFFSET_CONTROL + 2] = 0.0f;
    sab[GVT_SAB_OFFSET_CONTROL + 3] = 0.0f;

    host::update_camera(mouse_dx, mouse_dy, zoom_delta, dt, e->camera);
    if (!e->camera.valid()) e->camera = e->last_good; else e->last_good = e->camera;

    float* cam = sab + GVT_SAB_OFFSET_CAMERA;
    cam[0] = (float)e->camera.position.x; cam[1] = (float)e->camera.position.y; cam[2] = (float)e->camera.position.z;
    cam[4] = (float)e->camera.velocity.x; cam[5] = (float)e->camera.velocity.y; cam[6] = (float)e->camera.velocity.z;
    cam[8] = (float)e->camera.quat[0]; cam[9] = (float)e->camera.quat[1];
    cam[10] = (float)e->camera.quat[2]; cam[11] = (float)e->camera.quat[3];

    host::Hole bh(e->mass, e->spin);
    float* phys = sab + GVT_SAB_OFFSET_PHYSICS;
    phys[0] = (float)bh.horizon();
    phys[1] = (float)bh.isco(true);
    phys[2] = (float)e->mass;
    phys[3] = (float)e->spin;
    const host::Vec3 p = e->camera.position;
    const double r_cam = std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z);
    if (r_cam > 0.0) {
        const double theta_obs = std::acos(p.y / r_cam);
        const auto curve = host::bardeen_shadow(bh, theta_obs, 32);
        for (int i = 0; i < 128; i++) phys[16 + i] = 0.0f;
        const size_t np = std::min<size_t>(curve.size(), 64);
        phys[15] = (float)np;
        for (size_t i = 0; i < np; i++) { phys[16 + 2 * i] = (float)curve[i].first; phys[16 + 2 * i + 1] = (float)curve[i].second; }
        double min_a = 0.0, max_a = 0.0;
        if (!curve.empty()) {
            min_a = max_a = curve[0].first;
            for (const auto& c : curve) { if (c.first < min_a) min_a = c.first; if (c.first > max_a) max_a = c.first; }
        }
        phys[4] = (float)min_a; phys[5] = (float)max_a;
    }
    sab[GVT_SAB_OFFSET_TELEMETRY] += 1.0f;  // lib.rs:407: an f32 increment on the engine's buffer
    return GVT_OK;
}

static int32_t engine_ensure_device(gvt_engine* e, uint64_t n) {
    if (!e->device_ready) {
        int nd = 0;
        if (cudaGetDeviceCount(&nd) != cudaSuccess || nd == 0)
            return fail(GVT_ERR_NO_DEVICE, "no CUDA device: geodesic integration has no CPU path");
        CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
        e->device_ready = true;
    }
    if (n > e->cap) {
        engine_free_device(e);
        CK(cudaMalloc(&e->d_in, n * 8 * sizeof(double)));
        CK(cudaMalloc(&e->d_out, n * 8 * sizeof(double)));
        CK(cudaMalloc(&e->d_drift, n * sizeof(double)));
        CK(cudaMalloc(&e->d_term, n * sizeof(uint32_t)));
        CK(cudaMalloc(&e->d_steps, n * sizeof(uint32_t)));
        CK(cudaMalloc(&e->d_rhs, n * sizeof(uint32_t)));
        e->cap = n;
    }
    return GVT_OK;
}

extern "C" int32_t gvt_engine_integrate_rays(gvt_engine* e, const GvtRenderParams* o, uint64_t n, const double* in_xp,
                                             double* out_xp, uint32_t* term, uint32_t* steps_taken, double* max_drift,
                                             uint32_t* rhs_evals) {
    if (!e || !o || !in_xp || !out_xp) return fail(GVT_ERR_INVALID, "null argument");
    if (n == 0) return GVT_OK;
    if (o->renormalize_interval == 0) return fail(GVT_ERR_INVALID, "renormalize_interval must be > 0");
    int32_t rc = engine_ensure_device(e, n);
    if (rc != GVT_OK) return rc;
    host::Hole bh(e->mass, e->spin);
    RayBatchParams p{};
    p.M = bh.mass; p.a = bh.a(); p.rh = bh.horizon(); p.r_term = p.rh * 1.001; p.escape_r = o->escape_radius;
    p.tol = o->tolerance; p.h0 = o->initial_step;
    p.max_steps = o->max_steps; p.renorm_interval = o->renormalize_interval; p.step_rule = o->step_rule;
    p.method = o->method; p.coords = o->coords; p.n = n;
    p.in_xp = e->d_in; p.out_xp = e->d_out; p.term = e->d_term; p.steps = e->d_steps; p.drift = e->d_drift; p.rhs = e->d_rhs;
    CK(cudaMemcpyAsync(e->d_in, in_xp, n * 8 * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CK(launch_integrate_rays(p, e->stream));
    CK(cudaMemcpyAsync(out_xp, e->d_out, n * 8 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    if (term) CK(cudaMemcpyAsync(term, e->d_term, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    if (steps_taken) CK(cudaMemcpyAsync(steps_taken, e->d_steps, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    if (max_drift) CK(cudaMemcpyAsync(max_drift, e->d_drift, n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    if (rhs_evals) CK(cudaMemcpyAsync(rhs_evals, e->d_rhs, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return GVT_OK;
}

// lib.rs:422-464
extern "C" int32_t gvt_engine_integrate_ray(gvt_engine* e, const double in8[8], uint64_t steps, double tolerance,
                                            int32_t use_ks, double out8[8], uint32_t* term, uint64_t* steps_taken,
                                            double* max_drift) {
    if (!e || !in8 || !out8) return fail(GVT_ERR_INVALID, "null argument");
    GvtRenderParams o{};
    o.struct_size = sizeof(o);
    o.method = GVT_METHOD_RKF45; o.precision = GVT_PRECISION_F64;
    o.coords = use_ks ? GVT_COORDS_KERR_SCHILD : GVT_COORDS_BOYER_LINDQUIST;
    o.step_rule = GVT_STEP_CONSTANT;
    o.max_steps = steps > 0xffffffffull ? 0xffffffffu : (uint32_t)steps;
    o.renormalize_interval = 10; o.tolerance = tolerance; o.initial_step = 0.01; o.escape_radius = 1000.0;
    uint32_t st = 0;
    int32_t rc = gvt_engine_integrate_rays(e, &o, 1, in8, out8, term, &st, max_drift, nullptr);
    if (steps_taken) *steps_taken = st;
    return rc;
}

// --------------------------------------------------------------------------------------------------
// Seam B: renderer
// --------------------------------------------------------------------------------------------------
struct gvt_renderer {
    int device = 0, rank = 0, world = 1, sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint32_t width = 0, height = 0, padded_height = 0, rows_per_rank = 0;
    float4* cur = nullptr;      // trace output
    float4* frame = nullptr;    // finished frame of the last call (TAA-resolved and all-gathered)
    float4* hist = nullptr;     // previous finished frame (TAA history)
    float4* buf[2] = {nullptr, nullptr};            // the two allocations frame/hist alternate between
    float4* peer[GVT_MAX_PEERS][2] = {};            // peer ranks' buf[0], buf[1] (CUDA IPC), GVT_FLAG_PEER_STORE
    bool peer_open[GVT_MAX_PEERS] = {};
    void* half_frame = nullptr; // RGBA16F staging
    bool history_valid = false;
    float4* last_peer_target = nullptr;   // the buffer the peers filled on the previous GVT_FLAG_PEER_STORE frame
    uint32_t rows_override[2] = {0u, 0u};  // gvt_render_rows
    bool frame_f16 = false;               // cur / frame / hist hold RGBA16F (gvt_render_set_frame_format)
    size_t px_bytes() const { return frame_f16 ? 8 : 16; }
    uint32_t storage_format() const { return frame_f16 ? GVT_FORMAT_RGBA16F : GVT_FORMAT_RGBA32F; }
    // byte address of pixel `px` in one of the frame-chain buffers
    char* at(float4* base, size_t px) const { return reinterpret_cast<char*>(base) + px * px_bytes(); }
    FrameBlock* d_block = nullptr; FrameBlock* h_block = nullptr;   // device / pinned host
    Counters* d_counters = nullptr; Counters* h_counters = nullptr;
    float4* d_spectrum = nullptr; uint32_t spec_w = 0, spec_h = 0;
    std::vector<float> tdisk; double tdisk_rin = 0.0, tdisk_rout = 0.0;
    double lut_mass = 0.0, lut_spin = 0.0; bool luts_ready = false;
    ncclComm_t_ comm = nullptr;
    float* d_sink = nullptr;
    // debug buffers for gvt_trace_states
    double* d_xp = nullptr; double* d_drift = nullptr; double* d_rgba = nullptr; uint32_t* d_term = nullptr; uint32_t* d_steps = nullptr;
    size_t dbg_cap = 0;
    // WebGL2 fragment-shader path: red channels of the two noise textures, per-pixel parity hooks
    uint8_t* d_noise_r = nullptr; uint8_t* d_blue_r = nullptr;
    uint32_t* d_gsteps = nullptr; uint32_t* d_ghit = nullptr; size_t g_cap = 0; bool g_valid = false;
    // bloom: display-referred output + RGBA16F scratch (half-res bright texture, two quarter-res blur textures)
    float4* display = nullptr; uint2* bloom_half = nullptr; uint2* bloom_q1 = nullptr; uint2* bloom_q2 = nullptr;
    uint32_t bloom_w = 0, bloom_h = 0;
};

extern "C" int32_t gvt_render_params_default(GvtRenderParams* p) {
    if (!p) return fail(GVT_ERR_INVALID, "null params");
    memset(p, 0, sizeof(*p));
    p->struct_size = sizeof(*p);
    p->method = GVT_METHOD_SYMPLECTIC; p->precision = GVT_PRECISION_F64; p->coords = GVT_COORDS_KERR_SCHILD;
    p->step_rule = GVT_STEP_WGSL; p->max_steps = 512; p->renormalize_interval = 10; p->flags = 0;
    p->output_format = GVT_FORMAT_RGBA32F;
    p->tolerance = 1e-8; p->initial_step = 0.01; p->escape_radius = 1000.0; p->disk_r_out = 50.0;
    p->taa_blend = 0.75f; p->taa_camera_moving = 0;   // webgl/renderer.ts:380-385
    return GVT_OK;
}

extern "C" int32_t gvt_render_create(const GvtDeviceConfig* cfg, gvt_renderer** out) {
    if (!cfg || !out) return fail(GVT_ERR_INVALID, "null argument");
    int nd = 0;
    if (cudaGetDeviceCount(&nd) != cudaSuccess || nd == 0)
        return fail(GVT_ERR_NO_DEVICE, "no CUDA device: the renderer has no CPU path");
    if (cfg->device < 0 || cfg->device >= nd) return fail(GVT_ERR_INVALID, "device %d out of range (%d devices)", cfg->device, nd);
    if (cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size) return fail(GVT_ERR_INVALID, "bad rank/world_size");
    CK(cudaSetDevice(cfg->device));
    gvt_renderer* r = new (std::nothrow) gvt_renderer();
    if (!r) return fail(GVT_ERR_INVALID, "out of memory");
    struct Guard {   // a failing step below must not leak the half-built renderer
        gvt_renderer* r;
        ~Guard() { if (r) { const std::string keep = g_err; gvt_render_destroy(r); g_err = keep; } }
    } guard{r};
    r->device = cfg->device; r->rank = cfg->rank; r->world = cfg->world_size;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    r->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
        return fail(GVT_ERR_UNSUPPORTED, "device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
    }
    CK(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    for (auto& e : r->ev) CK(cudaEventCreate(&e));
    CK(cudaMalloc(&r->d_block, sizeof(FrameBlock)));
    CK(cudaMallocHost(&r->h_block, sizeof(FrameBlock)));
    CK(cudaMalloc(&r->d_counters, sizeof(Counters)));
    CK(cudaMallocHost(&r->h_counters, sizeof(Counters)));
    CK(cudaMalloc(&r->d_sink, 256));
    if (r->world > 1) {
        if (!g_nccl.load()) { return fail(GVT_ERR_NCCL, "libnccl.so.2 not loadable: %s", dlerror()); }
        NcclId id;
        memcpy(id.internal, cfg->nccl_id, 128);
        int rc = g_nccl.CommInitRank(&r->comm, r->world, id, r->rank);
        if (rc != 0) return fail(GVT_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(rc));
    }
    guard.r = nullptr;
    *out = r;
    return GVT_OK;
}

static void close_peers(gvt_renderer* r) {
    for (int p = 0; p < GVT_MAX_PEERS; p++) {
        if (r->peer_open[p]) { cudaIpcCloseMemHandle(r->peer[p][0]); cudaIpcCloseMemHandle(r->peer[p][1]); }
        r->peer_open[p] = false; r->peer[p][0] = r->peer[p][1] = nullptr;
    }
}
static void free_frames(gvt_renderer* r) {
    close_peers(r);
    if (r->cur) cudaFree(r->cur);
    if (r->frame) cudaFree(r->frame);
    if (r->hist) cudaFree(r->hist);
    if (r->half_frame) cudaFree(r->half_frame);
    r->cur = r->frame = r->hist = nullptr; r->half_frame = nullptr;
}

extern "C" int32_t gvt_render_destroy(gvt_renderer* r) {
    if (!r) return GVT_OK;
    cudaSetDevice(r->device);
    if (r->stream) cudaStreamSynchronize(r->stream);
    if (r->comm) g_nccl.CommDestroy(r->comm);
    free_frames(r);
    if (r->d_block) cudaFree(r->d_block);
    if (r->h_block) cudaFreeHost(r->h_block);
    if (r->d_counters) cudaFree(r->d_counters);
    if (r->h_counters) cudaFreeHost(r->h_counters);
    if (r->d_spectrum) cudaFree(r->d_spectrum);
    if (r->d_sink) cudaFree(r->d_sink);
    if (r->d_xp) cudaFree(r->d_xp);
    if (r->d_drift) cudaFree(r->d_drift);
    if (r->d_rgba) cudaFree(r->d_rgba);
    if (r->d_term) cudaFree(r->d_term);
    if (r->d_steps) cudaFree(r->d_steps);
    if (r->d_noise_r) cudaFree(r->d_noise_r);
    if (r->d_blue_r) cudaFree(r->d_blue_r);
    if (r->d_gsteps) cudaFree(r->d_gsteps);
    if (r->d_ghit) cudaFree(r->d_ghit);
    if (r->display) cudaFree(r->display);
    if (r->bloom_half) cudaFree(r->bloom_half);
    if (r->bloom_q1) cudaFree(r->bloom_q1);
    if (r->bloom_q2) cudaFree(r->bloom_q2);
    for (auto& e : r->ev) if (e) cudaEventDestroy(e);
    if (r->stream) cudaStreamDestroy(r->stream);
    delete r;
    return GVT_OK;
}

extern "C" int32_t gvt_render_set_luts(gvt_renderer* r, const float* spectrum, uint32_t w, uint32_t h, const float* tdisk,
                                       uint32_t n, double rin, double rout) {
    if (!r || !spectrum || !tdisk || w == 0 || h == 0) return fail(GVT_ERR_INVALID, "bad argument");
    if (n != 512) return fail(GVT_ERR_INVALID, "disk LUT must have 512 entries (lib.rs:65), got %u", n);
    CK(cudaSetDevice(r->device));
    if (r->d_spectrum) { CK(cudaFree(r->d_spectrum)); r->d_spectrum = nullptr; }
    CK(cudaMalloc(&r->d_spectrum, (size_t)w * h * sizeof(float4)));
    CK(cudaMemcpy(r->d_spectrum, spectrum, (size_t)w * h * sizeof(float4), cudaMemcpyHostToDevice));
    r->spec_w = w; r->spec_h = h;
    r->tdisk.assign(tdisk, tdisk + n);
    r->tdisk_rin = rin; r->tdisk_rout = rout;
    r->luts_ready = true;
    return GVT_OK;
}

extern "C" int32_t gvt_render_init_luts(gvt_renderer* r, double mass, double spin, uint32_t w, uint32_t h, double max_temp) {
    if (!r || w == 0 || h == 0) return fail(GVT_ERR_INVALID, "bad argument");
    std::vector<float> spec((size_t)w * h * 4), td(512);
    host::spectrum_lut(w, h, max_temp, spec.data());
    host::Hole bh(mass, spin);
    host::disk_lut(bh, 512, td.data());
    int32_t rc = gvt_render_set_luts(r, spec.data(), w, h, td.data(), 512, bh.isco(true), 50.0 * bh.mass);
    if (rc == GVT_OK) { r->lut_mass = mass; r->lut_spin = spin; }
    return rc;
}

extern "C" int32_t gvt_render_resize(gvt_renderer* r, uint32_t width, uint32_t height) {
    if (!r || width == 0 || height == 0) return fail(GVT_ERR_INVALID, "bad argument");
    if (r->width == width && r->height == height && r->cur) return GVT_OK;
    CK(cudaSetDevice(r->device));
    CK(cudaStreamSynchronize(r->stream));
    free_frames(r);
    r->width = width; r->height = height;
    r->rows_per_rank = (height + r->world - 1) / r->world;
    r->padded_height = r->rows_per_rank * r->world;  // all-gather needs equal blocks
    const size_t bytes = (size_t)width * r->padded_height * r->px_bytes();
    CK(cudaMalloc(&r->cur, bytes));
    CK(cudaMalloc(&r->frame, bytes));
    CK(cudaMalloc(&r->hist, bytes));
    r->buf[0] = r->frame; r->buf[1] = r->hist;
    r->last_peer_target = nullptr;
    CK(cudaMemsetAsync(r->cur, 0, bytes, r->stream));
    CK(cudaMemsetAsync(r->frame, 0, bytes, r->stream));
    CK(cudaMemsetAsync(r->hist, 0, bytes, r->stream));  // WebGPU textures start zeroed: so does the history
    r->history_valid = false;
    return GVT_OK;
}

// The format of the frame chain (cur / frame / hist): RGBA32F (default, the parity format) or RGBA16F, the reference's own
// texture format (rendering/reprojection.ts:120-140, webgpu/renderer.ts:161-180). Re-allocates the buffers, clears the history.
extern "C" int32_t gvt_render_set_frame_format(gvt_renderer* r, uint32_t format) {
    if (!r || (format != GVT_FORMAT_RGBA32F && format != GVT_FORMAT_RGBA16F)) return fail(GVT_ERR_INVALID, "frame format: RGBA32F or RGBA16F");
    const bool f16 = format == GVT_FORMAT_RGBA16F;
    if (f16 == r->frame_f16) return GVT_OK;
    r->frame_f16 = f16;
    if (!r->cur) return GVT_OK;
    const uint32_t w = r->width, h = r->height;
    r->width = r->height = 0;                     // force gvt_render_resize to re-allocate
    return gvt_render_resize(r, w, h);
}

extern "C" int32_t gvt_render_get_size(gvt_renderer* r, uint32_t* width, uint32_t* height) {
    if (!r || !width || !height) return fail(GVT_ERR_INVALID, "null argument");
    *width = r->width; *height = r->height;
    return GVT_OK;
}

extern "C" int32_t gvt_render_reset_history(gvt_renderer* r) {
    if (!r) return fail(GVT_ERR_INVALID, "null renderer");
    if (r->hist) CK(cudaMemsetAsync(r->hist, 0, (size_t)r->width * r->padded_height * r->px_bytes(), r->stream));
    r->history_valid = false;
    return GVT_OK;
}

// compute.wgsl.ts:135-145
static double halton(uint32_t index, uint32_t base) {
    double result = 0.0, f = 1.0 / (double)base;
    for (uint32_t i = index; i > 0u; i /= base) { result += f * (double)(i % base); f = f / (double)base; }
    return result;
}

// Fill the TMA-staged block and the kernel parameters from the reference-layout uniforms.
static int32_t build_frame(gvt_renderer* r, const GvtCamera* cam, const GvtPhysicsParams* phys, const GvtRenderParams* rp,
                           FrameParams& P) {
    if (rp->struct_size < offsetof(GvtRenderParams, disk_r_out) + sizeof(double))
        return fail(GVT_ERR_INVALID, "GvtRenderParams.struct_size = %u: fill the struct with gvt_render_params_default first", rp->struct_size);
    if (!r->luts_ready) return fail(GVT_ERR_INVALID, "LUTs not initialised: call gvt_render_init_luts / gvt_render_set_luts first");
    if (rp->coords != GVT_COORDS_KERR_SCHILD)
        return fail(GVT_ERR_UNSUPPORTED, "the render path traces in Kerr-Schild coordinates (lib.rs:64,454); use gvt_engine_integrate_rays for Boyer-Lindquist");
    if (rp->method > GVT_METHOD_VERLET_GLSL || (rp->precision > GVT_PRECISION_F32 && rp->precision != GVT_PRECISION_MIXED) ||
        rp->output_format > GVT_FORMAT_RGBA8_UNORM)
        return fail(GVT_ERR_INVALID, "bad method/precision/output_format");
    if (rp->precision == GVT_PRECISION_MIXED && rp->method != GVT_METHOD_SYMPLECTIC)
        return fail(GVT_ERR_UNSUPPORTED, "GVT_PRECISION_MIXED is defined for GVT_METHOD_SYMPLECTIC only (f32 predictors of the implicit midpoint)");
    if (rp->renormalize_interval == 0) return fail(GVT_ERR_INVALID, "renormalize_interval must be > 0");
    const uint32_t W = (uint32_t)phys->resolution[0], H = (uint32_t)phys->resolution[1];
    if (W == 0 || H == 0) return fail(GVT_ERR_INVALID, "zero resolution");
    host::Hole bh((double)phys->mass, (double)phys->spin);
    FrameBlock* b = r->h_block;
    for (int i = 0; i < 16; i++) { b->inv_proj[i] = (double)cam->inv_proj[i]; b->inv_view[i] = (double)cam->inv_view[i]; }
    const double cx = (double)cam->position[0], cy = (double)cam->position[1], cz = (double)cam->position[2];
    b->r0 = std::sqrt(cx * cx + cy * cy + cz * cz);
    if (!(b->r0 > 0.0)) return fail(GVT_ERR_INVALID, "camera at the origin");
    b->theta0 = std::acos(std::min(1.0, std::max(-1.0, cy / b->r0)));
    b->phi0 = std::atan2(cz, cx);
    b->st = std::sin(b->theta0); b->ct = std::cos(b->theta0); b->sp = std::sin(b->phi0); b->cp = std::cos(b->phi0);
    b->safe_st = std::max(b->st, 1e-4);
    b->inv_width = 1.0 / (double)W; b->inv_height = 1.0 / (double)H;
    b->jx = b->jy = 0.0;
    if (rp->flags & GVT_FLAG_JITTER) {
        b->jx = (halton((phys->frame_index % 8u) + 1u, 2u) - 0.5) / (double)W;
        b->jy = (halton((phys->frame_index % 8u) + 1u, 3u) - 0.5) / (double)H;
    }
    b->cam_pos[0] = cx; b->cam_pos[1] = cy; b->cam_pos[2] = cz;
    {   // kerr_photon_sphere, chunks/metric.ts:32-37
        const double as = std::min(0.9999, std::max(-0.9999, bh.a() / bh.mass));
        b->rph = 2.0 * bh.mass * (1.0 + std::cos((2.0 / 3.0) * std::acos(std::min(1.0, std::max(-1.0, -as)))));
    }
    memcpy(b->tdisk, r->tdisk.data(), 512 * sizeof(float));

    memset(&P, 0, sizeof(P));
    P.M = bh.mass; P.a = bh.a(); P.spin = bh.spin; P.rh = bh.horizon(); P.r_term = P.rh * 1.001;
    P.a2 = P.a * P.a; P.twoM = 2.0 * P.M;
    P.sqrtM = std::sqrt(bh.mass);
    P.escape_r = rp->escape_radius; P.r_in = bh.isco(true); P.r_out = rp->disk_r_out;
    P.tol = rp->tolerance; P.h0 = rp->initial_step;
    {   // Zone radii. A chunk is 8 steps and |dr/dlambda| <= 1 for a null ray with p_t = -1, so a ray moves at most
        // 8 h_max in r within a chunk.
        const double h_max = rp->step_rule == GVT_STEP_WGSL ? 1.0 : std::fabs(rp->initial_step);
        const double travel = 1.5 * 8.0 * h_max;   // |dr/dlambda| <= (r^2 + a^2 + a |L|) / Sigma: up to ~1.1 near the hole, 1.5 is generous
        // step rule saturated (compute.wgsl.ts:213: 0.15 (r - r+) >= 1), with one chunk of margin; the constant rule is
        // saturated everywhere, but the horizon test still needs r > 1.001 r+ for the whole chunk
        P.h_const = rp->step_rule == GVT_STEP_WGSL ? 1.0 : rp->initial_step;
        P.r_hconst = (rp->step_rule == GVT_STEP_WGSL ? P.rh + 1.0 / 0.15 : P.r_term) + travel + 1e-3;
        P.r_escape_guard = rp->escape_radius - travel - 1e-3;
        P.r_sat = rp->step_rule == GVT_STEP_WGSL ? P.rh + 1.0 / 0.15 + 1e-9 : P.r_term;
        // high-word windows: hi(r) in [hi(lo) + 1, hi(escape_r) - 1] implies lo < r < escape_r for positive finite lo, escape_r
        auto hi_of = [](double x) { uint64_t b; memcpy(&b, &x, 8); return (uint32_t)(b >> 32); };
        const double los[2] = {P.r_term, P.r_sat};
        for (int z = 0; z < 2; z++) {
            const uint32_t lo = hi_of(los[z]) + 1u, hi = hi_of(rp->escape_radius);
            P.alive_lo[z] = lo;
            P.alive_span[z] = (rp->escape_radius > los[z] && los[z] > 0.0 && hi > lo) ? hi - lo : 0u;   // 0: always take the exact tests
        }
        // GVT_PRECISION_MIXED: r_switch = 35 M
        const char* e = getenv("GVT_MIXED_RSWITCH");   // tuning experiments only
        const double r_switch = (e && atof(e) > 0.0 ? atof(e) : 35.0) * bh.mass;
        P.r_far = std::max(r_switch + travel, P.r_hconst);
        {   // f64 zone 3: rotated trigonometry (step_symplectic_rot) and no equatorial-crossing test. It opens where a chunk
            // cannot reach the disk's outer edge (compute.wgsl.ts:216-218 shades crossings inside r_out only) and, per ray,
            // where a whole step turns theta by at most 2^-8: r > travel + sqrt(256 h B), B^2 = Q + a^2 (k_trace_tile).
            P.r_rot = std::max(P.r_far, rp->disk_r_out + travel + 1e-3);
            const double rmin = P.r_rot - travel, hc = std::max(std::fabs(P.h_const), 1e-300);
            P.rot_travel = travel; P.rot_k = 256.0 * hc;
            P.f32_rot_travel = (float)travel; P.f32_rot_inv_k = (float)(1.0 / P.rot_k);
            P.rot_q_max = 1e300;
            // 32x the h^2 k = 4 estimate (3/4 h^2 / r_min^4). Measured on the headline frame: with trig_full every step a 4x margin
            // reproduced the oracle's census on all 8.3 M pixels; the chained rotation (rounding a few ulp looser) needs 16x for
            // the last marginal ray, a pixel column next to the image of the spin axis.
            P.rot_stab = 24.0 * hc * hc / (rmin * rmin * rmin * rmin);
            if (const char* m = getenv("GVT_ROT_STAB_MULT")) if (atof(m) > 0.0) P.rot_stab *= atof(m);   // tuning experiments only
            if (getenv("GVT_NO_ROT")) P.rot_q_max = -1.0;                    // diagnostics: f64 zone 3 closed
        }
        P.f32_M = (float)bh.mass; P.f32_a = (float)bh.a(); P.f32_a2 = (float)(bh.a() * bh.a()); P.f32_twoM = (float)(2.0 * bh.mass); P.f32_hconst = (float)P.h_const;
    }
    P.tdisk_rin = r->tdisk_rin; P.tdisk_scale = 511.0 / (r->tdisk_rout - r->tdisk_rin);
    P.width = W; P.height = H;
    P.max_steps = rp->max_steps; P.renorm_interval = rp->renormalize_interval; P.step_rule = rp->step_rule;
    P.tdisk_n = 512; P.spec_w = r->spec_w; P.spec_h = r->spec_h;
    P.lut_in_smem = ((size_t)r->spec_w * r->spec_h * sizeof(float4) <= kMaxSmemLutBytes) ? 1u : 0u;
    P.block = r->d_block; P.spectrum = r->d_spectrum; P.counters = r->d_counters;
    return GVT_OK;
}

static void fill_taa(const GvtCamera* cam, uint32_t W, uint32_t H, TaaParams& T) {
    memset(&T, 0, sizeof(T));
    // column-major mat4 (gl-matrix / WGSL): m[4 c + r]. clip = (cx, cy, 1, 1)  (ataa.wgsl.ts:56-58)
    const float* ip = cam->inv_proj; const float* iv = cam->inv_view; const float* pv = cam->prev_view_proj;
    double vA[4], vB[4], vC[4];
    for (int r = 0; r < 4; r++) { vA[r] = ip[r]; vB[r] = ip[4 + r]; vC[r] = (double)ip[8 + r] + (double)ip[12 + r]; }
    // G = 12 * PV[rows 0,1,3][:, :3] * IV[:3, :3]   (world direction has w = 0; reprojectDepth = 12, ataa.wgsl.ts:64)
    const int rows[3] = {0, 1, 3};
    for (int k = 0; k < 3; k++) {
        double G[3];
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int m = 0; m < 3; m++) s += (double)pv[4 * m + rows[k]] * (double)iv[4 * j + m];
            G[j] = 12.0 * s;
        }
        T.gA[k] = (float)(G[0] * vA[0] + G[1] * vA[1] + G[2] * vA[2]);
        T.gB[k] = (float)(G[0] * vB[0] + G[1] * vB[1] + G[2] * vB[2]);
        T.gC[k] = (float)(G[0] * vC[0] + G[1] * vC[1] + G[2] * vC[2]);
        T.c0[k] = (float)((double)pv[rows[k]] * cam->position[0] + (double)pv[4 + rows[k]] * cam->position[1] +
                          (double)pv[8 + rows[k]] * cam->position[2] + (double)pv[12 + rows[k]]);
    }
    for (int r = 0; r < 4; r++) { T.vA[r] = (float)vA[r]; T.vB[r] = (float)vB[r]; T.vC[r] = (float)vC[r]; }
    T.width = W; T.height = H; T.row0 = 0; T.row1 = H; T.host_out = nullptr; T.n_peer = 0;
    T.stripe = StripeMap{0u, 0u, 1u, 0u}; T.n_stripes = 0;
    T.mode = 0; T.blend = 0.75f; T.moving = 0;
    memcpy(T.m_inv_proj, cam->inv_proj, 64); memcpy(T.m_inv_view, cam->inv_view, 64); memcpy(T.m_prev_vp, cam->prev_view_proj, 64);
    memcpy(T.cam_pos, cam->position, 16);
}

static size_t format_bytes(uint32_t f) { return f == GVT_FORMAT_RGBA32F ? 16 : (f == GVT_FORMAT_RGBA16F ? 8 : 4); }
// converts pixels [px0, px0 + npx) of the finished frame into the staging buffer in `format` (!= the storage format)
static int32_t convert_frame(gvt_renderer* r, uint32_t format, size_t px0, size_t npx) {
    const size_t n_px = (size_t)r->width * r->height;
    if (!r->half_frame) CK(cudaMalloc(&r->half_frame, n_px * 16));
    char* dst = static_cast<char*>(r->half_frame) + px0 * format_bytes(format);
    const float4* src = reinterpret_cast<const float4*>(r->at(r->frame, px0));
    if (format == GVT_FORMAT_RGBA16F) CK(launch_f32_to_f16(src, dst, npx, r->stream));                    // storage is RGBA32F here
    else if (format == GVT_FORMAT_RGBA32F) CK(launch_f16_to_f32(src, reinterpret_cast<float4*>(dst), npx, r->stream));   // storage is RGBA16F
    else CK(launch_tonemap_rgba8(src, dst, npx, format == GVT_FORMAT_RGBA8_ACES ? 1 : (format == GVT_FORMAT_RGBA8_UNORM ? 2 : 0), r->stream, r->frame_f16));
    return GVT_OK;
}

extern "C" int32_t gvt_render_read_frame(gvt_renderer* r, uint32_t format, void* host_rgba) {
    if (!r || !host_rgba || !r->frame || format > GVT_FORMAT_RGBA8_UNORM) return fail(GVT_ERR_INVALID, "bad argument / no frame");
    CK(cudaSetDevice(r->device));
    const size_t n_px = (size_t)r->width * r->height;
    if (format != r->storage_format()) {
        int32_t rc = convert_frame(r, format, 0, n_px);
        if (rc != GVT_OK) return rc;
        CK(cudaMemcpyAsync(host_rgba, r->half_frame, n_px * format_bytes(format), cudaMemcpyDefault, r->stream));
    } else {
        CK(cudaMemcpyAsync(host_rgba, r->frame, n_px * r->px_bytes(), cudaMemcpyDefault, r->stream));
    }
    CK(cudaStreamSynchronize(r->stream));
    return GVT_OK;
}

// One frame through the shared pipeline: [barrier] -> producing kernel (the fused geodesic trace, or the WebGL2
// fragment shader when `glsl` is given) -> optional TAA resolve -> gather / peer-store barrier -> host delivery.
static int32_t render_impl(gvt_renderer* r, const GvtCamera* cam, const GvtPhysicsParams* phys, const GvtRenderParams* rp,
                           const GvtGlslUniforms* glsl, uint32_t glsl_precision, void* host_rgba, GvtFrameStats* stats) {
    if (!r || !rp || (!glsl && (!cam || !phys))) return fail(GVT_ERR_INVALID, "null argument");
    CK(cudaSetDevice(r->device));
    const uint32_t W = glsl ? (uint32_t)glsl->resolution[0] : (uint32_t)phys->resolution[0];
    const uint32_t H = glsl ? (uint32_t)glsl->resolution[1] : (uint32_t)phys->resolution[1];
    if (W != r->width || H != r->height || !r->cur) {  // renderer.ts:281-284
        int32_t rc = gvt_render_resize(r, W, H);
        if (rc != GVT_OK) return rc;
    }
    FrameParams P;
    GlslParams G;
    if (glsl) {
        if (!r->d_noise_r || !r->d_blue_r) return fail(GVT_ERR_INVALID, "noise textures not set: call gvt_render_set_noise_textures first");
        if ((rp->flags & GVT_FLAG_TAA) && !(rp->flags & GVT_FLAG_TAA_WEBGL))
            return fail(GVT_ERR_INVALID, "the fragment-shader path resolves with GVT_FLAG_TAA_WEBGL (reprojection.glsl.ts), it has no CameraUniforms");
        memset(&G, 0, sizeof(G));
        G.u = *glsl;
        G.width = W; G.height = H; G.ys = 1;
        G.noise_r = r->d_noise_r; G.blue_r = r->d_blue_r; G.counters = r->d_counters;
        const size_t n_px = (size_t)W * H;
        const bool counts = (rp->flags & GVT_FLAG_DEBUG_COUNTS) != 0;
        r->g_valid = counts;
        if (counts && r->g_cap < n_px) {
            if (r->d_gsteps) cudaFree(r->d_gsteps);
            if (r->d_ghit) cudaFree(r->d_ghit);
            r->d_gsteps = nullptr; r->d_ghit = nullptr; r->g_cap = 0;
            CK(cudaMalloc(&r->d_gsteps, n_px * sizeof(uint32_t)));
            CK(cudaMalloc(&r->d_ghit, n_px * sizeof(uint32_t)));
            r->g_cap = n_px;
        }
        G.dbg_steps = counts ? r->d_gsteps : nullptr; G.dbg_hit = counts ? r->d_ghit : nullptr;
    } else {
        int32_t rc = build_frame(r, cam, phys, rp, P);
        if (rc != GVT_OK) return rc;
    }
    const bool taa = (rp->flags & GVT_FLAG_TAA) != 0;
    const bool budget = !glsl && (rp->flags & GVT_FLAG_BUDGET) != 0 && (rp->method == GVT_METHOD_RK4 || rp->method == GVT_METHOD_SYMPLECTIC);
    uint32_t row0 = std::min(H, (uint32_t)r->rank * r->rows_per_rank);
    uint32_t row1 = std::min(H, row0 + r->rows_per_rank);
    if (r->rows_override[1] > r->rows_override[0]) {   // gvt_render_rows: one row block of the frame on a single renderer
        row0 = std::min(H, r->rows_override[0]); row1 = std::min(H, r->rows_override[1]);
    }
    // TAA needs a one-pixel halo of the current frame: trace one redundant row above and below the block
    const uint32_t ty0 = (taa && row0 > 0) ? row0 - 1 : row0, ty1 = (taa && row1 < H) ? row1 + 1 : row1;
    // GVT_FLAG_ROW_INTERLEAVE (peer stores only: nothing needs contiguous blocks): stripes dealt round-robin to the ranks.
    // Without TAA a stripe is one row (rows rank, rank + world, ...: the finest balance); with TAA it is 16 rows traced
    // with one halo row on each side, so the 3x3 resolve stays local for 12 % redundant rows.
    const bool interleave = (rp->flags & GVT_FLAG_ROW_INTERLEAVE) != 0 && r->world > 1;
    if (interleave && (!(rp->flags & GVT_FLAG_PEER_STORE) || (rp->flags & GVT_FLAG_NO_GATHER)))
        return fail(GVT_ERR_INVALID, "GVT_FLAG_ROW_INTERLEAVE needs the GVT_FLAG_PEER_STORE gather");
    if (interleave && (rp->flags & GVT_FLAG_TAA_PRECISE))
        return fail(GVT_ERR_UNSUPPORTED, "GVT_FLAG_TAA_PRECISE (the validation build of the resolve) runs on row blocks only");
    StripeMap sm = {0u, 0u, (uint32_t)r->world, (uint32_t)r->rank};
    uint32_t n_my = 0, n_own = row1 - row0;
    if (interleave) {
        sm.s = taa ? 16u : 1u; sm.halo = taa ? 1u : 0u;
        const uint32_t n_total = (H + sm.s - 1u) / sm.s;
        n_my = (uint32_t)r->rank < n_total ? (n_total - (uint32_t)r->rank + (uint32_t)r->world - 1u) / (uint32_t)r->world : 0u;
        n_own = 0;
        for (uint32_t t = 0; t < n_my; t++) n_own += std::min(sm.s, H - (t * (uint32_t)r->world + (uint32_t)r->rank) * sm.s);
    }
    const uint32_t lattice_rows = n_my * (sm.s + 2u * sm.halo);
    if (glsl) {
        G.y0 = interleave ? 0u : ty0; G.y1 = interleave ? H : ty1; G.ys = 1u;
        G.stripe = sm; G.n_lattice_rows = lattice_rows;
    } else if (interleave) {
        P.x0 = 0; P.xs = 1; P.y0 = 0; P.y1 = H; P.ys = 1; P.nx = W; P.ny = lattice_rows; P.stripe = sm;
    } else {
        P.x0 = 0; P.xs = 1; P.y0 = ty0; P.y1 = ty1; P.ys = 1; P.nx = W; P.ny = ty1 - ty0; P.stripe = sm;
    }
    // ---- everything that can be refused is refused HERE, before any state changes or anything is enqueued: a rank that
    //      bailed out later would leave its ping-pong parity flipped against its peers and them blocked in a collective ----
    const bool own_early = (rp->flags & GVT_FLAG_D2H_OWN_ROWS) != 0 && r->world > 1;
    if (host_rgba && own_early && interleave && rp->output_format != r->storage_format())
        return fail(GVT_ERR_UNSUPPORTED, "interleaved own-row delivery needs output_format == the frame chain's format");
    if ((rp->flags & GVT_FLAG_PEER_STORE) != 0 && r->world > 1 && !(rp->flags & GVT_FLAG_NO_GATHER)) {
        if (!r->d_sink) return fail(GVT_ERR_INVALID, "no barrier buffer");
        for (int p = 0; p < r->world; p++)
            if (p != r->rank && (p >= GVT_MAX_PEERS || !r->peer_open[p]))
                return fail(GVT_ERR_INVALID, "GVT_FLAG_PEER_STORE: frames of rank %d not imported (gvt_render_import_peer_frames)", p);
    }
    // A CUDA / NCCL failure after this point is a broken device or communicator, not a caller error; the guard still
    // drains the stream and puts the ping-pong buffers back so that this renderer's own state stays consistent.
    struct LateGuard {
        gvt_renderer* r; bool swapped = false, armed = true;
        ~LateGuard() { if (armed) { cudaStreamSynchronize(r->stream); if (swapped) std::swap(r->frame, r->hist); } }
    } late{r};
    const bool peer_req = (rp->flags & GVT_FLAG_PEER_STORE) != 0 && r->world > 1 && !(rp->flags & GVT_FLAG_NO_GATHER);
    // The last finished frame becomes the history (TAA), and under the fused gather the two frame buffers ping-pong on
    // every frame, TAA or not: peers then never write into the buffer a rank may still be reading (TAA history taps, the
    // D2H copy of the previous frame), which is what lets ONE barrier per frame close the exchange (see below).
    if ((taa && r->history_valid) || (peer_req && !taa)) { std::swap(r->frame, r->hist); late.swapped = true; }
    float4* trace_out = taa ? r->cur : r->frame;
    if (glsl) { G.frame = trace_out; G.frame_f16 = r->frame_f16 ? 1u : 0u; } else { P.frame = trace_out; P.frame_f16 = r->frame_f16 ? 1u : 0u; }
    uint32_t launches = 0;
    uint64_t h2d = 0, d2h = 0;
    // Host frame delivery. If the caller's buffer is page-locked (gvt_host_alloc / gvt_host_register) and this rank
    // produces every pixel it has to deliver (single GPU, or GVT_FLAG_D2H_OWN_ROWS), the producing kernel stores each
    // finished pixel straight into host memory over PCIe (132.7 MB spread over the whole kernel: ~2 GB/s) and no D2H
    // copy follows. Otherwise the finished frame is copied after the last kernel.
    // The same holds for a frame target in DEVICE memory -- a buffer of the presenter's graphics API imported through
    // gvt_external_import_fd (cudaImportExternalMemory), or any device allocation of the caller: the producing kernel
    // stores into it directly and the frame never crosses host memory.
    const bool own = (rp->flags & GVT_FLAG_D2H_OWN_ROWS) != 0 && r->world > 1;
    float4* host_alias = nullptr;
    bool device_target = false;
    if (host_rgba) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, host_rgba) == cudaSuccess) {
            device_target = at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
            if (rp->output_format == r->storage_format() && (r->world == 1 || own)) {
                if (at.type == cudaMemoryTypeHost && at.devicePointer) host_alias = static_cast<float4*>(at.devicePointer);
                // (a target on ANOTHER device is reached by the copy below, not by this kernel's stores: peer access to it
                // is the caller's business and is not assumed)
                else if (device_target && at.devicePointer && (at.type == cudaMemoryTypeManaged || at.device == r->device))
                    host_alias = static_cast<float4*>(at.devicePointer);
            }
        } else {
            (void)cudaGetLastError();   // pageable memory on an old driver: not an error, just the copy path
        }
    }
    if (glsl) G.host_frame = taa ? nullptr : host_alias; else P.host_frame = taa ? nullptr : host_alias;
    // Fused gather: the producing kernel writes each finished pixel into the same frame on every peer.
    const bool peer_store = (rp->flags & GVT_FLAG_PEER_STORE) != 0 && r->world > 1 && !(rp->flags & GVT_FLAG_NO_GATHER);
    float4* peer_targets[GVT_MAX_PEERS];
    uint32_t n_peer = 0;
    if (peer_store) {
        const int idx = (r->frame == r->buf[0]) ? 0 : 1;
        for (int p = 0; p < r->world; p++) {
            if (p == r->rank) continue;
            peer_targets[n_peer++] = r->peer[p][idx];          // imported: checked before the swap
        }
        if (!taa && glsl) { for (uint32_t q = 0; q < n_peer; q++) G.peer_frame[q] = peer_targets[q]; G.n_peer = n_peer; }
        else if (!taa) { for (uint32_t q = 0; q < n_peer; q++) P.peer_frame[q] = peer_targets[q]; P.n_peer = n_peer; }
    }

    CK(cudaEventRecord(r->ev[0], r->stream));
    if (!glsl) {
        CK(cudaMemcpyAsync(r->d_block, r->h_block, sizeof(FrameBlock), cudaMemcpyHostToDevice, r->stream));
        h2d += sizeof(FrameBlock);
    } else {
        h2d += sizeof(GvtGlslUniforms);   // travels in the kernel parameter block
    }
    CK(cudaMemsetAsync(r->d_counters, 0, sizeof(Counters), r->stream));
    if (peer_store && r->frame == r->last_peer_target) {
        // Peers are about to write into the SAME buffer they filled last frame (no ping-pong happened: first TAA frame
        // after a history reset): everything this rank still had queued on that buffer (D2H copy) must be done first --
        // it is ahead of this barrier in stream order on every rank. When the buffers alternated, the closing barrier of
        // the previous frame already orders those reads before any peer's next write to that buffer, and this one is
        // skipped: one 4-byte all-reduce per frame instead of two.
        int nrc = g_nccl.AllReduce(r->d_sink, r->d_sink, 1, kNcclFloat32, kNcclSum, r->comm, r->stream);
        if (nrc != 0) return fail(GVT_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString(nrc));
    }
    if (peer_store) r->last_peer_target = r->frame;
    CK(cudaEventRecord(r->ev[1], r->stream));
    if (glsl) {
        if (interleave ? lattice_rows > 0 : G.y1 > G.y0) {
            if (glsl_precision == GVT_PRECISION_F32_FAST) CK(launch_fragment_glsl_fast(G, r->sm_count, r->stream));
            else CK(launch_fragment_glsl(G, (int)glsl_precision, r->sm_count, r->stream));
            launches++;
        }
    } else if (P.ny > 0) {
        const char* tl = getenv("GVT_TIMELINE_DUMP");          // diagnostics: per-warp start / end / tile count of this launch
        DevBuf tl_buf;
        const size_t tl_n = (size_t)r->sm_count * 32 * 3;
        if (tl && *tl) { CK(tl_buf.alloc(tl_n * 8)); CK(cudaMemsetAsync(tl_buf.as<char>(), 0, tl_n * 8, r->stream)); P.timeline = tl_buf.as<unsigned long long>(); }
        CK(launch_trace(P, rp->method, rp->precision, budget, false, r->sm_count, r->stream));
        launches++;
        if (P.timeline) {
            std::vector<unsigned long long> h(tl_n);
            CK(cudaMemcpyAsync(h.data(), P.timeline, tl_n * 8, cudaMemcpyDeviceToHost, r->stream));
            CK(cudaStreamSynchronize(r->stream));
            if (FILE* f = fopen(tl, "wb")) { fwrite(h.data(), 8, tl_n, f); fclose(f); }
        }
    }
    CK(cudaEventRecord(r->ev[2], r->stream));
    if (taa && (interleave ? n_my > 0 : row1 > row0)) {
        TaaParams T;
        if (cam) fill_taa(cam, W, H, T);
        else { memset(&T, 0, sizeof(T)); T.width = W; T.height = H; }
        T.cur = r->cur; T.hist = r->hist; T.out = r->frame; T.frame_f16 = r->frame_f16 ? 1u : 0u;
        T.row0 = row0; T.row1 = row1;
        T.stripe = sm; T.n_stripes = n_my;
        if (rp->flags & GVT_FLAG_TAA_WEBGL) {
            const bool has = rp->struct_size >= offsetof(GvtRenderParams, taa_camera_moving) + sizeof(uint32_t);
            T.mode = 1; T.blend = has ? rp->taa_blend : 0.75f; T.moving = has ? rp->taa_camera_moving : 0u;
        }
        T.host_out = host_alias;
        if (peer_store) { for (uint32_t q = 0; q < n_peer; q++) T.peer_out[q] = peer_targets[q]; T.n_peer = n_peer; }
        if (rp->flags & GVT_FLAG_TAA_PRECISE) CK(launch_taa_precise(T, r->stream));
        else CK(launch_taa(T, r->sm_count, r->stream));
        launches++;
    }
    CK(cudaEventRecord(r->ev[3], r->stream));
    if (peer_store) {
        // all ranks' stores must have landed before anyone reads its frame: a 1-element all-reduce is the barrier
        int nrc = g_nccl.AllReduce(r->d_sink, r->d_sink, 1, kNcclFloat32, kNcclSum, r->comm, r->stream);
        if (nrc != 0) return fail(GVT_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString(nrc));
    } else if (r->world > 1 && !(rp->flags & GVT_FLAG_NO_GATHER)) {
        const size_t count = (size_t)r->rows_per_rank * W * r->px_bytes() / 4;  // 4-byte words per rank block
        int nrc = g_nccl.AllGather(reinterpret_cast<const float*>(r->frame) + (size_t)r->rank * count, r->frame, count,
                                   kNcclFloat32, r->comm, r->stream);
        if (nrc != 0) return fail(GVT_ERR_NCCL, "ncclAllGather: %s", g_nccl.GetErrorString(nrc));
    }
    CK(cudaEventRecord(r->ev[4], r->stream));
    CK(cudaMemcpyAsync(r->h_counters, r->d_counters, sizeof(Counters), cudaMemcpyDeviceToHost, r->stream));
    d2h += sizeof(Counters);
    const uint64_t d2h_before_frame = d2h;
    if (host_rgba && host_alias) {
        d2h += (size_t)n_own * W * r->px_bytes();   // delivered by the kernel's own stores
    } else if (host_rgba && own && interleave) {
        const size_t row_bytes = (size_t)W * r->px_bytes();                 // (storage format only: checked before the swap)
        if (sm.s == 1u) {   // every world-th row: one strided copy
            const size_t pitch = row_bytes * (size_t)r->world;
            if (n_own)
                CK(cudaMemcpy2DAsync(static_cast<char*>(host_rgba) + (size_t)r->rank * row_bytes, pitch,
                                     reinterpret_cast<const char*>(r->frame) + (size_t)r->rank * row_bytes, pitch, row_bytes, n_own,
                                     cudaMemcpyDefault, r->stream));
        } else {            // 16-row stripes: one contiguous copy each
            for (uint32_t t = 0; t < n_my; t++) {
                const size_t y = (size_t)(t * (uint32_t)r->world + (uint32_t)r->rank) * sm.s;
                const size_t n = std::min<size_t>(sm.s, H - y);
                CK(cudaMemcpyAsync(static_cast<char*>(host_rgba) + y * row_bytes, reinterpret_cast<const char*>(r->frame) + y * row_bytes,
                                   n * row_bytes, cudaMemcpyDefault, r->stream));
            }
        }
        d2h += (size_t)n_own * row_bytes;
    } else if (host_rgba) {
        const size_t n_px = (size_t)W * H;
        // which pixels go back: the whole frame, or only this rank's row block at its place in the host frame
        const size_t px0 = own ? (size_t)row0 * W : 0, npx = own ? (size_t)(row1 - row0) * W : n_px;
        if (rp->output_format != r->storage_format()) {
            const size_t bpp = format_bytes(rp->output_format);
            if (npx) {
                int32_t crc = convert_frame(r, rp->output_format, px0, npx);
                if (crc != GVT_OK) return crc;
                launches++;
                CK(cudaMemcpyAsync(static_cast<char*>(host_rgba) + px0 * bpp, static_cast<char*>(r->half_frame) + px0 * bpp,
                                   npx * bpp, cudaMemcpyDefault, r->stream));
            }
            d2h += npx * bpp;
        } else {
            if (npx)
                CK(cudaMemcpyAsync(static_cast<char*>(host_rgba) + px0 * r->px_bytes(), r->at(r->frame, px0), npx * r->px_bytes(),
                                   cudaMemcpyDefault, r->stream));
            d2h += npx * r->px_bytes();
        }
    }
    if (device_target) d2h = d2h_before_frame;   // a device-memory target: the frame stayed on the device
    CK(cudaEventRecord(r->ev[5], r->stream));
    CK(cudaStreamSynchronize(r->stream));
    if (r->comm && g_nccl.CommGetAsyncError) {
        // a collective can fail after it was enqueued (peer death, link error): surface it instead of a silent bad frame
        int aerr = 0;
        const int qrc = g_nccl.CommGetAsyncError(r->comm, &aerr);
        if (qrc != 0 || aerr != 0)
            return fail(GVT_ERR_NCCL, "NCCL asynchronous error: %s", g_nccl.GetErrorString(qrc != 0 ? qrc : aerr));
    }
    late.armed = false;
    if (taa) r->history_valid = true;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r->ev[1], r->ev[2]); stats->trace_ms = ms;
        cudaEventElapsedTime(&ms, r->ev[2], r->ev[3]); stats->taa_ms = taa ? ms : 0.0;
        cudaEventElapsedTime(&ms, r->ev[3], r->ev[4]); stats->gather_ms = (r->world > 1) ? ms : 0.0;
        cudaEventElapsedTime(&ms, r->ev[0], r->ev[5]); stats->total_ms = ms;
        const Counters& c = *r->h_counters;
        stats->steps_committed = c.steps_committed; stats->steps_executed = c.steps_executed; stats->rhs_evals = c.rhs_evals;
        stats->n_horizon = c.n_horizon; stats->n_escape = c.n_escape; stats->n_maxsteps = c.n_maxsteps; stats->n_disk = c.n_disk;
        stats->h2d_bytes = h2d; stats->d2h_bytes = d2h; stats->kernel_launches = launches;
        stats->rows_begin = interleave ? (uint32_t)r->rank : row0; stats->rows_end = interleave ? H : row1;   // interleaved: every world-th row from rows_begin
    }
    return GVT_OK;
}

// webgpu/renderer.ts:280-411 render(camera, physics)
extern "C" int32_t gvt_render_frame(gvt_renderer* r, const GvtCamera* cam, const GvtPhysicsParams* phys,
                                    const GvtRenderParams* rp, void* host_rgba, GvtFrameStats* stats) {
    if (!r || !cam || !phys || !rp) return fail(GVT_ERR_INVALID, "null argument");
    return render_impl(r, cam, phys, rp, nullptr, 0, host_rgba, stats);
}

// One row block [row0, row1) of the frame on a single-GPU renderer: what one rank of an N-GPU run traces, for tiled
// rendering by a host that schedules the blocks itself and for measuring a rank's share of a frame on one device.
extern "C" int32_t gvt_render_rows(gvt_renderer* r, const GvtCamera* cam, const GvtPhysicsParams* phys, const GvtRenderParams* rp,
                                   uint32_t row0, uint32_t row1, void* host_rgba, GvtFrameStats* stats) {
    if (!r || !cam || !phys || !rp) return fail(GVT_ERR_INVALID, "null argument");
    if (r->world != 1) return fail(GVT_ERR_UNSUPPORTED, "gvt_render_rows is for single-GPU renderers (ranks already own a row block)");
    if (row1 <= row0 || (rp->flags & (GVT_FLAG_TAA | GVT_FLAG_ROW_INTERLEAVE))) return fail(GVT_ERR_INVALID, "bad row block / flags");
    r->rows_override[0] = row0; r->rows_override[1] = row1;
    const int32_t rc = render_impl(r, cam, phys, rp, nullptr, 0, host_rgba, stats);
    r->rows_override[0] = r->rows_override[1] = 0u;
    return rc;
}

// webgl-utils.ts:259-303: createNoiseTexture / createBlueNoiseTexture (256x256 RGBA8). Only .r is ever sampled
// (chunks/noise.ts:7, fragment.glsl.ts:106), so only the red channels are kept on the device.
extern "C" int32_t gvt_render_set_noise_textures(gvt_renderer* r, const uint8_t* noise_rgba8, const uint8_t* blue_rgba8,
                                                 uint32_t size) {
    if (!r || !noise_rgba8 || !blue_rgba8) return fail(GVT_ERR_INVALID, "null argument");
    if (size != 256) return fail(GVT_ERR_INVALID, "noise textures are 256x256 (webgl/renderer.ts:124-125), got %u", size);
    CK(cudaSetDevice(r->device));
    std::vector<uint8_t> red(2 * 65536);
    for (size_t i = 0; i < 65536; i++) { red[i] = noise_rgba8[4 * i]; red[65536 + i] = blue_rgba8[4 * i]; }
    if (!r->d_noise_r) CK(cudaMalloc(&r->d_noise_r, 65536));
    if (!r->d_blue_r) CK(cudaMalloc(&r->d_blue_r, 65536));
    CK(cudaMemcpyAsync(r->d_noise_r, red.data(), 65536, cudaMemcpyHostToDevice, r->stream));
    CK(cudaMemcpyAsync(r->d_blue_r, red.data() + 65536, 65536, cudaMemcpyHostToDevice, r->stream));
    CK(cudaStreamSynchronize(r->stream));
    return GVT_OK;
}

// webgl/renderer.ts:173-420 render(params, mouse): the fragment-shader pass (+ the WebGL2 TAA resolve on request)
extern "C" int32_t gvt_render_fragment_glsl(gvt_renderer* r, const GvtGlslUniforms* u, uint32_t precision, uint32_t flags,
                                            uint32_t output_format, float taa_blend, uint32_t taa_camera_moving,
                                            void* host_rgba, GvtFrameStats* stats) {
    if (!r || !u) return fail(GVT_ERR_INVALID, "null argument");
    if (u->struct_size != sizeof(GvtGlslUniforms)) return fail(GVT_ERR_INVALID, "GvtGlslUniforms.struct_size = %u, expected %zu", u->struct_size, sizeof(GvtGlslUniforms));
    if (precision > GVT_PRECISION_F32_FAST || output_format > GVT_FORMAT_RGBA8_UNORM) return fail(GVT_ERR_INVALID, "bad precision / output format");
    if (!(u->resolution[0] >= 1.0f) || !(u->resolution[1] >= 1.0f) || u->resolution[0] > 65536.0f || u->resolution[1] > 65536.0f)
        return fail(GVT_ERR_INVALID, "bad u_resolution");
    if (!(u->mass > 0.0f)) return fail(GVT_ERR_INVALID, "u_mass must be positive");
    GvtRenderParams rp;
    gvt_render_params_default(&rp);
    rp.flags = flags & ~(uint32_t)(GVT_FLAG_BUDGET | GVT_FLAG_JITTER);
    rp.output_format = output_format;
    rp.taa_blend = taa_blend; rp.taa_camera_moving = taa_camera_moving;
    return render_impl(r, nullptr, nullptr, &rp, u, precision, host_rgba, stats);
}

// rendering/bloom.ts:446-632 on the finished frame
extern "C" int32_t gvt_render_bloom(gvt_renderer* r, const GvtBloomConfig* cfg, uint32_t output_format, void* host_out, double* ms) {
    if (!r || !cfg) return fail(GVT_ERR_INVALID, "null argument");
    if (cfg->struct_size < offsetof(GvtBloomConfig, precise) || cfg->struct_size > sizeof(GvtBloomConfig))
        return fail(GVT_ERR_INVALID, "GvtBloomConfig.struct_size = %u", cfg->struct_size);
    const bool bloom_precise = cfg->struct_size >= sizeof(GvtBloomConfig) && cfg->precise != 0;
    if (!r->frame || r->width == 0) return fail(GVT_ERR_INVALID, "no frame rendered yet");
    if (output_format != GVT_FORMAT_RGBA32F && output_format != GVT_FORMAT_RGBA16F && output_format != GVT_FORMAT_RGBA8_UNORM)
        return fail(GVT_ERR_INVALID, "bloom output is display-referred: RGBA32F, RGBA16F or RGBA8_UNORM");
    if (cfg->blur_passes > 16) return fail(GVT_ERR_INVALID, "blur_passes > 16");
    CK(cudaSetDevice(r->device));
    const int W = (int)r->width, H = (int)r->height;
    if (r->bloom_w != r->width || r->bloom_h != r->height || !r->display) {
        CK(cudaStreamSynchronize(r->stream));
        if (r->display) cudaFree(r->display);
        if (r->bloom_half) cudaFree(r->bloom_half);
        if (r->bloom_q1) cudaFree(r->bloom_q1);
        if (r->bloom_q2) cudaFree(r->bloom_q2);
        r->display = nullptr; r->bloom_half = r->bloom_q1 = r->bloom_q2 = nullptr; r->bloom_w = r->bloom_h = 0;
        const size_t hw = (size_t)std::max(1, W / 2), hh = (size_t)std::max(1, H / 2), bw = (size_t)std::max(1, W / 4), bh = (size_t)std::max(1, H / 4);
        CK(cudaMalloc(&r->display, (size_t)W * H * sizeof(float4)));
        CK(cudaMalloc(&r->bloom_half, hw * hh * sizeof(uint2)));
        CK(cudaMalloc(&r->bloom_q1, bw * bh * sizeof(uint2)));
        CK(cudaMalloc(&r->bloom_q2, bw * bh * sizeof(uint2)));
        r->bloom_w = r->width; r->bloom_h = r->height;
    }
    int launches = 0;
    CK(cudaEventRecord(r->ev[0], r->stream));
    CK(launch_bloom(r->frame, W, H, r->bloom_half, r->bloom_q1, r->bloom_q2, r->display, cfg->threshold, cfg->intensity,
                    (int)cfg->blur_passes, cfg->enabled ? 1 : 0, r->sm_count, r->stream, &launches, bloom_precise, r->frame_f16));
    CK(cudaEventRecord(r->ev[1], r->stream));
    if (host_out) {
        const size_t n_px = (size_t)W * H;
        if (output_format == GVT_FORMAT_RGBA32F) {
            CK(cudaMemcpyAsync(host_out, r->display, n_px * sizeof(float4), cudaMemcpyDefault, r->stream));
        } else {
            if (!r->half_frame) CK(cudaMalloc(&r->half_frame, n_px * 16));
            if (output_format == GVT_FORMAT_RGBA16F) CK(launch_f32_to_f16(r->display, r->half_frame, n_px, r->stream));
            else CK(launch_tonemap_rgba8(r->display, r->half_frame, n_px, 2, r->stream));
            CK(cudaMemcpyAsync(host_out, r->half_frame, n_px * format_bytes(output_format), cudaMemcpyDefault, r->stream));
        }
    }
    CK(cudaStreamSynchronize(r->stream));
    if (ms) { float t = 0.f; cudaEventElapsedTime(&t, r->ev[0], r->ev[1]); *ms = t; }
    return GVT_OK;
}

extern "C" int32_t gvt_render_fragment_glsl_debug(gvt_renderer* r, uint32_t* steps, uint32_t* hit) {
    if (!r || !r->d_gsteps || !r->d_ghit || !r->g_valid) return fail(GVT_ERR_INVALID, "the last fragment-shader frame was not rendered with GVT_FLAG_DEBUG_COUNTS");
    CK(cudaSetDevice(r->device));
    const size_t n = (size_t)r->width * r->height;
    if (n > r->g_cap) return fail(GVT_ERR_INVALID, "frame was resized since the last fragment-shader frame");
    if (steps) CK(cudaMemcpyAsync(steps, r->d_gsteps, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, r->stream));
    if (hit) CK(cudaMemcpyAsync(hit, r->d_ghit, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, r->stream));
    CK(cudaStreamSynchronize(r->stream));
    return GVT_OK;
}

static int32_t ensure_dbg(gvt_renderer* r, size_t n) {
    if (n <= r->dbg_cap) return GVT_OK;
    if (r->d_xp) cudaFree(r->d_xp);
    if (r->d_drift) cudaFree(r->d_drift);
    if (r->d_rgba) cudaFree(r->d_rgba);
    if (r->d_term) cudaFree(r->d_term);
    if (r->d_steps) cudaFree(r->d_steps);
    r->d_xp = r->d_drift = r->d_rgba = nullptr; r->d_term = r->d_steps = nullptr; r->dbg_cap = 0;
    CK(cudaMalloc(&r->d_xp, n * 8 * sizeof(double)));
    CK(cudaMalloc(&r->d_drift, n * sizeof(double)));
    CK(cudaMalloc(&r->d_rgba, n * 4 * sizeof(double)));
    CK(cudaMalloc(&r->d_term, n * sizeof(uint32_t)));
    CK(cudaMalloc(&r->d_steps, n * sizeof(uint32_t)));
    r->dbg_cap = n;
    return GVT_OK;
}

extern "C" int32_t gvt_trace_states(gvt_renderer* r, const GvtCamera* cam, const GvtPhysicsParams* phys,
                                    const GvtRenderParams* rp, uint32_t x0, uint32_t xs, uint32_t y0, uint32_t y1,
                                    uint32_t ys, double* out_xp8, uint32_t* term, uint32_t* steps, double* max_drift,
                                    double* rgba64) {
    if (!r || !cam || !phys || !rp) return fail(GVT_ERR_INVALID, "null argument");
    CK(cudaSetDevice(r->device));
    FrameParams P;
    int32_t rc = build_frame(r, cam, phys, rp, P);
    if (rc != GVT_OK) return rc;
    if (xs == 0 || ys == 0 || x0 >= P.width || y1 > P.height || y0 >= y1) return fail(GVT_ERR_INVALID, "bad pixel lattice");
    P.x0 = x0; P.xs = xs; P.y0 = y0; P.y1 = y1; P.ys = ys;
    P.nx = (P.width - x0 + xs - 1) / xs; P.ny = (y1 - y0 + ys - 1) / ys;
    const size_t n = (size_t)P.nx * P.ny;
    rc = ensure_dbg(r, n);
    if (rc != GVT_OK) return rc;
    P.frame = nullptr;
    P.dbg_xp = r->d_xp; P.dbg_term = r->d_term; P.dbg_steps = r->d_steps; P.dbg_drift = r->d_drift; P.dbg_rgba = r->d_rgba;
    const bool budget = (rp->flags & GVT_FLAG_BUDGET) != 0 && (rp->method == GVT_METHOD_RK4 || rp->method == GVT_METHOD_SYMPLECTIC);
    CK(cudaMemcpyAsync(r->d_block, r->h_block, sizeof(FrameBlock), cudaMemcpyHostToDevice, r->stream));
    CK(cudaMemsetAsync(r->d_counters, 0, sizeof(Counters), r->stream));
    CK(launch_trace(P, rp->method, rp->precision, budget, true, r->sm_count, r->stream));
    if (out_xp8) CK(cudaMemcpyAsync(out_xp8, r->d_xp, n * 8 * sizeof(double), cudaMemcpyDeviceToHost, r->stream));
    if (term) CK(cudaMemcpyAsync(term, r->d_term, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, r->stream));
    if (steps) CK(cudaMemcpyAsync(steps, r->d_steps, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, r->stream));
    if (max_drift) CK(cudaMemcpyAsync(max_drift, r->d_drift, n * sizeof(double), cudaMemcpyDeviceToHost, r->stream));
    if (rgba64) CK(cudaMemcpyAsync(rgba64, r->d_rgba, n * 4 * sizeof(double), cudaMemcpyDeviceToHost, r->stream));
    CK(cudaStreamSynchronize(r->stream));
    return GVT_OK;
}

static int32_t taa_host_frames(gvt_renderer* r, TaaParams& T, uint32_t width, uint32_t height, const void* cur, const void* hist,
                               void* out, int32_t precise, double* ms_out) {
    CK(cudaSetDevice(r->device));
    const size_t bytes = (size_t)width * height * (T.frame_f16 ? 8 : 16);
    DevBuf b_cur, b_hist, b_out;
    CK(b_cur.alloc(bytes)); CK(b_hist.alloc(bytes)); CK(b_out.alloc(bytes));
    float4 *d_cur = b_cur.as<float4>(), *d_hist = b_hist.as<float4>(), *d_out = b_out.as<float4>();
    CK(cudaMemcpyAsync(d_cur, cur, bytes, cudaMemcpyHostToDevice, r->stream));
    CK(cudaMemcpyAsync(d_hist, hist, bytes, cudaMemcpyHostToDevice, r->stream));
    T.cur = d_cur; T.hist = d_hist; T.out = d_out;
    CK(cudaEventRecord(r->ev[0], r->stream));
    if (precise) CK(launch_taa_precise(T, r->stream)); else CK(launch_taa(T, r->sm_count, r->stream));
    CK(cudaEventRecord(r->ev[1], r->stream));
    CK(cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, r->stream));
    CK(cudaStreamSynchronize(r->stream));
    if (ms_out) { float t = 0.f; cudaEventElapsedTime(&t, r->ev[0], r->ev[1]); *ms_out = t; }
    return GVT_OK;
}

extern "C" int32_t gvt_taa_resolve_ex(gvt_renderer* r, const GvtCamera* cam, uint32_t width, uint32_t height, const void* cur,
                                      const void* hist, void* out, uint32_t frame_format, uint32_t webgl, float blend,
                                      int32_t camera_moving, int32_t precise, double* ms_out) {
    if (!r || (!cam && !webgl) || !cur || !hist || !out || width == 0 || height == 0) return fail(GVT_ERR_INVALID, "bad argument");
    if (frame_format != GVT_FORMAT_RGBA32F && frame_format != GVT_FORMAT_RGBA16F) return fail(GVT_ERR_INVALID, "frame_format: RGBA32F or RGBA16F");
    TaaParams T;
    if (cam) fill_taa(cam, width, height, T);
    else { memset(&T, 0, sizeof(T)); T.width = width; T.height = height; T.row1 = height; T.stripe = StripeMap{0u, 0u, 1u, 0u}; }
    if (webgl) { T.mode = 1; T.blend = blend; T.moving = camera_moving ? 1u : 0u; }
    T.frame_f16 = frame_format == GVT_FORMAT_RGBA16F ? 1u : 0u;
    return taa_host_frames(r, T, width, height, cur, hist, out, precise, ms_out);
}

extern "C" int32_t gvt_taa_resolve(gvt_renderer* r, const GvtCamera* cam, uint32_t width, uint32_t height, const float* cur,
                                   const float* hist, float* out) {
    if (!r || !cam || !cur || !hist || !out || width == 0 || height == 0) return fail(GVT_ERR_INVALID, "bad argument");
    CK(cudaSetDevice(r->device));
    const size_t bytes = (size_t)width * height * sizeof(float4);
    DevBuf b_cur, b_hist, b_out;
    CK(b_cur.alloc(bytes)); CK(b_hist.alloc(bytes)); CK(b_out.alloc(bytes));
    float4 *d_cur = b_cur.as<float4>(), *d_hist = b_hist.as<float4>(), *d_out = b_out.as<float4>();
    CK(cudaMemcpyAsync(d_cur, cur, bytes, cudaMemcpyHostToDevice, r->stream));
    CK(cudaMemcpyAsync(d_hist, hist, bytes, cudaMemcpyHostToDevice, r->stream));
    TaaParams T;
    fill_taa(cam, width, height, T);
    T.cur = d_cur; T.hist = d_hist; T.out = d_out;
    CK(launch_taa(T, r->sm_count, r->stream));
    CK(cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, r->stream));
    CK(cudaStreamSynchronize(r->stream));
    return GVT_OK;
}

extern "C" int32_t gvt_taa_resolve_webgl(gvt_renderer* r, uint32_t width, uint32_t height, const float* cur, const float* hist,
                                         float blend, int32_t camera_moving, float* out) {
    if (!r || !cur || !hist || !out || width == 0 || height == 0) return fail(GVT_ERR_INVALID, "bad argument");
    CK(cudaSetDevice(r->device));
    const size_t bytes = (size_t)width * height * sizeof(float4);
    DevBuf b_cur, b_hist, b_out;
    CK(b_cur.alloc(bytes)); CK(b_hist.alloc(bytes)); CK(b_out.alloc(bytes));
    float4 *d_cur = b_cur.as<float4>(), *d_hist = b_hist.as<float4>(), *d_out = b_out.as<float4>();
    CK(cudaMemcpyAsync(d_cur, cur, bytes, cudaMemcpyHostToDevice, r->stream));
    CK(cudaMemcpyAsync(d_hist, hist, bytes, cudaMemcpyHostToDevice, r->stream));
    TaaParams T;
    memset(&T, 0, sizeof(T));
    T.width = width; T.height = height; T.row0 = 0; T.row1 = height;
    T.mode = 1; T.blend = blend; T.moving = camera_moving ? 1u : 0u;
    T.cur = d_cur; T.hist = d_hist; T.out = d_out;
    CK(launch_taa(T, r->sm_count, r->stream));
    CK(cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, r->stream));
    CK(cudaStreamSynchronize(r->stream));
    return GVT_OK;
}

extern "C" int32_t gvt_host_alloc(size_t bytes, void** out) {
    if (!out || bytes == 0) return fail(GVT_ERR_INVALID, "bad argument");
    CK(cudaMallocHost(out, bytes));
    return GVT_OK;
}
extern "C" int32_t gvt_host_free(void* p) {
    if (p) CK(cudaFreeHost(p));
    return GVT_OK;
}
extern "C" int32_t gvt_host_register(void* p, size_t bytes) {
    if (!p || bytes == 0) return fail(GVT_ERR_INVALID, "bad argument");
    CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return GVT_OK;
}
extern "C" int32_t gvt_host_unregister(void* p) {
    if (p) CK(cudaHostUnregister(p));
    return GVT_OK;
}

// --------------------------------------------------------------------------------------------------
// External-memory hand-off: the presenter allocates the frame image / buffer in its graphics API with an
// exportable POSIX fd (VK_KHR_external_memory_fd, GL_EXT_memory_object_fd -- or a CUDA VMM allocation exported with
// cuMemExportToShareableHandle), this side maps it and every frame entry point accepts the mapped pointer as its target.
// --------------------------------------------------------------------------------------------------
struct gvt_external_buffer { cudaExternalMemory_t mem = nullptr; void* ptr = nullptr; int device = 0; };
struct gvt_external_semaphore { cudaExternalSemaphore_t sem = nullptr; int device = 0; };

extern "C" int32_t gvt_external_import_fd(int32_t device, int32_t fd, uint64_t bytes, int32_t dedicated,
                                          gvt_external_buffer** out, void** device_ptr) {
    if (!out || !device_ptr || fd < 0 || bytes == 0) return fail(GVT_ERR_INVALID, "bad argument");
    int nd = 0;
    if (cudaGetDeviceCount(&nd) != cudaSuccess || nd == 0) return fail(GVT_ERR_NO_DEVICE, "no CUDA device");
    if (device < 0 || device >= nd) return fail(GVT_ERR_INVALID, "device %d out of range (%d devices)", device, nd);
    CK(cudaSetDevice(device));
    cudaExternalMemoryHandleDesc hd;
    memset(&hd, 0, sizeof(hd));
    hd.type = cudaExternalMemoryHandleTypeOpaqueFd; hd.handle.fd = fd; hd.size = bytes;
    hd.flags = dedicated ? cudaExternalMemoryDedicated : 0;
    gvt_external_buffer* b = new (std::nothrow) gvt_external_buffer();
    if (!b) return fail(GVT_ERR_INVALID, "out of memory");
    b->device = device;
    cudaError_t e = cudaImportExternalMemory(&b->mem, &hd);     // on success the driver owns the fd
    if (e != cudaSuccess) { delete b; (void)cudaGetLastError(); return fail(GVT_ERR_CUDA, "cudaImportExternalMemory: %s", cudaGetErrorString(e)); }
    cudaExternalMemoryBufferDesc bd;
    memset(&bd, 0, sizeof(bd));
    bd.offset = 0; bd.size = bytes;
    e = cudaExternalMemoryGetMappedBuffer(&b->ptr, b->mem, &bd);
    if (e != cudaSuccess) { cudaDestroyExternalMemory(b->mem); delete b; (void)cudaGetLastError(); return fail(GVT_ERR_CUDA, "cudaExternalMemoryGetMappedBuffer: %s", cudaGetErrorString(e)); }
    *out = b; *device_ptr = b->ptr;
    return GVT_OK;
}
extern "C" int32_t gvt_external_release(gvt_external_buffer* b) {
    if (!b) return GVT_OK;
    cudaSetDevice(b->device);
    if (b->ptr) cudaFree(b->ptr);                               // the documented way to unmap a mapped external buffer
    if (b->mem) cudaDestroyExternalMemory(b->mem);
    delete b;
    return GVT_OK;
}
extern "C" int32_t gvt_external_semaphore_import_fd(int32_t device, int32_t fd, int32_t timeline, gvt_external_semaphore** out) {
    if (!out || fd < 0) return fail(GVT_ERR_INVALID, "bad argument");
    int nd = 0;
    if (cudaGetDeviceCount(&nd) != cudaSuccess || nd == 0) return fail(GVT_ERR_NO_DEVICE, "no CUDA device");
    if (device < 0 || device >= nd) return fail(GVT_ERR_INVALID, "device %d out of range (%d devices)", device, nd);
    CK(cudaSetDevice(device));
    cudaExternalSemaphoreHandleDesc hd;
    memset(&hd, 0, sizeof(hd));
    hd.type = timeline ? cudaExternalSemaphoreHandleTypeTimelineSemaphoreFd : cudaExternalSemaphoreHandleTypeOpaqueFd;
    hd.handle.fd = fd;
    gvt_external_semaphore* x = new (std::nothrow) gvt_external_semaphore();
    if (!x) return fail(GVT_ERR_INVALID, "out of memory");
    x->device = device;
    cudaError_t e = cudaImportExternalSemaphore(&x->sem, &hd);
    if (e != cudaSuccess) { delete x; (void)cudaGetLastError(); return fail(GVT_ERR_CUDA, "cudaImportExternalSemaphore: %s", cudaGetErrorString(e)); }
    *out = x;
    return GVT_OK;
}
extern "C" int32_t gvt_external_semaphore_release(gvt_external_semaphore* x) {
    if (!x) return GVT_OK;
    cudaSetDevice(x->device);
    if (x->sem) cudaDestroyExternalSemaphore(x->sem);
    delete x;
    return GVT_OK;
}
// Enqueued on the renderer's stream: the next frame's kernels start after the wait, a signal fires after everything
// already enqueued (every frame entry point returns with its stream drained, so a signal right after it is immediate).
extern "C" int32_t gvt_render_wait_external(gvt_renderer* r, gvt_external_semaphore* x, uint64_t value) {
    if (!r || !x) return fail(GVT_ERR_INVALID, "null argument");
    CK(cudaSetDevice(r->device));
    cudaExternalSemaphoreWaitParams wp;
    memset(&wp, 0, sizeof(wp));
    wp.params.fence.value = value;
    CK(cudaWaitExternalSemaphoresAsync(&x->sem, &wp, 1, r->stream));
    return GVT_OK;
}
extern "C" int32_t gvt_render_signal_external(gvt_renderer* r, gvt_external_semaphore* x, uint64_t value) {
    if (!r || !x) return fail(GVT_ERR_INVALID, "null argument");
    CK(cudaSetDevice(r->device));
    cudaExternalSemaphoreSignalParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.params.fence.value = value;
    CK(cudaSignalExternalSemaphoresAsync(&x->sem, &sp, 1, r->stream));
    return GVT_OK;
}

extern "C" int32_t gvt_measure_fma_peak(gvt_renderer* r, int32_t precision, double* out_tflops, double* out_ms) {
    if (!r || !out_tflops) return fail(GVT_ERR_INVALID, "null argument");
    CK(cudaSetDevice(r->device));
    double flops = 0.0;
    // warm-up, then a launch long enough (tens of ms) for the clocks to settle
    CK(launch_fma_peak(precision, r->sm_count, 2000, r->d_sink, r->stream, &flops));
    const unsigned long long iters = precision == GVT_PRECISION_F32 ? 400000ull : 200000ull;
    CK(cudaEventRecord(r->ev[0], r->stream));
    CK(launch_fma_peak(precision, r->sm_count, iters, r->d_sink, r->stream, &flops));
    CK(cudaEventRecord(r->ev[1], r->stream));
    CK(cudaStreamSynchronize(r->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, r->ev[0], r->ev[1]));
    *out_tflops = flops / ((double)ms * 1e-3) * 1e-12;
    if (out_ms) *out_ms = ms;
    return GVT_OK;
}

extern "C" int32_t gvt_device_info(gvt_renderer* r, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, char name[256]) {
    if (!r) return fail(GVT_ERR_INVALID, "null renderer");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, r->device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (name) { strncpy(name, prop.name, 255); name[255] = 0; }
    return GVT_OK;
}

extern "C" int32_t gvt_render_export_frames(gvt_renderer* r, uint8_t handles[2][64]) {
    if (!r || !handles || !r->buf[0]) return fail(GVT_ERR_INVALID, "bad argument / no frame buffers (call gvt_render_resize first)");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    CK(cudaSetDevice(r->device));
    for (int i = 0; i < 2; i++) {
        cudaIpcMemHandle_t h;
        CK(cudaIpcGetMemHandle(&h, r->buf[i]));
        memcpy(handles[i], &h, 64);
    }
    return GVT_OK;
}

extern "C" int32_t gvt_render_import_peer_frames(gvt_renderer* r, int32_t peer_rank, const uint8_t handles[2][64]) {
    if (!r || !handles || peer_rank < 0 || peer_rank >= GVT_MAX_PEERS || peer_rank >= r->world || peer_rank == r->rank)
        return fail(GVT_ERR_INVALID, "bad peer rank");
    CK(cudaSetDevice(r->device));
    if (r->peer_open[peer_rank]) {
        cudaIpcCloseMemHandle(r->peer[peer_rank][0]); cudaIpcCloseMemHandle(r->peer[peer_rank][1]);
        r->peer_open[peer_rank] = false;
    }
    for (int i = 0; i < 2; i++) {
        cudaIpcMemHandle_t h;
        memcpy(&h, handles[i], 64);
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        r->peer[peer_rank][i] = static_cast<float4*>(p);
    }
    r->peer_open[peer_rank] = true;
    return GVT_OK;
}
