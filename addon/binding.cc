// addon/binding.cc — thin N-API shim over libgravitas_b200.so (C ABI in include/gravitas_b200.h).
// Exposes the two seams of the reference to TypeScript:
//   class PhysicsEngine  — same method names as the wasm-bindgen class (physics-engine/gravitas-wasm/src/lib.rs:42-465)
//   class KerrRenderer   — init/resize/renderFrame for the src/rendering renderer API (rendering/webgpu/renderer.ts:82-411)
// Build (on a machine with node): node-gyp rebuild (addon/binding.gyp). Not load-tested in the build image (no node, no
// node headers); type-checked against addon/stub/node_api.h -- a verbatim subset of Node's node_api.h / js_native_api.h
// declarations -- by tests/test_abi_host.py. Runtime: a Node-API host with a DOM, i.e. an Electron / NW.js renderer (and
// its workers, nodeIntegrationInWorker), see INTEGRATION.md 4.
#include <node_api.h>
#include <stdint.h>
#include <string.h>
#include "gravitas_b200.h"

#define NAPI_OK(c) do { if ((c) != napi_ok) { napi_throw_error(env, nullptr, #c); return nullptr; } } while (0)
#define GVT(c) do { if ((c) != GVT_OK) { napi_throw_error(env, nullptr, gvt_last_error()); return nullptr; } } while (0)

namespace {

napi_value Undefined(napi_env env) { napi_value u; napi_get_undefined(env, &u); return u; }
double Num(napi_env env, napi_value v) { double d = 0; napi_get_value_double(env, v, &d); return d; }

template <class T> T* Self(napi_env env, napi_callback_info info, size_t* argc, napi_value* argv) {
    napi_value self; void* p = nullptr;
    if (napi_get_cb_info(env, info, argc, argv, &self, nullptr) != napi_ok) return nullptr;
    napi_unwrap(env, self, &p);
    return static_cast<T*>(p);
}
// What a JS PhysicsEngine wraps: the native engine plus the view it writes its SAB block into. The view is held by a
// strong reference so that the memory outlives every tick_sab (the engine keeps a raw pointer into it).
struct EngineBox {
    gvt_engine* e = nullptr;
    napi_ref sab_ref = nullptr;      // Float32Array (over a SharedArrayBuffer or ArrayBuffer) passed to attach_sab
    uint32_t sab_byte_offset = 0;    // its byteOffset inside that buffer: what get_sab_ptr() returns (wasm: offset into memory.buffer)
    bool attached = false;
};
gvt_engine* Eng(napi_env env, napi_callback_info info, size_t* argc, napi_value* argv, EngineBox** box_out = nullptr) {
    EngineBox* b = Self<EngineBox>(env, info, argc, argv);
    if (box_out) *box_out = b;
    return b ? b->e : nullptr;
}
// A caller-provided byte range: an ArrayBuffer, or any TypedArray / DataView-less view (also over a SharedArrayBuffer, which
// napi_get_arraybuffer_info rejects on most Node versions while napi_get_typedarray_info accepts views of it).
bool ByteRange(napi_env env, napi_value v, void** data, size_t* bytes, napi_typedarray_type* type_out = nullptr) {
    bool is_ta = false;
    if (napi_is_typedarray(env, v, &is_ta) == napi_ok && is_ta) {
        napi_typedarray_type t; size_t n = 0; napi_value ab; size_t off = 0;
        if (napi_get_typedarray_info(env, v, &t, &n, data, &ab, &off) != napi_ok) return false;
        static const size_t elem[] = {1, 1, 1, 2, 2, 4, 4, 4, 8, 8, 8};
        *bytes = n * ((size_t)t < sizeof(elem) / sizeof(elem[0]) ? elem[t] : 1);
        if (type_out) *type_out = t;
        return true;
    }
    bool is_ab = false;
    if (napi_is_arraybuffer(env, v, &is_ab) == napi_ok && is_ab) {
        if (type_out) *type_out = napi_uint8_array;
        return napi_get_arraybuffer_info(env, v, data, bytes) == napi_ok;
    }
    return false;
}
napi_value F32Array(napi_env env, size_t n, float** data) {
    napi_value ab, out; void* p;
    if (napi_create_arraybuffer(env, n * 4, &p, &ab) != napi_ok) return nullptr;
    *data = static_cast<float*>(p);
    napi_create_typedarray(env, napi_float32_array, n, ab, 0, &out);
    return out;
}

// ---------------------------------------------------------------- PhysicsEngine (Seam A)
napi_value EngineNew(napi_env env, napi_callback_info info) {                    // lib.rs:59
    size_t argc = 2; napi_value argv[2], self;
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, &self, nullptr));
    EngineBox* box = new EngineBox();
    if (gvt_engine_create(Num(env, argv[0]), Num(env, argv[1]), &box->e) != GVT_OK) {
        delete box;
        napi_throw_error(env, nullptr, gvt_last_error());
        return nullptr;
    }
    NAPI_OK(napi_wrap(env, self, box, [](napi_env fenv, void* d, void*) {
        EngineBox* b = static_cast<EngineBox*>(d);
        if (b->sab_ref) napi_delete_reference(fenv, b->sab_ref);
        gvt_engine_destroy(b->e);
        delete b;
    }, nullptr, nullptr));
    return self;
}
napi_value UpdateParams(napi_env env, napi_callback_info info) {                 // lib.rs:78
    size_t argc = 2; napi_value argv[2];
    gvt_engine* e = Eng(env, info, &argc, argv);
    GVT(gvt_engine_update_params(e, Num(env, argv[0]), Num(env, argv[1])));
    return Undefined(env);
}
napi_value TickSab(napi_env env, napi_callback_info info) {                      // lib.rs:308
    size_t argc = 1; napi_value argv[1];
    gvt_engine* e = Eng(env, info, &argc, argv);
    GVT(gvt_engine_tick_sab(e, Num(env, argv[0])));
    return Undefined(env);
}
// attach_sab(view: Float32Array)            lib.rs:74 -- the engine then writes its SAB block into the caller's memory in place.
// `view` is a Float32Array of >= 2048 elements over a SharedArrayBuffer (the worker's SAB, or the module "memory" of
// addon/ts/index.ts) or an ArrayBuffer; a bare ArrayBuffer is accepted too. The engine holds a strong reference to it.
napi_value AttachSab(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value argv[1]; EngineBox* box = nullptr;
    gvt_engine* e = Eng(env, info, &argc, argv, &box);
    if (!e || argc < 1) { napi_throw_error(env, nullptr, "attach_sab(view)"); return nullptr; }
    void* data = nullptr; size_t bytes = 0; napi_typedarray_type t = napi_uint8_array;
    if (!ByteRange(env, argv[0], &data, &bytes, &t)) {
        napi_throw_error(env, nullptr, "attach_sab: pass a Float32Array view of the (Shared)ArrayBuffer");
        return nullptr;
    }
    bool is_ta = false; napi_is_typedarray(env, argv[0], &is_ta);
    if (is_ta && t != napi_float32_array) { napi_throw_error(env, nullptr, "attach_sab: Float32Array expected"); return nullptr; }
    if (bytes < (size_t)GVT_SAB_INTERNAL_F32 * 4 || ((uintptr_t)data & 3u)) { napi_throw_range_error(env, nullptr, "SAB view smaller than 2048 f32 or misaligned"); return nullptr; }
    size_t byte_offset = 0;
    if (is_ta) { napi_typedarray_type tt; size_t n; void* d; napi_value ab; napi_get_typedarray_info(env, argv[0], &tt, &n, &d, &ab, &byte_offset); }
    napi_ref ref = nullptr;
    NAPI_OK(napi_create_reference(env, argv[0], 1, &ref));
    if (gvt_engine_attach_sab(e, static_cast<float*>(data)) != GVT_OK) {
        napi_delete_reference(env, ref);
        napi_throw_error(env, nullptr, gvt_last_error());
        return nullptr;
    }
    if (box->sab_ref) napi_delete_reference(env, box->sab_ref);     // re-attach: release the previous view
    box->sab_ref = ref; box->sab_byte_offset = (uint32_t)byte_offset; box->attached = true;
    return Undefined(env);
}
// get_sab_ptr()                              lib.rs:116 -- the wasm build returns the byte offset of the engine's 2048-f32 block inside
// `memory.buffer`; physics.worker.ts:153-163 and physics-bridge.ts:107 divide it by 4 and index a Float32Array over that
// buffer. Here it is the byteOffset of the attached view inside ITS buffer -- addon/ts/index.ts attaches every engine to a
// region of the one SharedArrayBuffer its init() hands out as `memory.buffer`, so the unchanged worker reads live values.
napi_value GetSabPtr(napi_env env, napi_callback_info info) {
    size_t argc = 0; EngineBox* box = nullptr;
    gvt_engine* e = Eng(env, info, &argc, nullptr, &box);
    if (!e || !box->attached) { napi_throw_error(env, nullptr, "get_sab_ptr: no SAB view attached (addon/ts/index.ts attaches one per engine)"); return nullptr; }
    napi_value v; napi_create_uint32(env, box->sab_byte_offset, &v);
    return v;
}
napi_value GetSab(napi_env env, napi_callback_info info) {                       // a snapshot copy of the engine-owned block (debugging aid)
    size_t argc = 0;
    gvt_engine* e = Eng(env, info, &argc, nullptr);
    const float* p = nullptr; float* out = nullptr;
    GVT(gvt_engine_get_sab_ptr(e, &p));
    napi_value arr = F32Array(env, GVT_SAB_INTERNAL_F32, &out);
    if (arr) memcpy(out, p, GVT_SAB_INTERNAL_F32 * 4);
    return arr;
}
napi_value SetCameraState(napi_env env, napi_callback_info info) {               // lib.rs:120
    size_t argc = 6; napi_value argv[6];
    gvt_engine* e = Eng(env, info, &argc, argv);
    GVT(gvt_engine_set_camera_state(e, Num(env, argv[0]), Num(env, argv[1]), Num(env, argv[2]), 0, 0, 0));
    return Undefined(env);
}
napi_value SetAutoSpin(napi_env env, napi_callback_info info) {                  // lib.rs:124
    size_t argc = 1; napi_value argv[1]; bool b = false;
    gvt_engine* e = Eng(env, info, &argc, argv);
    napi_get_value_bool(env, argv[0], &b);
    GVT(gvt_engine_set_auto_spin(e, b ? 1 : 0));
    return Undefined(env);
}
#define ENGINE_GETTER(NAME, CALL)                                          \
    napi_value NAME(napi_env env, napi_callback_info info) {                \
        size_t argc = 2; napi_value argv[2];                                \
        gvt_engine* e = Eng(env, info, &argc, argv);           \
        double out = 0;                                                     \
        GVT(CALL);                                                          \
        napi_value v; napi_create_double(env, out, &v); return v;           \
    }
ENGINE_GETTER(ComputeHorizon, gvt_engine_compute_horizon(e, &out))               // lib.rs:85
ENGINE_GETTER(ComputeIsco, gvt_engine_compute_isco(e, &out))                     // lib.rs:89
ENGINE_GETTER(ComputePhotonSphere, gvt_engine_compute_photon_sphere(e, &out))    // lib.rs:93
ENGINE_GETTER(ComputeDilation, gvt_engine_compute_dilation(e, Num(env, argv[0]), &out))             // lib.rs:97
ENGINE_GETTER(ComputeGFactor, gvt_engine_compute_g_factor(e, Num(env, argv[0]), Num(env, argv[1]), &out))   // lib.rs:203
ENGINE_GETTER(ComputeShadowRadius, gvt_engine_compute_shadow_radius(e, &out))    // lib.rs:173
ENGINE_GETTER(ComputeDiskFlux, gvt_engine_compute_disk_flux(e, Num(env, argv[0]), &out))            // lib.rs:199
napi_value ShadowCurve(napi_env env, napi_callback_info info) {                  // lib.rs:161
    size_t argc = 2; napi_value argv[2];
    gvt_engine* e = Eng(env, info, &argc, argv);
    const double theta = Num(env, argv[0]);
    const uint32_t n_points = (uint32_t)Num(env, argv[1]);
    uint32_t n = 0;
    GVT(gvt_engine_compute_shadow_curve(e, theta, n_points, nullptr, 0, &n));
    float* out = nullptr;
    napi_value arr = F32Array(env, (size_t)n * 2, &out);
    if (arr) GVT(gvt_engine_compute_shadow_curve(e, theta, n_points, out, n, &n));
    return arr;
}
napi_value ShadowShift(napi_env env, napi_callback_info info) {                  // lib.rs:179
    size_t argc = 1; napi_value argv[1];
    gvt_engine* e = Eng(env, info, &argc, argv);
    float* out = nullptr;
    napi_value arr = F32Array(env, 2, &out);
    if (arr) GVT(gvt_engine_compute_shadow_shift(e, Num(env, argv[0]), out));
    return arr;
}
// ---- spacetime-visualisation helpers (lib.rs:139-305) ----
ENGINE_GETTER(ComputeKretschner, gvt_engine_compute_kretschner(e, Num(env, argv[0]), Num(env, argv[1]), &out))        // lib.rs:213
ENGINE_GETTER(ComputeLightConeTilt, gvt_engine_compute_light_cone_tilt(e, Num(env, argv[0]), Num(env, argv[1]), &out)) // lib.rs:238
ENGINE_GETTER(ComputeFrameDragOmega, gvt_engine_compute_frame_drag_omega(e, Num(env, argv[0]), Num(env, argv[1]), &out)) // lib.rs:267
ENGINE_GETTER(ComputeFlammHeight, gvt_engine_compute_flamm_height(e, Num(env, argv[0]), &out))                       // lib.rs:296
napi_value ComputeProperDistance(napi_env env, napi_callback_info info) {                                           // lib.rs:302
    size_t argc = 3; napi_value argv[3];
    gvt_engine* e = Eng(env, info, &argc, argv);
    double out = 0;
    GVT(gvt_engine_compute_proper_distance(e, Num(env, argv[0]), Num(env, argv[1]), (uint32_t)Num(env, argv[2]), &out));
    napi_value v; napi_create_double(env, out, &v); return v;
}
typedef int32_t (*FieldFn)(gvt_engine*, double, double, uint32_t, uint32_t, float*);
template <FieldFn FN> napi_value Field(napi_env env, napi_callback_info info) {   // (rMin, rMax, n1, n2) -> Float32Array(3 n1 n2)
    size_t argc = 4; napi_value argv[4];
    gvt_engine* e = Eng(env, info, &argc, argv);
    const uint32_t n1 = (uint32_t)Num(env, argv[2]), n2 = (uint32_t)Num(env, argv[3]);
    float* out = nullptr;
    napi_value arr = F32Array(env, (size_t)3 * n1 * n2, &out);
    if (arr) GVT(FN(e, Num(env, argv[0]), Num(env, argv[1]), n1, n2, out));
    return arr;
}
napi_value ErgosphereMesh(napi_env env, napi_callback_info info) {                                                  // lib.rs:153
    size_t argc = 2; napi_value argv[2];
    gvt_engine* e = Eng(env, info, &argc, argv);
    const uint32_t n1 = (uint32_t)Num(env, argv[0]), n2 = (uint32_t)Num(env, argv[1]);
    float* out = nullptr;
    napi_value arr = F32Array(env, (size_t)3 * n1 * n2, &out);
    if (arr) GVT(gvt_engine_generate_ergosphere_mesh(e, n1, n2, out));
    return arr;
}
napi_value DiskLut(napi_env env, napi_callback_info info) {                      // lib.rs:107
    size_t argc = 0;
    gvt_engine* e = Eng(env, info, &argc, nullptr);
    float* out = nullptr;
    napi_value arr = F32Array(env, 512, &out);
    if (arr) GVT(gvt_engine_generate_disk_lut(e, out));
    return arr;
}
napi_value DiskLutPtr(napi_env env, napi_callback_info info) {                   // lib.rs:112: a copy of the engine-owned LUT
    size_t argc = 0;
    gvt_engine* e = Eng(env, info, &argc, nullptr);
    const float* p = nullptr; uint32_t n = 0;
    GVT(gvt_engine_get_disk_lut_ptr(e, &p, &n));
    float* out = nullptr;
    napi_value arr = F32Array(env, n, &out);
    if (arr && n) memcpy(out, p, (size_t)n * sizeof(float));
    return arr;
}
napi_value SpectrumLut(napi_env env, napi_callback_info info) {                  // lib.rs:128
    size_t argc = 3; napi_value argv[3];
    gvt_engine* e = Eng(env, info, &argc, argv);
    const uint32_t w = (uint32_t)Num(env, argv[0]), h = (uint32_t)Num(env, argv[1]);
    float* out = nullptr;
    napi_value arr = F32Array(env, (size_t)w * h * 4, &out);
    if (arr) GVT(gvt_engine_generate_spectrum_lut(e, w, h, Num(env, argv[2]), out));
    return arr;
}
napi_value IntegrateRay(napi_env env, napi_callback_info info) {                 // lib.rs:422
    size_t argc = 4; napi_value argv[4];
    gvt_engine* e = Eng(env, info, &argc, argv);
    uint32_t n = 0; napi_get_array_length(env, argv[0], &n);
    if (n < 8) return argv[0];                                                   // lib.rs:429-431: returned unchanged
    double in8[8], out8[8]; bool ks = false;
    for (uint32_t i = 0; i < 8; i++) { napi_value v; napi_get_element(env, argv[0], i, &v); in8[i] = Num(env, v); }
    napi_get_value_bool(env, argv[3], &ks);
    GVT(gvt_engine_integrate_ray(e, in8, (uint64_t)Num(env, argv[1]), Num(env, argv[2]), ks ? 1 : 0, out8, nullptr, nullptr, nullptr));
    void* data; napi_value ab, out;
    NAPI_OK(napi_create_arraybuffer(env, 64, &data, &ab));
    memcpy(data, out8, 64);
    NAPI_OK(napi_create_typedarray(env, napi_float64_array, 8, ab, 0, &out));
    return out;
}
napi_value SabLayout(napi_env env, napi_callback_info info) {                    // lib.rs:411
    size_t argc = 0;
    gvt_engine* e = Eng(env, info, &argc, nullptr);
    uint32_t l[5];
    GVT(gvt_engine_get_sab_layout(e, l));
    void* data; napi_value ab, out;
    NAPI_OK(napi_create_arraybuffer(env, 20, &data, &ab));
    memcpy(data, l, 20);
    NAPI_OK(napi_create_typedarray(env, napi_uint32_array, 5, ab, 0, &out));
    return out;
}

// ---------------------------------------------------------------- KerrRenderer (Seam B)
napi_value RendererNew(napi_env env, napi_callback_info info) {                  // new KerrRenderer(device = 0)
    size_t argc = 1; napi_value argv[1], self;
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, &self, nullptr));
    GvtDeviceConfig cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.struct_size = sizeof(cfg); cfg.device = argc > 0 ? (int32_t)Num(env, argv[0]) : 0; cfg.rank = 0; cfg.world_size = 1;
    gvt_renderer* r = nullptr;
    GVT(gvt_render_create(&cfg, &r));
    NAPI_OK(napi_wrap(env, self, r, [](napi_env, void* d, void*) { gvt_render_destroy(static_cast<gvt_renderer*>(d)); }, nullptr, nullptr));
    return self;
}
napi_value InitLuts(napi_env env, napi_callback_info info) {                     // (mass, spin, w, h, maxTemp)  spectral.ts:21-61
    size_t argc = 5; napi_value argv[5];
    gvt_renderer* r = Self<gvt_renderer>(env, info, &argc, argv);
    GVT(gvt_render_init_luts(r, Num(env, argv[0]), Num(env, argv[1]), (uint32_t)Num(env, argv[2]), (uint32_t)Num(env, argv[3]), Num(env, argv[4])));
    return Undefined(env);
}
napi_value Resize(napi_env env, napi_callback_info info) {                       // renderer.ts:269
    size_t argc = 2; napi_value argv[2];
    gvt_renderer* r = Self<gvt_renderer>(env, info, &argc, argv);
    GVT(gvt_render_resize(r, (uint32_t)Num(env, argv[0]), (uint32_t)Num(env, argv[1])));
    return Undefined(env);
}
napi_value SetFrameFormat(napi_env env, napi_callback_info info) {               // (format: 0 = RGBA32F, 1 = RGBA16F)  reprojection.ts:120-140
    size_t argc = 1; napi_value argv[1];
    gvt_renderer* r = Self<gvt_renderer>(env, info, &argc, argv);
    GVT(gvt_render_set_frame_format(r, (uint32_t)Num(env, argv[0])));
    return Undefined(env);
}
uint32_t OptU32(napi_env env, napi_value obj, const char* key, uint32_t dflt) {
    bool has = false; napi_value v;
    if (napi_has_named_property(env, obj, key, &has) != napi_ok || !has) return dflt;
    napi_get_named_property(env, obj, key, &v);
    return (uint32_t)Num(env, v);
}
// A frame target in DEVICE memory: importExternalFd(fd, bytes, device = 0, dedicated = false) maps memory
// the presenter exported from its graphics API (VK_KHR_external_memory_fd / GL_EXT_memory_object_fd) and returns an opaque
// value that renderFrame / renderFragment / bloom accept in place of the output ArrayBuffer: the frame never crosses host
// memory. The mapping is released when the value is collected.
struct ExternalTarget { gvt_external_buffer* h; void* ptr; uint64_t bytes; };
napi_value ImportExternalFd(napi_env env, napi_callback_info info) {
    size_t argc = 4; napi_value argv[4];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    if (argc < 2) { napi_throw_error(env, nullptr, "importExternalFd(fd, bytes, device = 0, dedicated = false)"); return nullptr; }
    bool dedicated = false;
    if (argc > 3) napi_get_value_bool(env, argv[3], &dedicated);
    ExternalTarget* t = new ExternalTarget{nullptr, nullptr, (uint64_t)Num(env, argv[1])};
    if (gvt_external_import_fd(argc > 2 ? (int32_t)Num(env, argv[2]) : 0, (int32_t)Num(env, argv[0]), t->bytes, dedicated ? 1 : 0, &t->h, &t->ptr) != GVT_OK) {
        delete t; napi_throw_error(env, nullptr, gvt_last_error()); return nullptr;
    }
    napi_value out;
    NAPI_OK(napi_create_external(env, t, [](napi_env, void* d, void*) { ExternalTarget* x = static_cast<ExternalTarget*>(d); gvt_external_release(x->h); delete x; }, nullptr, &out));
    return out;
}
// the output argument of the frame calls: an ArrayBuffer / typed array (host memory) or an imported external target
bool OutRange(napi_env env, napi_value v, void** data, size_t* bytes) {
    napi_valuetype vt;
    if (napi_typeof(env, v, &vt) == napi_ok && vt == napi_external) {
        void* p = nullptr;
        if (napi_get_value_external(env, v, &p) != napi_ok || !p) return false;
        *data = static_cast<ExternalTarget*>(p)->ptr; *bytes = (size_t)static_cast<ExternalTarget*>(p)->bytes;
        return true;
    }
    return ByteRange(env, v, data, bytes);
}
// renderFrame(cameraF32x88, physF32x8, {maxSteps, method, precision, taa, jitter, f16}, outArrayBuffer | externalTarget) -> stats   renderer.ts:280
napi_value RenderFrame(napi_env env, napi_callback_info info) {
    size_t argc = 4; napi_value argv[4];
    gvt_renderer* r = Self<gvt_renderer>(env, info, &argc, argv);
    napi_typedarray_type t; void *cam, *phys, *out; size_t nbytes, outlen;
    if (!r || argc < 4) { napi_throw_error(env, nullptr, "renderFrame(camera, physics, options, out)"); return nullptr; }
    if (!ByteRange(env, argv[0], &cam, &nbytes, &t) || t != napi_float32_array || nbytes < sizeof(GvtCamera)) { napi_throw_range_error(env, nullptr, "camera: Float32Array(88) expected"); return nullptr; }
    // PhysicsParams is 6 f32 + 2 u32 (types/webgpu.ts:42-64): any 32-bit view of >= 32 bytes
    if (!ByteRange(env, argv[1], &phys, &nbytes, &t) || (t != napi_float32_array && t != napi_uint32_array && t != napi_int32_array) || nbytes < sizeof(GvtPhysicsParams)) { napi_throw_range_error(env, nullptr, "physics: 32-bit typed array of 32 bytes expected"); return nullptr; }
    if (!OutRange(env, argv[3], &out, &outlen)) { napi_throw_error(env, nullptr, "out: ArrayBuffer, typed array or external target expected"); return nullptr; }
    GvtRenderParams p; gvt_render_params_default(&p);
    p.max_steps = OptU32(env, argv[2], "maxSteps", p.max_steps);
    p.method = OptU32(env, argv[2], "method", p.method);
    p.precision = OptU32(env, argv[2], "precision", p.precision);
    if (OptU32(env, argv[2], "taa", 0)) p.flags |= GVT_FLAG_TAA;
    if (OptU32(env, argv[2], "jitter", 0)) p.flags |= GVT_FLAG_JITTER;
    if (OptU32(env, argv[2], "f16", 0)) p.output_format = GVT_FORMAT_RGBA16F;
    const GvtPhysicsParams* ph = static_cast<const GvtPhysicsParams*>(phys);
    const size_t need = (size_t)ph->resolution[0] * (size_t)ph->resolution[1] * (p.output_format == GVT_FORMAT_RGBA16F ? 8 : 16);
    if (outlen < need) { napi_throw_range_error(env, nullptr, "output buffer too small"); return nullptr; }
    GvtFrameStats st;
    GVT(gvt_render_frame(r, static_cast<const GvtCamera*>(cam), ph, &p, out, &st));
    napi_value o, v;
    NAPI_OK(napi_create_object(env, &o));
    napi_create_double(env, st.total_ms, &v); napi_set_named_property(env, o, "totalMs", v);
    napi_create_double(env, st.trace_ms, &v); napi_set_named_property(env, o, "traceMs", v);
    napi_create_double(env, (double)st.steps_committed, &v); napi_set_named_property(env, o, "steps", v);
    return o;
}

// setNoiseTextures(noiseU8x262144, blueU8x262144): the two 256x256 RGBA8 textures of webgl-utils.ts:259-303
napi_value SetNoiseTextures(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value argv[2];
    gvt_renderer* r = Self<gvt_renderer>(env, info, &argc, argv);
    napi_typedarray_type t0, t1; size_t n0 = 0, n1 = 0; void *a, *b;
    if (!r || !ByteRange(env, argv[0], &a, &n0, &t0) || !ByteRange(env, argv[1], &b, &n1, &t1) ||
        (t0 != napi_uint8_array && t0 != napi_uint8_clamped_array) || (t1 != napi_uint8_array && t1 != napi_uint8_clamped_array) ||
        n0 < 262144 || n1 < 262144) { napi_throw_range_error(env, nullptr, "Uint8Array(256*256*4) expected"); return nullptr; }
    GVT(gvt_render_set_noise_textures(r, static_cast<const uint8_t*>(a), static_cast<const uint8_t*>(b), 256));
    return Undefined(env);
}
// renderFragment(uniformsF32x155, {precision, flags, format, taaBlend, cameraMoving}, outArrayBuffer) -> stats
// uniforms: the GvtGlslUniforms block as 155 32-bit words (struct_size, features and max_ray_steps are integer words:
// build it with a DataView / Uint32Array alias, see addon/ts/webgl_b200_renderer.ts)      webgl/renderer.ts:173
napi_value RenderFragment(napi_env env, napi_callback_info info) {
    size_t argc = 3; napi_value argv[3];
    gvt_renderer* r = Self<gvt_renderer>(env, info, &argc, argv);
    napi_typedarray_type t; size_t nbytes; void *u, *out; size_t outlen;
    if (!r || !ByteRange(env, argv[0], &u, &nbytes, &t) || (t != napi_float32_array && t != napi_uint32_array && t != napi_int32_array) ||
        nbytes < sizeof(GvtGlslUniforms)) { napi_throw_range_error(env, nullptr, "uniforms: 155 32-bit words expected"); return nullptr; }
    if (!OutRange(env, argv[2], &out, &outlen)) { napi_throw_error(env, nullptr, "out: ArrayBuffer, typed array or external target expected"); return nullptr; }
    const uint32_t precision = OptU32(env, argv[1], "precision", GVT_PRECISION_F32_FAST);
    const uint32_t flags = OptU32(env, argv[1], "flags", 0), format = OptU32(env, argv[1], "format", GVT_FORMAT_RGBA32F);
    const uint32_t moving = OptU32(env, argv[1], "cameraMoving", 0);
    const GvtGlslUniforms* gu = static_cast<const GvtGlslUniforms*>(u);
    const size_t need = (size_t)gu->resolution[0] * (size_t)gu->resolution[1] * (format == GVT_FORMAT_RGBA32F ? 16 : format == GVT_FORMAT_RGBA16F ? 8 : 4);
    if (outlen == 0) out = nullptr;                                               // frame stays on the device (bloom follows)
    else if (outlen < need) { napi_throw_range_error(env, nullptr, "output ArrayBuffer too small"); return nullptr; }
    GvtFrameStats st;
    GVT(gvt_render_fragment_glsl(r, gu, precision, flags, format, 0.75f, moving, out, &st));
    napi_value o, v;
    NAPI_OK(napi_create_object(env, &o));
    napi_create_double(env, st.total_ms, &v); napi_set_named_property(env, o, "totalMs", v);
    napi_create_double(env, st.trace_ms, &v); napi_set_named_property(env, o, "shaderMs", v);
    napi_create_double(env, (double)st.steps_committed, &v); napi_set_named_property(env, o, "steps", v);
    return o;
}

// bloom({enabled, intensity, threshold, blurPasses, format}, outArrayBuffer) -> ms      rendering/bloom.ts:446-632
napi_value Bloom(napi_env env, napi_callback_info info) {
    size_t argc = 2; napi_value argv[2];
    gvt_renderer* r = Self<gvt_renderer>(env, info, &argc, argv);
    GvtBloomConfig cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.struct_size = sizeof(cfg);
    cfg.enabled = OptU32(env, argv[0], "enabled", 1); cfg.blur_passes = OptU32(env, argv[0], "blurPasses", 2);
    cfg.intensity = 0.5f; cfg.threshold = 0.8f;                                   // bloom.ts:32-39
    bool has = false; napi_value v;
    if (napi_has_named_property(env, argv[0], "intensity", &has) == napi_ok && has) { napi_get_named_property(env, argv[0], "intensity", &v); cfg.intensity = (float)Num(env, v); }
    if (napi_has_named_property(env, argv[0], "threshold", &has) == napi_ok && has) { napi_get_named_property(env, argv[0], "threshold", &v); cfg.threshold = (float)Num(env, v); }
    cfg.precise = OptU32(env, argv[0], "precise", 0);
    void* out; size_t outlen;
    if (!r || !OutRange(env, argv[1], &out, &outlen)) { napi_throw_error(env, nullptr, "bloom(options, out: ArrayBuffer | typed array | external target)"); return nullptr; }
    const uint32_t format = OptU32(env, argv[0], "format", GVT_FORMAT_RGBA8_UNORM);
    double ms = 0;
    if (outlen == 0) out = nullptr;                                               // result stays on the device
    else {
        // the library writes width*height pixels of `format`: the caller's buffer must hold them (it knows the frame size)
        uint32_t w = 0, h = 0;
        GVT(gvt_render_get_size(r, &w, &h));
        const size_t need = (size_t)w * h * (format == GVT_FORMAT_RGBA32F ? 16 : format == GVT_FORMAT_RGBA16F ? 8 : 4);
        if (outlen < need) { napi_throw_range_error(env, nullptr, "bloom: output buffer smaller than width*height*bytes-per-pixel"); return nullptr; }
    }
    GVT(gvt_render_bloom(r, &cfg, format, out, &ms));
    napi_value o; napi_create_double(env, ms, &o);
    return o;
}

}  // namespace

NAPI_MODULE_INIT() {
    const napi_property_descriptor engine[] = {
        {"update_params", 0, UpdateParams, 0, 0, 0, napi_default, 0}, {"tick_sab", 0, TickSab, 0, 0, 0, napi_default, 0},
        {"attach_sab", 0, AttachSab, 0, 0, 0, napi_default, 0}, {"get_sab_ptr", 0, GetSabPtr, 0, 0, 0, napi_default, 0},
        {"get_sab", 0, GetSab, 0, 0, 0, napi_default, 0},
        {"get_sab_layout", 0, SabLayout, 0, 0, 0, napi_default, 0},
        {"set_camera_state", 0, SetCameraState, 0, 0, 0, napi_default, 0}, {"set_auto_spin", 0, SetAutoSpin, 0, 0, 0, napi_default, 0},
        {"compute_horizon", 0, ComputeHorizon, 0, 0, 0, napi_default, 0}, {"compute_isco", 0, ComputeIsco, 0, 0, 0, napi_default, 0},
        {"compute_photon_sphere", 0, ComputePhotonSphere, 0, 0, 0, napi_default, 0},
        {"compute_dilation", 0, ComputeDilation, 0, 0, 0, napi_default, 0}, {"compute_g_factor", 0, ComputeGFactor, 0, 0, 0, napi_default, 0},
        {"compute_shadow_curve", 0, ShadowCurve, 0, 0, 0, napi_default, 0}, {"compute_shadow_shift", 0, ShadowShift, 0, 0, 0, napi_default, 0},
        {"compute_shadow_radius", 0, ComputeShadowRadius, 0, 0, 0, napi_default, 0}, {"compute_disk_flux", 0, ComputeDiskFlux, 0, 0, 0, napi_default, 0},
        {"compute_kretschner", 0, ComputeKretschner, 0, 0, 0, napi_default, 0}, {"compute_light_cone_tilt", 0, ComputeLightConeTilt, 0, 0, 0, napi_default, 0},
        {"compute_frame_drag_omega", 0, ComputeFrameDragOmega, 0, 0, 0, napi_default, 0}, {"compute_flamm_height", 0, ComputeFlammHeight, 0, 0, 0, napi_default, 0},
        {"compute_proper_distance", 0, ComputeProperDistance, 0, 0, 0, napi_default, 0},
        {"generate_curvature_field", 0, Field<gvt_engine_generate_curvature_field>, 0, 0, 0, napi_default, 0},
        {"generate_tilt_field", 0, Field<gvt_engine_generate_tilt_field>, 0, 0, 0, napi_default, 0},
        {"generate_frame_drag_field", 0, Field<gvt_engine_generate_frame_drag_field>, 0, 0, 0, napi_default, 0},
        {"generate_embedding_mesh", 0, Field<gvt_engine_generate_embedding_mesh>, 0, 0, 0, napi_default, 0},
        {"generate_ergosphere_mesh", 0, ErgosphereMesh, 0, 0, 0, napi_default, 0},
        {"generate_disk_lut", 0, DiskLut, 0, 0, 0, napi_default, 0}, {"get_disk_lut_ptr", 0, DiskLutPtr, 0, 0, 0, napi_default, 0}, {"generate_spectrum_lut", 0, SpectrumLut, 0, 0, 0, napi_default, 0},
        {"integrate_ray_relativistic", 0, IntegrateRay, 0, 0, 0, napi_default, 0}};
    const napi_property_descriptor renderer[] = {
        {"initLuts", 0, InitLuts, 0, 0, 0, napi_default, 0}, {"resize", 0, Resize, 0, 0, 0, napi_default, 0},
        {"setFrameFormat", 0, SetFrameFormat, 0, 0, 0, napi_default, 0},
        {"renderFrame", 0, RenderFrame, 0, 0, 0, napi_default, 0},
        {"setNoiseTextures", 0, SetNoiseTextures, 0, 0, 0, napi_default, 0}, {"renderFragment", 0, RenderFragment, 0, 0, 0, napi_default, 0},
        {"bloom", 0, Bloom, 0, 0, 0, napi_default, 0}};
    napi_value cls;
    NAPI_OK(napi_define_class(env, "PhysicsEngine", NAPI_AUTO_LENGTH, EngineNew, nullptr, sizeof(engine) / sizeof(engine[0]), engine, &cls));
    NAPI_OK(napi_set_named_property(env, exports, "PhysicsEngine", cls));
    NAPI_OK(napi_define_class(env, "KerrRenderer", NAPI_AUTO_LENGTH, RendererNew, nullptr, sizeof(renderer) / sizeof(renderer[0]), renderer, &cls));
    NAPI_OK(napi_set_named_property(env, exports, "KerrRenderer", cls));
    napi_value fn;
    NAPI_OK(napi_create_function(env, "importExternalFd", NAPI_AUTO_LENGTH, ImportExternalFd, nullptr, &fn));
    NAPI_OK(napi_set_named_property(env, exports, "importExternalFd", fn));
    return exports;
}
