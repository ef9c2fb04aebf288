"""ctypes binding of include/gravitas_b200.h. Fails loudly if the CUDA library has not been built."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GRAVITAS_B200_LIB selects another build of the SAME library (kernel tuning experiments); never a fallback.
_SO = os.environ.get("GRAVITAS_B200_LIB") or os.path.join(_HERE, "libgravitas_b200.so")
_LIB = None

# SAB v2 offsets, f32 element indices (physics-bridge.ts:5-11; lib.rs:36-40)
OFFSETS = {"CONTROL": 0, "CAMERA": 64, "PHYSICS": 128, "TELEMETRY": 256, "LUTS": 2048}

GVT_OK, GVT_ERR_INVALID, GVT_ERR_NO_DEVICE, GVT_ERR_CUDA, GVT_ERR_NCCL, GVT_ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
TERM_NONE, TERM_HORIZON, TERM_ESCAPE, TERM_MAXSTEPS, TERM_DISK = 0, 1, 2, 3, 4
COORDS_BL, COORDS_KS = 0, 1
METHOD_RKF45, METHOD_RK4, METHOD_SYMPLECTIC, METHOD_VERLET_GLSL = 0, 1, 2, 3
PRECISION_F64, PRECISION_F32, PRECISION_F32_FAST, PRECISION_MIXED = 0, 1, 2, 3
FORMAT_RGBA32F, FORMAT_RGBA16F, FORMAT_RGBA8_REINHARD, FORMAT_RGBA8_ACES, FORMAT_RGBA8_UNORM = 0, 1, 2, 3, 4
FORMAT_BYTES = {0: 16, 1: 8, 2: 4, 3: 4, 4: 4}
FORMAT_DTYPE = {0: "float32", 1: "float16", 2: "uint8", 3: "uint8", 4: "uint8"}
STEP_CONSTANT, STEP_WGSL = 0, 1
FLAG_JITTER, FLAG_BUDGET, FLAG_TAA, FLAG_NO_GATHER, FLAG_D2H_OWN_ROWS, FLAG_PEER_STORE, FLAG_TAA_WEBGL = 1, 2, 8, 16, 32, 64, 128
FLAG_ROW_INTERLEAVE = 256
FLAG_DEBUG_COUNTS = 512
FLAG_TAA_PRECISE = 1024


class GravitasError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gravitas_b200 error {code}: {msg}")
        self.code = code


class GvtCamera(C.Structure):  # 352 bytes, types/webgpu.ts:67-116
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16), ("inv_view", C.c_float * 16),
                ("inv_proj", C.c_float * 16), ("prev_view_proj", C.c_float * 16), ("position", C.c_float * 4),
                ("direction", C.c_float * 4)]


class GvtPhysicsParams(C.Structure):  # 32 bytes, types/webgpu.ts:42-64
    _fields_ = [("mass", C.c_float), ("spin", C.c_float), ("resolution", C.c_float * 2), ("time", C.c_float),
                ("dt", C.c_float), ("frame_index", C.c_uint32), ("_pad", C.c_uint32)]


class GvtRenderParams(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("method", C.c_uint32), ("precision", C.c_uint32), ("coords", C.c_uint32),
                ("step_rule", C.c_uint32), ("max_steps", C.c_uint32), ("renormalize_interval", C.c_uint32),
                ("flags", C.c_uint32), ("output_format", C.c_uint32), ("_pad", C.c_uint32), ("tolerance", C.c_double),
                ("initial_step", C.c_double), ("escape_radius", C.c_double), ("disk_r_out", C.c_double),
                ("taa_blend", C.c_float), ("taa_camera_moving", C.c_uint32)]


GLSL_LENSING, GLSL_DISK, GLSL_JETS, GLSL_STARS, GLSL_PHOTON_GLOW, GLSL_DOPPLER, GLSL_REDSHIFT = 1, 2, 4, 8, 16, 32, 64
GLSL_LINEAR_OUTPUT, GLSL_QUALITY_LOW = 128, 256


class GvtGlslUniforms(C.Structure):  # chunks/common.ts:9-38 uniforms + shader-manager #defines (manager.ts:55-82)
    _fields_ = [("struct_size", C.c_uint32), ("features", C.c_uint32), ("resolution", C.c_float * 2), ("time", C.c_float),
                ("mass", C.c_float), ("spin", C.c_float), ("disk_density", C.c_float), ("disk_temp", C.c_float),
                ("mouse", C.c_float * 2), ("zoom", C.c_float), ("lensing_strength", C.c_float), ("disk_size", C.c_float),
                ("disk_scale_height", C.c_float), ("max_ray_steps", C.c_int32), ("debug", C.c_float),
                ("show_redshift", C.c_float), ("show_kerr_shadow", C.c_float), ("shadow_count", C.c_float),
                ("cam_pos", C.c_float * 3), ("cam_quat", C.c_float * 4), ("shadow_curve", C.c_float * 128)]


class GvtBloomConfig(C.Structure):  # rendering/bloom.ts:22-39
    _fields_ = [("struct_size", C.c_uint32), ("enabled", C.c_uint32), ("intensity", C.c_float), ("threshold", C.c_float),
                ("blur_passes", C.c_uint32), ("precise", C.c_uint32)]


class GvtDeviceConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("rank", C.c_int32), ("world_size", C.c_int32),
                ("nccl_id", C.c_uint8 * 128)]


class GvtFrameStats(C.Structure):
    _fields_ = [("trace_ms", C.c_double), ("taa_ms", C.c_double), ("gather_ms", C.c_double), ("total_ms", C.c_double),
                ("steps_committed", C.c_uint64), ("steps_executed", C.c_uint64), ("rhs_evals", C.c_uint64),
                ("n_horizon", C.c_uint64), ("n_escape", C.c_uint64), ("n_maxsteps", C.c_uint64), ("n_disk", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("kernel_launches", C.c_uint32),
                ("rows_begin", C.c_uint32), ("rows_end", C.c_uint32)]


assert C.sizeof(GvtCamera) == 352 and C.sizeof(GvtPhysicsParams) == 32 and C.sizeof(GvtGlslUniforms) == 620

_d, _i32, _u32, _u64, _vp = C.c_double, C.c_int32, C.c_uint32, C.c_uint64, C.c_void_p
_pd, _pf, _pu32, _pu64 = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)

# every symbol include/gravitas_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "gvt_last_error": (C.c_char_p, []),
    "gvt_build_info": (C.c_char_p, []),
    "gvt_abi_version": (_i32, []),
    "gvt_device_count": (_i32, [C.POINTER(_i32)]),
    "gvt_engine_create": (_i32, [_d, _d, C.POINTER(_vp)]),
    "gvt_engine_destroy": (_i32, [_vp]),
    "gvt_engine_update_params": (_i32, [_vp, _d, _d]),
    "gvt_engine_attach_sab": (_i32, [_vp, _pf]),
    "gvt_engine_get_sab_ptr": (_i32, [_vp, C.POINTER(_pf)]),
    "gvt_engine_get_sab_layout": (_i32, [_vp, _pu32]),
    "gvt_engine_tick_sab": (_i32, [_vp, _d]),
    "gvt_engine_set_camera_state": (_i32, [_vp, _d, _d, _d, _d, _d, _d]),
    "gvt_engine_set_auto_spin": (_i32, [_vp, _i32]),
    "gvt_engine_compute_horizon": (_i32, [_vp, _pd]),
    "gvt_engine_compute_isco": (_i32, [_vp, _pd]),
    "gvt_engine_compute_photon_sphere": (_i32, [_vp, _pd]),
    "gvt_engine_compute_dilation": (_i32, [_vp, _d, _pd]),
    "gvt_engine_compute_g_factor": (_i32, [_vp, _d, _d, _pd]),
    "gvt_engine_compute_shadow_curve": (_i32, [_vp, _d, _u32, _pf, _u32, _pu32]),
    "gvt_engine_compute_shadow_radius": (_i32, [_vp, _pd]),
    "gvt_engine_compute_shadow_shift": (_i32, [_vp, _d, _pf]),
    "gvt_engine_compute_disk_flux": (_i32, [_vp, _d, _pd]),
    "gvt_engine_compute_kretschner": (_i32, [_vp, _d, _d, _pd]),
    "gvt_engine_generate_curvature_field": (_i32, [_vp, _d, _d, _u32, _u32, _pf]),
    "gvt_engine_compute_light_cone_tilt": (_i32, [_vp, _d, _d, _pd]),
    "gvt_engine_generate_tilt_field": (_i32, [_vp, _d, _d, _u32, _u32, _pf]),
    "gvt_engine_compute_frame_drag_omega": (_i32, [_vp, _d, _d, _pd]),
    "gvt_engine_generate_frame_drag_field": (_i32, [_vp, _d, _d, _u32, _u32, _pf]),
    "gvt_engine_compute_flamm_height": (_i32, [_vp, _d, _pd]),
    "gvt_engine_compute_proper_distance": (_i32, [_vp, _d, _d, _u32, _pd]),
    "gvt_engine_generate_embedding_mesh": (_i32, [_vp, _d, _d, _u32, _u32, _pf]),
    "gvt_engine_generate_ergosphere_mesh": (_i32, [_vp, _u32, _u32, _pf]),
    "gvt_engine_generate_disk_lut": (_i32, [_vp, _pf]),
    "gvt_engine_get_disk_lut_ptr": (_i32, [_vp, C.POINTER(_pf), _pu32]),
    "gvt_engine_generate_spectrum_lut": (_i32, [_vp, _u32, _u32, _d, _pf]),
    "gvt_engine_integrate_ray": (_i32, [_vp, _pd, _u64, _d, _i32, _pd, _pu32, _pu64, _pd]),
    "gvt_engine_integrate_rays": (_i32, [_vp, C.POINTER(GvtRenderParams), _u64, _pd, _pd, _pu32, _pu32, _pd, _pu32]),
    "gvt_nccl_unique_id": (_i32, [C.POINTER(C.c_uint8)]),
    "gvt_render_params_default": (_i32, [C.POINTER(GvtRenderParams)]),
    "gvt_render_create": (_i32, [C.POINTER(GvtDeviceConfig), C.POINTER(_vp)]),
    "gvt_render_destroy": (_i32, [_vp]),
    "gvt_render_init_luts": (_i32, [_vp, _d, _d, _u32, _u32, _d]),
    "gvt_render_set_luts": (_i32, [_vp, _pf, _u32, _u32, _pf, _u32, _d, _d]),
    "gvt_render_resize": (_i32, [_vp, _u32, _u32]),
    "gvt_render_rows": (_i32, [_vp, C.POINTER(GvtCamera), C.POINTER(GvtPhysicsParams), C.POINTER(GvtRenderParams), _u32, _u32, _vp,
                        C.POINTER(GvtFrameStats)]),
    "gvt_render_frame": (_i32, [_vp, C.POINTER(GvtCamera), C.POINTER(GvtPhysicsParams), C.POINTER(GvtRenderParams), _vp,
                                C.POINTER(GvtFrameStats)]),
    "gvt_render_read_frame": (_i32, [_vp, _u32, _vp]),
    "gvt_render_get_size": (_i32, [_vp, C.POINTER(_u32), C.POINTER(_u32)]),
    "gvt_trace_states": (_i32, [_vp, C.POINTER(GvtCamera), C.POINTER(GvtPhysicsParams), C.POINTER(GvtRenderParams), _u32,
                                _u32, _u32, _u32, _u32, _pd, _pu32, _pu32, _pd, _pd]),
    "gvt_taa_resolve": (_i32, [_vp, C.POINTER(GvtCamera), _u32, _u32, _pf, _pf, _pf]),
    "gvt_taa_resolve_webgl": (_i32, [_vp, _u32, _u32, _pf, _pf, C.c_float, _i32, _pf]),
    "gvt_taa_resolve_ex": (_i32, [_vp, C.POINTER(GvtCamera), _u32, _u32, _vp, _vp, _vp, _u32, _u32, C.c_float, _i32, _i32, C.POINTER(C.c_double)]),
    "gvt_render_set_frame_format": (_i32, [_vp, _u32]),
    "gvt_render_reset_history": (_i32, [_vp]),
    "gvt_render_set_noise_textures": (_i32, [_vp, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), _u32]),
    "gvt_render_fragment_glsl": (_i32, [_vp, C.POINTER(GvtGlslUniforms), _u32, _u32, _u32, C.c_float, _u32, _vp,
                                        C.POINTER(GvtFrameStats)]),
    "gvt_render_fragment_glsl_debug": (_i32, [_vp, _pu32, _pu32]),
    "gvt_render_bloom": (_i32, [_vp, C.POINTER(GvtBloomConfig), _u32, _vp, _pd]),
    "gvt_render_export_frames": (_i32, [_vp, C.POINTER(C.c_uint8)]),
    "gvt_render_import_peer_frames": (_i32, [_vp, _i32, C.POINTER(C.c_uint8)]),
    "gvt_host_alloc": (_i32, [C.c_size_t, C.POINTER(_vp)]),
    "gvt_host_free": (_i32, [_vp]),
    "gvt_host_register": (_i32, [_vp, C.c_size_t]),
    "gvt_host_unregister": (_i32, [_vp]),
    "gvt_external_import_fd": (_i32, [_i32, _i32, C.c_uint64, _i32, C.POINTER(_vp), C.POINTER(_vp)]),
    "gvt_external_release": (_i32, [_vp]),
    "gvt_external_semaphore_import_fd": (_i32, [_i32, _i32, _i32, C.POINTER(_vp)]),
    "gvt_external_semaphore_release": (_i32, [_vp]),
    "gvt_render_wait_external": (_i32, [_vp, _vp, C.c_uint64]),
    "gvt_render_signal_external": (_i32, [_vp, _vp, C.c_uint64]),
    "gvt_measure_fma_peak": (_i32, [_vp, _i32, _pd, _pd]),
    "gvt_device_info": (_i32, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.c_char_p]),
}


def lib_path():
    return _SO


def lib():
    """Load libgravitas_b200.so. There is no fallback implementation: a missing library is an error."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(_SO):
            raise GravitasError(GVT_ERR_NO_DEVICE,
                                f"{_SO} not built — run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(nvcc, sm_100a). This package has no CPU or PyTorch fallback.")
        L = C.CDLL(_SO)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        _LIB = L
    return _LIB


def check(rc):
    if rc != GVT_OK:
        msg = lib().gvt_last_error()
        raise GravitasError(rc, msg.decode() if msg else "unknown")
