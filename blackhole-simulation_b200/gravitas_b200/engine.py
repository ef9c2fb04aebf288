"""PhysicsEngine — mirror of the wasm-bindgen class in physics-engine/gravitas-wasm/src/lib.rs:42-465."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib


def init_hooks():  # lib.rs:30-33 installs the wasm panic hook; errors here are status codes -> GravitasError
    return None


class PhysicsEngine:
    """Same surface as the reference's JS-visible ``PhysicsEngine`` (method names and argument meaning kept).

    ``get_sab_ptr`` returns a numpy f32 view of the engine-owned 2048-float buffer (the wasm version returns a
    byte offset into ``memory.buffer`` that JS wraps in a Float32Array — physics.worker.ts:153-159); ``attach_sab``
    takes any writable f32 buffer of >= 2048 elements (a SharedArrayBuffer view in the reference)."""

    def __init__(self, mass, spin):  # lib.rs:59-72
        self._h = C.c_void_p()
        check(lib().gvt_engine_create(float(mass), float(spin), C.byref(self._h)))
        self._sab_keepalive = None

    def free(self):  # wasm-bindgen .free()
        if self._h:
            lib().gvt_engine_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = free

    def update_params(self, mass, spin):  # lib.rs:78-83
        check(lib().gvt_engine_update_params(self._h, float(mass), float(spin)))

    def _scalar(self, fn, *args):
        out = C.c_double()
        check(fn(self._h, *args, C.byref(out)))
        return out.value

    def compute_horizon(self):  # lib.rs:85-87
        return self._scalar(lib().gvt_engine_compute_horizon)

    def compute_isco(self):  # lib.rs:89-91
        return self._scalar(lib().gvt_engine_compute_isco)

    def compute_photon_sphere(self):  # lib.rs:93-95
        return self._scalar(lib().gvt_engine_compute_photon_sphere)

    def compute_dilation(self, r):  # lib.rs:97-105
        return self._scalar(lib().gvt_engine_compute_dilation, float(r))

    def compute_g_factor(self, r, lam):  # lib.rs:203-205
        return self._scalar(lib().gvt_engine_compute_g_factor, float(r), float(lam))

    def compute_shadow_curve(self, theta_obs, n_points):  # lib.rs:161-170 -> Float32Array [a0, b0, a1, b1, ...]
        n = C.c_uint32(0)
        check(lib().gvt_engine_compute_shadow_curve(self._h, float(theta_obs), int(n_points), None, 0, C.byref(n)))
        out = np.zeros(2 * n.value, np.float32)
        check(lib().gvt_engine_compute_shadow_curve(self._h, float(theta_obs), int(n_points),
                                                    out.ctypes.data_as(C.POINTER(C.c_float)), n.value, C.byref(n)))
        return out

    def compute_shadow_radius(self):  # lib.rs:173-175
        return self._scalar(lib().gvt_engine_compute_shadow_radius)

    def compute_shadow_shift(self, theta_obs):  # lib.rs:179-196 -> Vec<f32>[min_alpha, max_alpha]
        out = np.zeros(2, np.float32)
        check(lib().gvt_engine_compute_shadow_shift(self._h, float(theta_obs), out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def compute_disk_flux(self, r):  # lib.rs:199-201
        return self._scalar(lib().gvt_engine_compute_disk_flux, float(r))

    # ---- spacetime visualisation helpers (lib.rs:139-305): Float32Array results as flat numpy f32 arrays ----
    def _field(self, fn, r_min, r_max, n_radial, n_polar):
        out = np.zeros(3 * int(n_radial) * int(n_polar), np.float32)
        check(fn(self._h, float(r_min), float(r_max), int(n_radial), int(n_polar), out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def compute_kretschner(self, r, theta):  # lib.rs:213-215
        return self._scalar(lib().gvt_engine_compute_kretschner, float(r), float(theta))

    def generate_curvature_field(self, r_min, r_max, n_radial, n_polar):  # lib.rs:219-234 -> (r, theta, K) triples
        return self._field(lib().gvt_engine_generate_curvature_field, r_min, r_max, n_radial, n_polar)

    def compute_light_cone_tilt(self, r, theta):  # lib.rs:238-240
        return self._scalar(lib().gvt_engine_compute_light_cone_tilt, float(r), float(theta))

    def generate_tilt_field(self, r_min, r_max, n_radial, n_polar):  # lib.rs:244-263
        return self._field(lib().gvt_engine_generate_tilt_field, r_min, r_max, n_radial, n_polar)

    def compute_frame_drag_omega(self, r, theta):  # lib.rs:267-269
        return self._scalar(lib().gvt_engine_compute_frame_drag_omega, float(r), float(theta))

    def generate_frame_drag_field(self, r_min, r_max, n_radial, n_polar):  # lib.rs:273-292
        return self._field(lib().gvt_engine_generate_frame_drag_field, r_min, r_max, n_radial, n_polar)

    def compute_flamm_height(self, r):  # lib.rs:296-298
        return self._scalar(lib().gvt_engine_compute_flamm_height, float(r))

    def compute_proper_distance(self, r1, r2, n_steps):  # lib.rs:302-304
        return self._scalar(lib().gvt_engine_compute_proper_distance, float(r1), float(r2), int(n_steps))

    def generate_embedding_mesh(self, r_min, r_max, n_radial, n_angular):  # lib.rs:139-150 -> (x, y, z) triples
        return self._field(lib().gvt_engine_generate_embedding_mesh, r_min, r_max, n_radial, n_angular)

    def generate_ergosphere_mesh(self, n_polar, n_azimuthal):  # lib.rs:153-157
        out = np.zeros(3 * int(n_polar) * int(n_azimuthal), np.float32)
        check(lib().gvt_engine_generate_ergosphere_mesh(self._h, int(n_polar), int(n_azimuthal), out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def generate_disk_lut(self):  # lib.rs:107-110 -> Vec<f32>(512)
        out = np.zeros(512, np.float32)
        check(lib().gvt_engine_generate_disk_lut(self._h, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def get_disk_lut_ptr(self):  # lib.rs:112-114 -> a view of the engine-owned copy (empty before generate_disk_lut)
        p, n = C.POINTER(C.c_float)(), C.c_uint32(0)
        check(lib().gvt_engine_get_disk_lut_ptr(self._h, C.byref(p), C.byref(n)))
        return np.ctypeslib.as_array(p, shape=(n.value,)) if n.value else np.zeros(0, np.float32)

    def generate_spectrum_lut(self, width, height, max_temp):  # lib.rs:128-136 -> Float32Array(4wh)
        out = np.zeros(int(width) * int(height) * 4, np.float32)
        check(lib().gvt_engine_generate_spectrum_lut(self._h, int(width), int(height), float(max_temp),
                                                      out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def get_sab_ptr(self):  # lib.rs:116-118
        p = C.POINTER(C.c_float)()
        check(lib().gvt_engine_get_sab_ptr(self._h, C.byref(p)))
        return np.ctypeslib.as_array(p, shape=(2048,))

    def attach_sab(self, buf):  # lib.rs:74-76
        buf = np.asarray(buf)
        if buf.dtype != np.float32 or buf.size < 2048 or not buf.flags["C_CONTIGUOUS"] or not buf.flags["WRITEABLE"]:
            raise ValueError("attach_sab needs a writable contiguous float32 buffer of >= 2048 elements")
        self._sab_keepalive = buf
        check(lib().gvt_engine_attach_sab(self._h, buf.ctypes.data_as(C.POINTER(C.c_float))))

    def get_sab_layout(self):  # lib.rs:411-419
        out = (C.c_uint32 * 5)()
        check(lib().gvt_engine_get_sab_layout(self._h, out))
        return list(out)

    def set_camera_state(self, px, py, pz, lx=0.0, ly=0.0, lz=0.0):  # lib.rs:120-122
        check(lib().gvt_engine_set_camera_state(self._h, px, py, pz, lx, ly, lz))

    def set_auto_spin(self, enabled):  # lib.rs:124-126
        check(lib().gvt_engine_set_auto_spin(self._h, 1 if enabled else 0))

    def tick_sab(self, dt_override):  # lib.rs:308-409
        check(lib().gvt_engine_tick_sab(self._h, float(dt_override)))

    def integrate_ray_relativistic(self, initial_state, steps, tolerance, use_kerr_schild):  # lib.rs:422-464
        s = list(initial_state)
        if len(s) < 8:
            return s  # lib.rs:429-431: returned unchanged
        a = np.asarray(s[:8], dtype=np.float64)
        out = np.zeros(8)
        pd = C.POINTER(C.c_double)
        check(lib().gvt_engine_integrate_ray(self._h, a.ctypes.data_as(pd), int(steps), float(tolerance),
                                             1 if use_kerr_schild else 0, out.ctypes.data_as(pd), None, None, None))
        return out

    # batched form (not in the reference: one ray per call there); returns dict of arrays
    def integrate_rays(self, states, params):
        xp = np.ascontiguousarray(np.atleast_2d(states), dtype=np.float64)
        n = xp.shape[0]
        out = np.zeros_like(xp)
        term = np.zeros(n, np.uint32)
        steps = np.zeros(n, np.uint32)
        drift = np.zeros(n)
        rhs = np.zeros(n, np.uint32)
        pd, pu = C.POINTER(C.c_double), C.POINTER(C.c_uint32)
        check(lib().gvt_engine_integrate_rays(self._h, C.byref(params.c), n, xp.ctypes.data_as(pd), out.ctypes.data_as(pd),
                                              term.ctypes.data_as(pu), steps.ctypes.data_as(pu), drift.ctypes.data_as(pd),
                                              rhs.ctypes.data_as(pu)))
        return dict(xp=out, term=term, steps=steps, drift=drift, rhs=rhs)
