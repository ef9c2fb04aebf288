"""ctypes loader for the CPU ORACLE (oracle/liboracle.so) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module. The product package (blackhole-simulation_b200/gravitas_b200) never does.

PARITY STATUS: "parity unpinned" for integrate / RKF45 / LUT texels / RGBA (see gravitas_oracle.hpp).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

BOYER_LINDQUIST, KERR_SCHILD = 0, 1
METHOD_RKF45, METHOD_RK4, METHOD_SYMPLECTIC, METHOD_VERLET_GLSL = 0, 1, 2, 3
TERM_NONE, TERM_HORIZON, TERM_ESCAPE, TERM_MAXSTEPS, TERM_DISK = 0, 1, 2, 3, 4


class Options(C.Structure):
    _fields_ = [("method", C.c_int32), ("step_rule", C.c_int32), ("tolerance", C.c_double),
                ("initial_step", C.c_double), ("max_steps", C.c_uint64), ("escape_radius", C.c_double),
                ("renormalize_interval", C.c_uint64)]

    @staticmethod
    def default(**kw):
        # geodesic/integrator.rs:36-47 defaults
        o = Options(METHOD_RKF45, 0, 1e-8, 0.01, 10000, 1000.0, 10)
        for k, v in kw.items():
            setattr(o, k, v)
        return o


class RenderParams(C.Structure):
    _fields_ = [("mass", C.c_double), ("spin", C.c_double), ("width", C.c_uint32), ("height", C.c_uint32),
                ("frame_index", C.c_uint32), ("jitter", C.c_int32), ("coords", C.c_int32),
                ("precision", C.c_int32), ("opts", Options), ("disk_r_out", C.c_double),
                ("lut_max_temp", C.c_double), ("spectrum", C.POINTER(C.c_float)), ("spec_w", C.c_uint32),
                ("spec_h", C.c_uint32), ("tdisk", C.POINTER(C.c_float)), ("tdisk_n", C.c_uint32),
                ("_pad", C.c_uint32), ("tdisk_rin", C.c_double), ("tdisk_rout", C.c_double)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "gravitas_oracle.hpp", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def build_native():
    """-O3 -march=native build for the TIMING legs (bench.py cpu_baseline / --impl reference). Host-specific: rebuilt when the
    CPU model differs from the one it was built on (the repo snapshot travels to the GPU box; this directory does not)."""
    d = os.path.join(_HERE, "_native")
    so, stamp = os.path.join(d, "liboracle_native.so"), os.path.join(d, "host.stamp")
    try:
        with open("/proc/cpuinfo") as f:
            host = next((l.split(":", 1)[1].strip() for l in f if l.startswith("model name")), "unknown")
    except OSError:
        host = "unknown"
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "gravitas_oracle.hpp", "glsl_fragment_oracle.hpp", "Makefile")]
    fresh = os.path.exists(so) and os.path.exists(stamp) and open(stamp).read() == host and \
        all(os.path.getmtime(s) <= os.path.getmtime(so) for s in srcs)
    if not fresh:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "native"])
        with open(stamp, "w") as f:
            f.write(host)
    return so


_NATIVE = None


def lib(native=False):
    """native=True: the -O3 -march=native build (timing only; same source, same symbols)."""
    global _LIB, _NATIVE
    if native:
        if _NATIVE is None:
            _NATIVE = _bind(C.CDLL(build_native()))
        return _NATIVE
    if _LIB is None:
        _LIB = _bind(C.CDLL(build()))
    return _LIB


def _bind(L):
    if True:
        d, i, u32, u64 = C.c_double, C.c_int, C.c_uint32, C.c_uint64
        pd, pf = C.POINTER(C.c_double), C.POINTER(C.c_float)
        for name, res, args in [
            ("orc_num_threads", i, []), ("orc_set_num_threads", None, [i]),
            ("orc_event_horizon", d, [d, d, i]), ("orc_isco", d, [d, d, i]), ("orc_photon_sphere", d, [d, d]),
            ("orc_time_dilation", d, [d, d, d, d]),
            ("orc_contravariant", None, [d, d, i, d, d, pd]),
            ("orc_hamiltonian_derivs", None, [d, d, i, d, d, pd, pd]),
            ("orc_hamiltonian", d, [d, d, i, pd]), ("orc_rhs", None, [d, d, i, pd, pd]),
            ("orc_renormalize", None, [d, d, i, pd]), ("orc_rkf45_step", d, [d, d, i, pd, d, pd]),
            ("orc_stepper_step", d, [d, d, i, d, pd, d]),
            ("orc_step_symplectic", None, [d, d, i, pd, d]), ("orc_step_rk4", None, [d, d, i, pd, d]),
            ("orc_integrate", None, [d, d, i, C.POINTER(Options), u64, pd, pd, C.POINTER(u32), C.POINTER(u64),
                                     pd, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
            ("orc_g_factor", d, [d, d, d, d]),
            ("orc_spectrum_lut", None, [u32, u32, d, pf]), ("orc_spectrum_lut_serial", None, [u32, u32, d, pf]),
            ("orc_disk_lut", None, [d, d, u32, pf]), ("orc_page_thorne_flux", d, [d, d, d, d]),
            ("orc_planck", d, [d, d]),
            ("orc_camera_ray", None, [pf, C.POINTER(RenderParams), u32, u32, pd]),
            ("orc_render", d, [pf, C.POINTER(RenderParams), u32, u32, u32, u32, u32, pd, pd, C.POINTER(u32),
                               C.POINTER(u32), pd, C.POINTER(u32), C.POINTER(u64), C.POINTER(u64)]),
            ("orc_flop_census", None, [i, d, pd, d, C.POINTER(u64)]),
            ("orc_fragment_glsl", d, [C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), i, u32, u32, u32, u32, u32, pd,
                                      C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(u64)]),
        ]:
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
    return L


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _pf(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def contravariant(m, a, coords, r, theta):
    g = np.zeros(16)
    lib().orc_contravariant(m, a, coords, r, theta, _pd(g))
    return g


def hamiltonian(m, a, coords, xp):
    xp = np.ascontiguousarray(xp, dtype=np.float64)
    return lib().orc_hamiltonian(m, a, coords, _pd(xp))


def hamiltonian_derivs(m, a, coords, r, theta, p):
    p = np.ascontiguousarray(p, dtype=np.float64)
    out = np.zeros(2)
    lib().orc_hamiltonian_derivs(m, a, coords, r, theta, _pd(p), _pd(out))
    return out


def rhs(m, a, coords, xp):
    xp = np.ascontiguousarray(xp, dtype=np.float64)
    out = np.zeros(8)
    lib().orc_rhs(m, a, coords, _pd(xp), _pd(out))
    return out


def renormalize(m, a, coords, xp):
    xp = np.array(xp, dtype=np.float64)
    lib().orc_renormalize(m, a, coords, _pd(xp))
    return xp


def rkf45_step(m, a, coords, xp, h):
    xp = np.ascontiguousarray(xp, dtype=np.float64)
    out = np.zeros(8)
    err = lib().orc_rkf45_step(m, a, coords, _pd(xp), h, _pd(out))
    return out, err


def stepper_step(m, a, coords, tol, xp, h_try):
    xp = np.array(xp, dtype=np.float64)
    hn = lib().orc_stepper_step(m, a, coords, tol, _pd(xp), h_try)
    return xp, hn


def step_symplectic(m, a, coords, xp, h):
    xp = np.array(xp, dtype=np.float64)
    lib().orc_step_symplectic(m, a, coords, _pd(xp), h)
    return xp


def step_rk4(m, a, coords, xp, h):
    xp = np.array(xp, dtype=np.float64)
    lib().orc_step_rk4(m, a, coords, _pd(xp), h)
    return xp


def integrate(m, a, coords, opts, xp, native=False, threads=None):
    """geodesic::integrate over rays xp[n,8] -> dict of arrays."""
    xp = np.ascontiguousarray(np.atleast_2d(xp), dtype=np.float64)
    n = xp.shape[0]
    out = np.zeros_like(xp)
    term = np.zeros(n, np.uint32)
    steps = np.zeros(n, np.uint64)
    drift = np.zeros(n)
    att = np.zeros(n, np.uint64)
    rej = np.zeros(n, np.uint64)
    rhs_ = np.zeros(n, np.uint64)
    u32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    L = lib(native)
    L.orc_set_num_threads(int(threads) if threads else 0)
    L.orc_integrate(m, a, coords, C.byref(opts), n, _pd(xp), _pd(out), term.ctypes.data_as(u32p),
                        steps.ctypes.data_as(u64p), _pd(drift), att.ctypes.data_as(u64p), rej.ctypes.data_as(u64p),
                        rhs_.ctypes.data_as(u64p))
    return dict(xp=out, term=term, steps=steps, drift=drift, attempts=att, rejects=rej, rhs=rhs_)


def spectrum_lut(w, h, max_temp, serial=False):
    out = np.zeros(w * h * 4, np.float32)
    (lib().orc_spectrum_lut_serial if serial else lib().orc_spectrum_lut)(w, h, max_temp, _pf(out))
    return out


def disk_lut(mass, spin, n=512):
    out = np.zeros(n, np.float32)
    lib().orc_disk_lut(mass, spin, n, _pf(out))
    return out


def make_render_params(width, height, mass=1.0, spin=0.999, opts=None, coords=KERR_SCHILD, precision=0,
                       frame_index=0, jitter=0, disk_r_out=50.0, lut_max_temp=1e7, spectrum=None, spec_w=0,
                       spec_h=0, tdisk=None):
    """Returns (RenderParams, keepalive) — LUT arrays must outlive the struct."""
    rp = RenderParams()
    rp.mass, rp.spin, rp.width, rp.height = mass, spin, width, height
    rp.frame_index, rp.jitter, rp.coords, rp.precision = frame_index, jitter, coords, precision
    rp.opts = opts if opts is not None else Options.default()
    rp.disk_r_out, rp.lut_max_temp = disk_r_out, lut_max_temp
    keep = []
    if spectrum is not None:
        spectrum = np.ascontiguousarray(spectrum, dtype=np.float32)
        keep.append(spectrum)
        rp.spectrum, rp.spec_w, rp.spec_h = _pf(spectrum), spec_w, spec_h
    if tdisk is not None:
        tdisk = np.ascontiguousarray(tdisk, dtype=np.float32)
        keep.append(tdisk)
        rp.tdisk, rp.tdisk_n = _pf(tdisk), tdisk.size
        rp.tdisk_rin = lib().orc_isco(mass, spin, 1)
        rp.tdisk_rout = 50.0 * mass
    return rp, keep


def camera_ray(cam88, rp, px, py):
    cam88 = np.ascontiguousarray(cam88, dtype=np.float32)
    out = np.zeros(8)
    lib().orc_camera_ray(_pf(cam88), C.byref(rp), px, py, _pd(out))
    return out


def render(cam88, rp, x0=0, xs=1, y0=0, y1=None, ys=1, want=("rgba", "xp", "term", "steps", "drift", "crossings"), native=False,
           threads=None):
    """Composite RGBA oracle over the lattice x = x0 + i*xs, y = y0 + j*ys (< y1). native=True: the -O3 -march=native
    build (timing legs only); threads: worker threads for this call (default: all hardware threads)."""
    cam88 = np.ascontiguousarray(cam88, dtype=np.float32)
    if y1 is None:
        y1 = rp.height
    nx = (rp.width - x0 + xs - 1) // xs
    ny = (y1 - y0 + ys - 1) // ys
    n = nx * ny
    u32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    res = {}
    rgba = np.zeros((ny, nx, 4)) if "rgba" in want else None
    xp = np.zeros((ny, nx, 8)) if "xp" in want else None
    term = np.zeros((ny, nx), np.uint32) if "term" in want else None
    steps = np.zeros((ny, nx), np.uint32) if "steps" in want else None
    drift = np.zeros((ny, nx)) if "drift" in want else None
    cross = np.zeros((ny, nx), np.uint32) if "crossings" in want else None
    ts, tr = C.c_uint64(0), C.c_uint64(0)
    L = lib(native)
    L.orc_set_num_threads(int(threads) if threads else 0)
    secs = L.orc_render(
        _pf(cam88), C.byref(rp), x0, xs, y0, y1, ys,
        _pd(rgba) if rgba is not None else None, _pd(xp) if xp is not None else None,
        term.ctypes.data_as(u32p) if term is not None else None,
        steps.ctypes.data_as(u32p) if steps is not None else None,
        _pd(drift) if drift is not None else None,
        cross.ctypes.data_as(u32p) if cross is not None else None, C.byref(ts), C.byref(tr))
    res.update(rgba=rgba, xp=xp, term=term, steps=steps, drift=drift, crossings=cross, seconds=secs,
               total_steps=ts.value, total_rhs=tr.value, n=n, nx=nx, ny=ny)
    return res


def flop_census(what, spin, xp, h=0.1):
    xp = np.ascontiguousarray(xp, dtype=np.float64)
    out = np.zeros(7, np.uint64)
    lib().orc_flop_census(what, spin, _pd(xp), h, out.ctypes.data_as(C.POINTER(C.c_uint64)))
    keys = ("add", "mul", "div", "sqrt", "trig", "pow", "cmp")
    d = {k: int(v) for k, v in zip(keys, out)}
    d["flops"] = d["add"] + d["mul"] + d["div"] + d["sqrt"]
    return d


# ---- the production WebGL2 fragment shader (glsl_fragment_oracle.hpp) ------------------------------------------
class GlslUniforms(C.Structure):  # chunks/common.ts:9-38 + shader-manager #defines; same layout as GvtGlslUniforms
    _fields_ = [("struct_size", C.c_uint32), ("features", C.c_uint32), ("resolution", C.c_float * 2), ("time", C.c_float),
                ("mass", C.c_float), ("spin", C.c_float), ("disk_density", C.c_float), ("disk_temp", C.c_float),
                ("mouse", C.c_float * 2), ("zoom", C.c_float), ("lensing_strength", C.c_float), ("disk_size", C.c_float),
                ("disk_scale_height", C.c_float), ("max_ray_steps", C.c_int32), ("debug", C.c_float),
                ("show_redshift", C.c_float), ("show_kerr_shadow", C.c_float), ("shadow_count", C.c_float),
                ("cam_pos", C.c_float * 3), ("cam_quat", C.c_float * 4), ("shadow_curve", C.c_float * 128)]


def fragment_glsl(uniform_bytes, noise_r, blue_r, precision=0, x0=0, xs=1, y0=0, y1=None, ys=1):
    """Run the GLSL-shader oracle over a pixel lattice. `uniform_bytes`: the raw bytes of a GvtGlslUniforms /
    GlslUniforms structure. noise_r / blue_r: (256, 256) uint8 (the .r channels of u_noiseTex / u_blueNoiseTex)."""
    u = GlslUniforms.from_buffer_copy(bytes(uniform_bytes))
    W, H = int(u.resolution[0]), int(u.resolution[1])
    y1 = H if y1 is None else y1
    nx, ny = (W - x0 + xs - 1) // xs, (y1 - y0 + ys - 1) // ys
    noise_r = np.ascontiguousarray(noise_r, np.uint8); blue_r = np.ascontiguousarray(blue_r, np.uint8)
    assert noise_r.size == 65536 and blue_r.size == 65536
    rgba = np.zeros((ny, nx, 4)); steps = np.zeros((ny, nx), np.uint32); hit = np.zeros((ny, nx), np.uint32)
    photon = np.zeros((ny, nx), np.uint32); total = C.c_uint64(0)
    pu8 = C.POINTER(C.c_uint8); pu32 = C.POINTER(C.c_uint32)
    secs = lib().orc_fragment_glsl(C.addressof(u), noise_r.ctypes.data_as(pu8), blue_r.ctypes.data_as(pu8), precision,
                                   x0, xs, y0, y1, ys, _pd(rgba), steps.ctypes.data_as(pu32), hit.ctypes.data_as(pu32),
                                   photon.ctypes.data_as(pu32), C.byref(total))
    return {"rgba": rgba, "steps": steps, "hit": hit, "photon": photon, "total_steps": total.value, "seconds": secs}
