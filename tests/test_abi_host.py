"""CPU-side checks of the product: the C-ABI library loads and exports every symbol include/gravitas_b200.h
declares; host-side (once-per-tick) functions agree with the oracle bit-for-bit; compute entry points fail loudly
without a GPU. No compute kernels run here."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "gravitas_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gvt_[a-z0-9_]+)\s*\(", src)))


def test_every_method_of_the_reference_class_is_mirrored(built):
    """Every `pub fn` of the wasm-bindgen PhysicsEngine (gravitas-wasm/src/lib.rs; list frozen here because
    /root/reference does not exist on the GPU box) has a same-named method on the Python mirror and a C-ABI symbol."""
    ref = """new update_params compute_horizon compute_isco compute_photon_sphere compute_dilation generate_disk_lut
             get_disk_lut_ptr get_sab_ptr attach_sab set_camera_state set_auto_spin generate_spectrum_lut
             generate_embedding_mesh generate_ergosphere_mesh compute_shadow_curve compute_shadow_radius compute_shadow_shift
             compute_disk_flux compute_g_factor compute_kretschner generate_curvature_field compute_light_cone_tilt
             generate_tilt_field compute_frame_drag_omega generate_frame_drag_field compute_flamm_height
             compute_proper_distance tick_sab get_sab_layout integrate_ray_relativistic""".split()
    syms = header_symbols()
    for m in ref:
        if m == "new":
            assert "gvt_engine_create" in syms
            continue
        assert hasattr(built.PhysicsEngine, m), m
        c_name = {"integrate_ray_relativistic": "gvt_engine_integrate_ray"}.get(m, "gvt_engine_" + m)
        assert c_name in syms, c_name
    assert callable(built.init_hooks)
    # Seam B: the public methods of WebGPURenderer (webgpu/renderer.ts:82-280) and WebGLRenderer (webgl/renderer.ts:36-471)
    for m in ("init", "init_pipelines", "update_settings", "get_format", "resize", "render"):
        assert hasattr(built.KerrRenderer, m), m
    for m in ("init", "resize", "render", "cleanup"):
        assert hasattr(built.WebGLRenderer, m), m
    w = built.WebGLRenderer.__new__(built.WebGLRenderer)
    built.WebGLRenderer.__init__(w)
    assert w.error is None and w.on_metrics_update is None and built.KerrRenderer().get_format() == "rgba32float"
    src = os.path.join("/root/reference", "physics-engine", "gravitas-wasm", "src", "lib.rs")
    if os.path.exists(src):       # in the build container: the frozen list really is the reference's
        names = set(re.findall(r"pub fn (\w+)", open(src).read())) - {"init_hooks"}
        assert names == set(ref), names ^ set(ref)


def test_library_exports_every_declared_symbol(built):
    from gravitas_b200 import _lib
    L = C.CDLL(built.lib_path())
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/gravitas_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in gravitas_b200/_lib.py"
    assert built.lib().gvt_abi_version() == 1


def test_struct_layouts_match_reference_uniforms(built):
    from gravitas_b200 import _lib
    assert C.sizeof(_lib.GvtCamera) == 352          # types/webgpu.ts:25 CAMERA_UNIFORM_SIZE
    assert C.sizeof(_lib.GvtPhysicsParams) == 32    # types/webgpu.ts:42 PHYSICS_PARAM_SIZE
    assert _lib.GvtCamera.inv_view.offset == 128 and _lib.GvtCamera.inv_proj.offset == 192
    assert _lib.GvtCamera.prev_view_proj.offset == 256 and _lib.GvtCamera.position.offset == 320
    assert _lib.GvtPhysicsParams.frame_index.offset == 24
    assert built.OFFSETS == {"CONTROL": 0, "CAMERA": 64, "PHYSICS": 128, "TELEMETRY": 256, "LUTS": 2048}


def test_engine_scalars_match_oracle_bitwise(built, oracle):
    L = oracle.lib()
    for m, a in [(1.0, 0.0), (1.0, 0.5), (1.0, 0.9), (1.0, 0.999), (2.5, -0.7), (1.0, 1.0)]:
        e = built.PhysicsEngine(m, a)
        assert e.compute_horizon() == L.orc_event_horizon(m, a, 0)
        assert e.compute_isco() == L.orc_isco(m, a, 1)
        assert e.compute_photon_sphere() == L.orc_photon_sphere(m, a)
        for r in (1.2, 2.0, 3.0, 10.0):
            td = L.orc_time_dilation(m, a, r, np.pi / 2)
            assert e.compute_dilation(r) == (100.0 if td <= 0 else 1.0 / td)   # lib.rs:97-105
            assert e.compute_g_factor(r + 5, 1.3) == L.orc_g_factor(r + 5, m, a, 1.3)
        assert e.get_sab_layout() == [0, 64, 128, 256, 2048]


def test_luts_match_oracle_bitwise(built, oracle):
    e = built.PhysicsEngine(1.0, 0.999)
    assert np.array_equal(e.generate_disk_lut(), oracle.disk_lut(1.0, 0.999))
    assert np.array_equal(e.generate_spectrum_lut(48, 12, 1e7), oracle.spectrum_lut(48, 12, 1e7, serial=True))
    e2 = built.PhysicsEngine(1.0, 0.0)
    assert np.array_equal(e2.generate_disk_lut(), oracle.disk_lut(1.0, 0.0))
    e_new = built.PhysicsEngine(1.0, 0.9)
    assert e_new.get_disk_lut_ptr().size == 0                    # lib.rs:66: the engine-owned copy starts empty
    lut = e.generate_disk_lut()
    assert lut.shape == (512,) and lut[0] == 0.0 and lut.max() == 1.0
    assert np.array_equal(e.get_disk_lut_ptr(), lut)              # lib.rs:108-114
    assert built.init_hooks() is None


def test_tick_sab_protocol(built):
    e = built.PhysicsEngine(1.0, 0.9)
    sab = e.get_sab_ptr()
    assert sab.shape == (2048,) and not sab.any()
    sab[1], sab[2], sab[3], sab[4] = 0.25, -0.5, 0.1, 0.016     # CONTROL: mouse_dx, mouse_dy, zoom, dt
    e.tick_sab(0.0)                                             # dt <= 0 -> read control[4] (lib.rs:319-323)
    assert sab[1] == 0 and sab[2] == 0 and sab[3] == 0 and sab[4] == np.float32(0.016)   # inputs consumed
    cam = sab[64:76]
    yaw = -0.25 * 2.0 * float(np.float32(0.016))
    zf = 1.0 + float(np.float32(0.1)) * float(np.float32(0.016))
    np.testing.assert_allclose(cam[0:3], [20 * np.sin(yaw) * zf, 0.0, 20 * np.cos(yaw) * zf], rtol=1e-6, atol=1e-7)
    assert list(cam[8:12]) == [0.0, 1.0, 0.0, 0.0]              # orientation xyzw (camera.rs:31)
    phys = sab[128:256]
    assert phys[0] == np.float32(e.compute_horizon()) and phys[1] == np.float32(e.compute_isco())
    assert phys[2] == 1.0 and phys[3] == np.float32(0.9) and phys[15] == 64.0
    assert phys[4] < 0 < phys[5]                                # shadow extents min/max alpha
    # The reference writes 64 (alpha,beta) pairs from PHYSICS+16, i.e. f32 indices 144..271: the last 16 floats
    # spill into the TELEMETRY block and the "sequence" increment lands on curve[56].alpha (lib.rs:381-407). Kept.
    curve_tail = sab[256:272].copy()
    assert np.any(curve_tail != 0)
    # external SAB (attach_sab, lib.rs:74-76) receives the same writes
    ext = np.zeros(4096, np.float32)
    e2 = built.PhysicsEngine(1.0, 0.9)
    e2.attach_sab(ext)
    ext[1], ext[2], ext[3], ext[4] = 0.25, -0.5, 0.1, 0.016
    e2.tick_sab(0.0)
    assert np.array_equal(ext[:2048], sab)
    # NaN guard: camera falls back to last-good (lib.rs:339-343)
    e2.set_camera_state(float("nan"), 0.0, 0.0)
    e2.tick_sab(0.016)
    assert np.all(np.isfinite(ext[64:76]))
    # auto spin rotates about Y at 0.15 rad/s (camera.rs:60-65)
    e3 = built.PhysicsEngine(1.0, 0.0)
    e3.set_auto_spin(True)
    e3.tick_sab(1.0)
    s3 = e3.get_sab_ptr()
    np.testing.assert_allclose(s3[64:67], [20 * np.sin(0.15), 0, 20 * np.cos(0.15)], rtol=1e-6, atol=1e-6)
    assert s3[143] == 32.0                                      # a = 0 -> 32-point circle (shadow.rs:90-99)
    np.testing.assert_allclose(np.hypot(s3[144:208:2], s3[145:208:2]), 3 * np.sqrt(3), rtol=1e-6)


def test_camera_uniforms_are_consistent(built):
    from gravitas_b200 import camera
    cam, vp = camera.default_camera(1920, 1080)
    V, P = cam[0:16].reshape(4, 4).T.astype(np.float64), cam[16:32].reshape(4, 4).T.astype(np.float64)
    IV, IP = cam[32:48].reshape(4, 4).T.astype(np.float64), cam[48:64].reshape(4, 4).T.astype(np.float64)
    np.testing.assert_allclose(V @ IV, np.eye(4), atol=1e-5)
    np.testing.assert_allclose(P @ IP, np.eye(4), atol=1e-4)
    np.testing.assert_allclose(np.linalg.norm(cam[80:83]), 30.0, rtol=1e-6)
    assert abs(np.degrees(np.arccos(cam[81] / 30.0)) - 97.0) < 1e-4
    np.testing.assert_allclose(vp.reshape(4, 4).T, P @ V, rtol=1e-5, atol=1e-5)
    assert cam[83] == 0 and cam[87] == 0


def test_compute_fails_loudly_without_gpu(built):
    n = C.c_int32(-1)
    rc = built.lib().gvt_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present; the no-device path is exercised on CPU-only hosts")
    e = built.PhysicsEngine(1.0, 0.9)
    with pytest.raises(built.GravitasError) as ei:
        e.integrate_ray_relativistic([0, 20, 1.57, 0, -1, -1, 0, 3.5], 100, 1e-8, True)
    assert ei.value.code == -2
    assert list(e.integrate_ray_relativistic([1, 2, 3], 100, 1e-8, True)) == [1, 2, 3]   # lib.rs:429-431
    r = built.KerrRenderer()
    with pytest.raises(built.GravitasError) as ei:
        r.init()
    assert ei.value.code == -2


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "blackhole-simulation_b200")
    for dp, _, fns in os.walk(pkg):
        if os.path.basename(dp) in ("build", "__pycache__"):
            continue
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                assert "liboracle" not in txt and "gravitas_oracle" not in txt and "import oracle" not in txt, fn


def test_napi_shim_compiles_against_the_header():
    """addon/binding.cc (the N-API shim of INTEGRATION.md) is type-checked against include/gravitas_b200.h and a
    declaration-only subset of Node's node_api.h written out in addon/stub/ (node and its headers are not in this image):
    this proves the C++ is well-formed against those signatures, nothing about loading under node."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    p = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "addon", "stub"),
                        "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "addon", "binding.cc")],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    src = open(os.path.join(ROOT, "addon", "binding.cc")).read()
    for name in ("tick_sab", "attach_sab", "update_params", "set_camera_state", "set_auto_spin", "compute_horizon",
                 "compute_isco", "compute_photon_sphere", "compute_dilation", "generate_disk_lut",
                 "generate_spectrum_lut", "integrate_ray_relativistic", "get_sab_layout", "get_sab_ptr"):
        assert f'"{name}"' in src, f"PhysicsEngine.{name} (gravitas-wasm/src/lib.rs) missing from the shim"


def test_shadow_and_flux_methods(built, oracle):
    """compute_shadow_curve / _shift / _radius / compute_disk_flux (lib.rs:161-201): the Bardeen curve against an
    independent Python restatement of shadow.rs, the flux against the oracle's page_thorne_flux, and the
    reference's own unit-test properties (shadow.rs tests: a=0 circle of radius 3 sqrt 3 M; D-shape shifts with spin)."""
    import pyref
    import ctypes as C
    from gravitas_b200 import _lib
    L = oracle.lib()
    for m, a, th, n in [(1.0, 0.9, 1.2, 32), (1.0, 0.999, math.radians(97.0), 32), (2.5, -0.5, 0.4, 17), (1.0, 0.0, 1.0, 12),
                        (1.0, 0.7, 0.0, 8), (1.0, 0.3, math.pi / 2, 1)]:
        e = built.PhysicsEngine(m, a)
        got = e.compute_shadow_curve(th, n)
        ref = np.array(pyref.bardeen_shadow(m, a, th, n), np.float64).astype(np.float32).ravel()
        assert got.shape == ref.shape, (m, a, th, n)
        np.testing.assert_allclose(got, ref, rtol=2e-6, atol=2e-6)
        shift = e.compute_shadow_shift(th)
        alphas = np.array(pyref.bardeen_shadow(m, a, th, 32))[:, 0]
        np.testing.assert_allclose(shift, [alphas.min(), alphas.max()], rtol=2e-6, atol=2e-6)
        assert e.compute_shadow_radius() == 3.0 * math.sqrt(3.0) * m
        for r in (e.compute_isco() * 0.9, e.compute_isco() + 0.5, 10.0 * m, 45.0 * m):
            assert e.compute_disk_flux(r) == L.orc_page_thorne_flux(r, m, a, 1.0)
    # capacity smaller than the curve: only `capacity` pairs are written, the full length is still reported
    e = built.PhysicsEngine(1.0, 0.9)
    buf = np.full(8, -7.0, np.float32)
    n = C.c_uint32(0)
    _lib.check(_lib.lib().gvt_engine_compute_shadow_curve(e._h, 1.2, 32, buf.ctypes.data_as(C.POINTER(C.c_float)), 3, C.byref(n)))
    assert n.value == 64 and np.all(buf[6:] == -7.0) and np.all(buf[:6] != -7.0)
    # tick_sab publishes the same curve (lib.rs:381-392)
    e.set_camera_state(0.0, 20.0 * math.cos(1.2), 20.0 * math.sin(1.2))
    e.tick_sab(0.016)
    sab = e.get_sab_ptr()
    cam = sab[64:67].astype(np.float64)
    th_cam = math.acos(cam[1] / np.linalg.norm(cam))
    np.testing.assert_allclose(sab[144:144 + 112], e.compute_shadow_curve(th_cam, 32)[:112], rtol=1e-5, atol=1e-5)


def test_spacetime_visualisation_helpers(built):
    """The PhysicsEngine's visualisation helpers (lib.rs:139-305 over gravitas-core/src/spacetime/) against an independent
    Python restatement of the Rust, plus the reference's own unit tests for them (embedding.rs:113-129) and closed forms."""
    import pyref
    # embedding.rs tests: flamm_height(2, 1) == 0; at r = 100: |z - 2 sqrt(2 * 98)| < 0.1
    e1 = built.PhysicsEngine(1.0, 0.0)
    assert e1.compute_flamm_height(2.0) == 0.0
    assert abs(e1.compute_flamm_height(100.0) - 2.0 * math.sqrt(2.0 * 98.0)) < 0.1
    # Schwarzschild closed forms: K = 48 M^2 / r^6 (curvature.rs:41-43); no frame dragging; tan(tilt) = 1 - 2M/r
    assert abs(e1.compute_kretschner(5.0, 1.0) - 48.0 / 5.0 ** 6) < 1e-18
    assert e1.compute_frame_drag_omega(7.0, 1.1) == 0.0
    assert abs(math.tan(e1.compute_light_cone_tilt(10.0, 0.7)) - (1.0 - 2.0 / 10.0)) < 1e-14
    assert e1.compute_light_cone_tilt(1.9, 0.7) == math.pi / 2                      # inside the horizon
    # proper distance from 4M to 10M in Schwarzschild: closed form of int dr / sqrt(1 - 2M/r)
    F = lambda r: math.sqrt(r * (r - 2.0)) + 2.0 * math.atanh(math.sqrt(1.0 - 2.0 / r))
    assert abs(e1.compute_proper_distance(4.0, 10.0, 4000) - (F(10.0) - F(4.0))) < 1e-6
    assert e1.compute_proper_distance(10.0, 4.0, 50) == e1.compute_proper_distance(4.0, 10.0, 50)
    for m, a in [(1.0, 0.9), (2.5, -0.6), (1.0, 0.0), (0.7, 0.999)]:
        e = built.PhysicsEngine(m, a)
        for r, th in [(3.0 * m, 0.4), (6.0 * m, 1.5707963267948966), (20.0 * m, 2.9), (1.2 * m, 1.0)]:
            np.testing.assert_allclose(e.compute_kretschner(r, th), pyref.kretschner_kerr(r, th, m, a), rtol=1e-13)
            np.testing.assert_allclose(e.compute_light_cone_tilt(r, th), pyref.light_cone_tilt_bl(m, a, r, th), rtol=1e-14, atol=1e-16)
            np.testing.assert_allclose(e.compute_frame_drag_omega(r, th), pyref.frame_dragging(m, a, r, th), rtol=1e-14, atol=1e-300)
        np.testing.assert_allclose(e.compute_proper_distance(3.0 * m, 9.0 * m, 64), pyref.proper_distance(m, a, 3.0 * m, 9.0 * m, 64), rtol=1e-13)
        nr, npol = 5, 7
        lat = list(pyref.field_lattice(2.5 * m, 12.0 * m, nr, npol))
        for gen, fn in [(e.generate_curvature_field, lambda r, t: pyref.kretschner_kerr(r, t, m, a)),
                        (e.generate_tilt_field, lambda r, t: pyref.light_cone_tilt_bl(m, a, r, t)),
                        (e.generate_frame_drag_field, lambda r, t: pyref.frame_dragging(m, a, r, t))]:
            got = gen(2.5 * m, 12.0 * m, nr, npol).reshape(-1, 3)
            ref = np.array([[r, t, fn(r, t)] for r, t in lat]).astype(np.float32)
            np.testing.assert_allclose(got, ref, rtol=2e-7, atol=1e-30)
        np.testing.assert_allclose(e.generate_embedding_mesh(2.2 * m, 15.0 * m, 6, 8),
                                   np.array(pyref.embedding_mesh(m, a, 2.2 * m, 15.0 * m, 6, 8)).astype(np.float32), rtol=2e-7, atol=1e-6)
        np.testing.assert_allclose(e.generate_ergosphere_mesh(9, 12),
                                   np.array(pyref.ergosphere_mesh(m, a, 9, 12)).astype(np.float32), rtol=2e-7, atol=1e-6)
        # the ergosphere touches the horizon at the poles and reaches 2M at the equator
        mesh = e.generate_ergosphere_mesh(9, 12).reshape(9, 12, 3)
        assert abs(np.linalg.norm(mesh[4, 0]) - 2.0 * m) < 1e-5 * m and abs(abs(mesh[0, 0, 1]) - e.compute_horizon()) < 1e-5 * m
    with pytest.raises(built.GravitasError):
        e1.generate_curvature_field(2.0, 10.0, 1, 5)          # the reference divides by (n - 1)


def test_unchanged_worker_sequence_against_the_module_contract(built):
    """Replays src/workers/physics.worker.ts:60-68,111-176 VERBATIM against the module shape of addon/ts/index.ts (its
    Python twin gravitas_b200/wasm_module.py): `init()` -> `.memory.buffer`, `new PhysicsEngine`, `tick_sab`, seqlock +1,
    `get_sab_ptr() / 4` -> `wasmF32.subarray(...)` copies of the CAMERA and PHYSICS blocks into the app's SAB, seqlock +1.
    The copied values must be the ones an engine writing straight into an attached SAB produces, the pointer a non-zero
    byte offset inside memory.buffer, and two engines must not share a region."""
    from gravitas_b200 import wasm_module as M
    OFF = built.OFFSETS
    # --- worker INIT (physics.worker.ts:38-68)
    sab = np.zeros(2 * 1024 * 1024 // 4, np.float32)                     # data.sab: SharedArrayBuffer(2 MiB), physics-bridge.ts:60
    sab_seq = sab.view(np.int32)                                         # sabSequenceView = new Int32Array(sab)
    wasm_module = M.init()                                               # await wasmModuleWrap.default()
    engine = M.PhysicsEngine(1.0, 0.9)                                   # new PhysicsEngine(mass, spin)
    wasm_memory = wasm_module.memory                                     # (self as any).wasmMemory = wasmModule.memory
    # --- a reference engine that writes an attached SAB directly (lib.rs:74): the ground truth for the copied blocks
    direct = built.PhysicsEngine(1.0, 0.9)
    direct_sab = np.zeros(4096, np.float32)
    direct.attach_sab(direct_sab)
    for e in (engine, direct):
        e.set_camera_state(0.0, -3.7, 29.8, 0.0, 0.0, 0.0)
        e.set_auto_spin(True)
    ptr0 = engine.get_sab_ptr()
    assert isinstance(ptr0, int) and ptr0 > 0 and ptr0 % 4 == 0 and ptr0 + 2048 * 4 <= wasm_memory.buffer.nbytes
    for k in range(5):
        # --- calculate() (physics.worker.ts:111-176)
        wasm_f32 = wasm_memory.buffer.view(np.float32)                   # wasmF32 = new Float32Array(memory.buffer)
        engine.tick_sab(0.016)                                           # engine.tick_sab(clampedDt)
        direct.tick_sab(0.016)
        sab_seq[OFF["TELEMETRY"]] += 1                                   # Atomics.add(sabSequenceView, OFFSETS.TELEMETRY, 1)
        start = engine.get_sab_ptr() // 4                                # wasmSABPtr / 4
        sab[OFF["CAMERA"]:OFF["PHYSICS"]] = wasm_f32[start + OFF["CAMERA"]:start + OFF["PHYSICS"]]
        sab[OFF["PHYSICS"]:OFF["TELEMETRY"]] = wasm_f32[start + OFF["PHYSICS"]:start + OFF["TELEMETRY"]]
        sab_seq[OFF["TELEMETRY"]] += 1
        assert sab_seq[OFF["TELEMETRY"]] == 2 * (k + 1)                  # an int32 counter, +2 per tick
        assert np.array_equal(sab[OFF["CAMERA"]:OFF["TELEMETRY"]], direct_sab[OFF["CAMERA"]:OFF["TELEMETRY"]])
        assert sab[OFF["PHYSICS"]] == np.float32(engine.compute_horizon()) and np.any(sab[OFF["CAMERA"]:OFF["CAMERA"] + 3] != 0)
    # --- physics-bridge.ts:105-125 fallback views: inputs written through memory.buffer at ptr + CONTROL*4 reach the engine
    ctrl = wasm_memory.buffer[ptr0 + OFF["CONTROL"] * 4: ptr0 + OFF["CONTROL"] * 4 + 64].view(np.float32)
    ctrl[1] = 0.25; ctrl[4] = 0.016
    cam_before = wasm_memory.buffer.view(np.float32)[ptr0 // 4 + OFF["CAMERA"]: ptr0 // 4 + OFF["CAMERA"] + 3].copy()
    engine.tick_sab(0.0)                                                 # dt from CONTROL[4] (lib.rs:317-328)
    assert ctrl[1] == 0.0                                                # consumed and zeroed by the engine
    assert np.any(wasm_memory.buffer.view(np.float32)[ptr0 // 4 + OFF["CAMERA"]: ptr0 // 4 + OFF["CAMERA"] + 3] != cam_before)
    other = M.PhysicsEngine(1.0, 0.5)
    assert other.get_sab_ptr() != ptr0 and abs(other.get_sab_ptr() - ptr0) >= 2048 * 4
    assert M.init().memory is wasm_memory                                # one memory object per module, like a wasm instance


def test_addon_sources_honour_the_memory_contract():
    """Static checks of the shim sources (node is absent here): index.ts allocates one SharedArrayBuffer, attaches a region
    per engine and never hands out an unrelated buffer; binding.cc exports get_sab_ptr, holds a reference on the attached
    view, and validates every caller buffer it writes into."""
    ts = open(os.path.join(ROOT, "addon", "ts", "index.ts")).read()
    cc = open(os.path.join(ROOT, "addon", "binding.cc")).read()
    assert ts.count("new SharedArrayBuffer(") == 1 and "this.attach_sab(region)" in ts and "return { memory }" in ts
    assert '"get_sab_ptr"' in cc and "napi_create_reference" in cc and "napi_delete_reference" in cc
    assert "gvt_render_get_size" in cc and cc.count("output buffer") + cc.count("out: ArrayBuffer") >= 3
    for name in ("attach_sab", "get_sab_ptr", "tick_sab"):
        assert name in ts
