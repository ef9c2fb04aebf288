"""Host logic of the WebGLRenderer mirror (gravitas_b200/webgl.py) against what src/rendering/webgl/renderer.ts:173-358 does
each frame: uniform values, feature-toggle -> #define mapping, shadow-curve hand-over from the SAB, fallbacks. No GPU."""
import math

import numpy as np
import pytest


def test_uniform_block_follows_renderer_ts(built):
    from gravitas_b200 import webgl, _lib
    p = dict(mass=2.0, spin=0.6, zoom=25.0, diskSize=40.0, diskScaleHeight=0.15, diskDensity=3.0, diskTemp=12000.0, lensing=0.9)
    u = webgl.make_uniforms(1280, 720, p, mouse=(0.3, 0.6), time=4.2, features=dict(webgl.PRESETS["balanced"]))
    assert (u.resolution[0], u.resolution[1]) == (1280.0, 720.0) and u.time == np.float32(4.2)
    assert u.mass == 2.0 and u.spin == np.float32(0.6 * 2.0)                    # renderer.ts:325-326
    assert u.zoom == 50.0                                                       # :327 zoom * 2
    assert u.disk_size == 40.0 and u.disk_scale_height == np.float32(0.15) and u.disk_density == 3.0
    assert u.disk_temp == np.float32(12000.0 * 2.0 ** -0.25)                    # :351-354 Shakura-Sunyaev mass scaling
    assert u.lensing_strength == np.float32(0.9) and (u.mouse[0], u.mouse[1]) == (np.float32(0.3), np.float32(0.6))
    assert u.max_ray_steps == 64 and u.show_redshift == 0.0 and u.show_kerr_shadow == 0.0 and u.debug == 0.0
    assert tuple(u.cam_pos) == (0.0, 0.0, 0.0) and tuple(u.cam_quat) == (0.0, 0.0, 0.0, 1.0)   # :314-315 fallback camera
    # no telemetry -> 64-point Schwarzschild circle of radius 3 sqrt 3 M (:289-298)
    c = np.array(u.shadow_curve[:]).reshape(64, 2)
    assert u.shadow_count == 64.0
    np.testing.assert_allclose(np.hypot(c[:, 0], c[:, 1]), 3.0 * math.sqrt(3.0) * 2.0, rtol=1e-6)
    assert c[0, 0] > 0 and abs(c[0, 1]) < 1e-6 and c[16, 1] > 0
    # defaults of types/simulation.ts + configs/simulation.config.ts
    d = webgl.make_uniforms(64, 36)
    assert d.mass == 1.0 and d.spin == 0.5 and d.zoom == 60.0 and d.lensing_strength == np.float32(0.7)
    assert d.disk_temp == 9500.0 and d.disk_density == 4.0 and d.max_ray_steps == 128
    assert d.features == webgl.feature_bits(webgl.PRESETS["high-quality"])
    for q, n in webgl.RAY_TRACING_STEPS.items():
        assert webgl.make_uniforms(8, 8, features=dict(webgl.DEFAULT_FEATURES, rayTracingQuality=q)).max_ray_steps == n
    assert webgl.make_uniforms(8, 8, has_post=True).features & _lib.GLSL_LINEAR_OUTPUT
    assert webgl.make_uniforms(8, 8, features=dict(webgl.DEFAULT_FEATURES, kerrShadow=True, gravitationalRedshift=True)).show_kerr_shadow == 1.0


def test_shadow_curve_comes_from_the_sab_when_a_bridge_is_attached(built):
    import gravitas_b200 as g
    from gravitas_b200 import webgl
    w = webgl.WebGLRenderer()
    w.width, w.height = 320, 180
    eng = g.PhysicsEngine(1.0, 0.8)
    eng.set_camera_state(0.0, 5.0, -59.0)
    w.physics_bridge = eng
    u = w.uniforms(dict(mass=1.0, spin=0.8), (0.5, 0.5))
    sab = eng.get_sab_ptr()
    assert u.shadow_count == sab[143] == 64.0
    np.testing.assert_array_equal(np.array(u.shadow_curve[:]), sab[144:272])
    alphas = np.array(u.shadow_curve[0::2])
    assert alphas.max() - alphas.min() > 8.0 and abs(alphas.mean()) > 0.1        # a Kerr D-shape, shifted off-centre
    # telemetry without a curve (count 0) -> the Schwarzschild fallback
    u2 = webgl.make_uniforms(320, 180, dict(mass=1.0), shadow_curve=np.zeros(128, np.float32), shadow_count=0)
    assert u2.shadow_count == 64.0 and abs(math.hypot(u2.shadow_curve[0], u2.shadow_curve[1]) - 3.0 * math.sqrt(3.0)) < 1e-5


def test_webgl_renderer_fails_loudly_without_a_gpu(built):
    import ctypes as C
    from gravitas_b200 import webgl
    n = C.c_int32(0)
    built.lib().gvt_device_count(C.byref(n))
    if n.value > 0:
        pytest.skip("a GPU is present")
    w = webgl.WebGLRenderer()
    assert w.init() is False and w.error and "no CUDA device" in w.error          # like WebGLRenderer.init -> false + .error
