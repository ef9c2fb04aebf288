"""Edge cases of the C ABI on the GPU: degenerate sizes, zero budgets, argument validation, formats, LUT residency."""
import ctypes as C
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SPIN = float(np.float32(0.999))


def _setup(renderer, W, H, **kw):
    from gravitas_b200 import camera, renderer as R
    renderer.init_pipelines(mass=1.0, spin=SPIN, spec_w=64, spec_h=16, max_temp=1e7)
    renderer.params = R.RenderParams(**kw)
    cam, _ = camera.default_camera(W, H)
    return cam, R.pack_physics(1.0, SPIN, W, H)


@pytest.mark.parametrize("W,H", [(1, 1), (7, 3), (8, 4), (9, 5), (33, 1), (1, 37)])
def test_tiny_and_ragged_frames(renderer, oracle, W, H):
    cam, phys = _setup(renderer, W, H, max_steps=64)
    frame = np.array(renderer.render(cam, phys))
    assert frame.shape == (H, W, 4) and np.all(frame[..., 3] == 1.0) and np.isfinite(frame).all()
    spec, td = oracle.spectrum_lut(64, 16, 1e7), oracle.disk_lut(1.0, SPIN)
    opts = oracle.Options.default(method=2, step_rule=1, max_steps=64)
    rp, keep = oracle.make_render_params(W, H, 1.0, SPIN, opts, spectrum=spec, spec_w=64, spec_h=16, tdisk=td)
    ref = oracle.render(cam, rp, want=("rgba",))
    peak = max(float(ref["rgba"][..., :3].max()), 1e-30)
    assert (np.abs(frame - ref["rgba"]) <= 1e-6 * np.maximum(np.abs(ref["rgba"]), 1e-3 * peak) + 1e-30).all()
    st = renderer.last_stats
    assert st.n_horizon + st.n_escape + st.n_maxsteps + st.n_disk == W * H


def test_zero_step_budget_and_single_step(renderer):
    cam, phys = _setup(renderer, 40, 24, max_steps=0)
    frame = np.array(renderer.render(cam, phys))
    assert not frame[..., :3].any() and renderer.last_stats.steps_committed == 0
    assert renderer.last_stats.n_maxsteps == 40 * 24            # integrate(): loop body never runs -> MaxSteps
    renderer.params.c.max_steps = 1
    renderer.render(cam, phys)
    assert renderer.last_stats.steps_committed == 40 * 24 and renderer.last_stats.rhs_evals == 3 * 40 * 24


def test_argument_validation(renderer, built):
    from gravitas_b200 import camera, renderer as R, _lib
    cam, phys = _setup(renderer, 32, 16, max_steps=8)
    renderer.params.c.renormalize_interval = 0
    with pytest.raises(built.GravitasError):
        renderer.render(cam, phys)
    renderer.params = R.RenderParams(coords=_lib.COORDS_BL)      # render path is Kerr-Schild only
    with pytest.raises(built.GravitasError) as ei:
        renderer.render(cam, phys)
    assert ei.value.code == _lib.GVT_ERR_UNSUPPORTED
    renderer.params = R.RenderParams(method=7)
    with pytest.raises(built.GravitasError):
        renderer.render(cam, phys)
    zero = R.pack_physics(1.0, SPIN, 0, 16)
    renderer.params = R.RenderParams()
    with pytest.raises(built.GravitasError):
        renderer.render(cam, zero)
    cam0 = np.array(cam).copy(); cam0[80:83] = 0.0               # camera at the origin
    with pytest.raises(built.GravitasError):
        renderer.render(cam0, phys)
    with pytest.raises(built.GravitasError):
        renderer.trace_states(cam, phys, x0=40)                  # lattice outside the frame
    assert built.lib().gvt_render_frame(None, None, None, None, None, None) == _lib.GVT_ERR_INVALID
    fresh = built.KerrRenderer(device=0)
    fresh.init()
    with pytest.raises(built.GravitasError):                     # LUTs not initialised
        fresh.render(cam, phys)
    fresh.cleanup()
    bad = built.KerrRenderer(device=99)
    with pytest.raises(built.GravitasError):
        bad.init()


def test_spin_is_clamped_like_kerr_new(renderer, oracle):
    """Kerr::new clamps a* to [-1, 1] (kerr.rs:49-55): spin 1.7 renders exactly like spin 1.0."""
    from gravitas_b200 import camera, renderer as R
    W, H = 48, 28
    renderer.init_pipelines(mass=1.0, spin=1.0, spec_w=64, spec_h=16, max_temp=1e7)
    renderer.params = R.RenderParams(max_steps=96)
    cam, _ = camera.default_camera(W, H)
    a = np.array(renderer.render(cam, R.pack_physics(1.0, 1.7, W, H)))
    b = np.array(renderer.render(cam, R.pack_physics(1.0, 1.0, W, H)))
    assert np.array_equal(a, b)


def test_rgba16f_output_and_pageable_host_buffer(renderer):
    from gravitas_b200 import _lib
    cam, phys = _setup(renderer, 96, 54, max_steps=128)
    f32 = np.array(renderer.render(cam, phys))                   # pinned buffer: kernel stores straight to host
    assert np.array_equal(renderer.read_frame(), f32)            # the device-resident frame says the same
    renderer.params.c.output_format = _lib.FORMAT_RGBA16F
    f16 = np.array(renderer.render(cam, phys))
    assert f16.dtype == np.float16 and np.array_equal(f16, f32.astype(np.float16))
    assert np.array_equal(renderer.read_frame(_lib.FORMAT_RGBA16F), f16)
    # pageable (ordinary numpy) destination goes through the copy path and gives the same bits
    renderer.params.c.output_format = _lib.FORMAT_RGBA32F
    out = np.zeros((54, 96, 4), np.float32)
    st = _lib.GvtFrameStats()
    from gravitas_b200.renderer import pack_camera
    cam_s = pack_camera(cam)
    rc = renderer._h and _lib.lib().gvt_render_frame(renderer._h, C.byref(cam_s), C.byref(phys), C.byref(renderer.params.c),
                                                     out.ctypes.data_as(C.c_void_p), C.byref(st))
    assert rc == 0 and np.array_equal(out, f32) and st.d2h_bytes >= out.nbytes


def test_large_lut_stays_in_global_memory(renderer, oracle):
    """A LUT above the shared-memory budget (512x128 RGBA32F = 1 MB) is sampled from global memory: same pixels as
    the oracle with the same LUT."""
    from gravitas_b200 import camera, renderer as R
    W, H, steps = 96, 54, 128
    renderer.init_pipelines(mass=1.0, spin=SPIN, spec_w=512, spec_h=128, max_temp=1e7)
    renderer.params = R.RenderParams(max_steps=steps)
    cam, _ = camera.default_camera(W, H)
    phys = R.pack_physics(1.0, SPIN, W, H)
    frame = np.array(renderer.render(cam, phys))
    spec, td = oracle.spectrum_lut(512, 128, 1e7), oracle.disk_lut(1.0, SPIN)
    opts = oracle.Options.default(method=2, step_rule=1, max_steps=steps)
    rp, keep = oracle.make_render_params(W, H, 1.0, SPIN, opts, spectrum=spec, spec_w=512, spec_h=128, tdisk=td)
    ref = oracle.render(cam, rp, want=("rgba",))
    peak = float(ref["rgba"][..., :3].max())
    assert (np.abs(frame - ref["rgba"]) <= 1e-6 * np.maximum(np.abs(ref["rgba"]), 1e-3 * peak)).all()


def test_default_integration_options_ray(built, oracle):
    """integrate_ray_relativistic with the library defaults' 10000-step budget (integrator.rs:36-47) on rays that
    escape, fall in, and orbit."""
    e = built.PhysicsEngine(1.0, 0.9)
    for ray in ([0, 20, math.pi / 2, 0, -1, -1, 0, 3.5], [0, 6, 1.2, 0.5, -1, -1, 0.3, 1.0], [0, 8, math.pi / 2, 0, -1, -0.9, 0, 4.6]):
        for ks in (True, False):
            got = e.integrate_ray_relativistic(ray, 10000, 1e-8, ks)
            ref = oracle.integrate(1, 0.9, 1 if ks else 0, oracle.Options.default(), ray)["xp"][0]
            np.testing.assert_allclose(got, ref, rtol=1e-7, atol=1e-10)


def test_display_ready_rgba8_outputs(renderer):
    """GVT_FORMAT_RGBA8_REINHARD = the WebGPU blit (webgpu/renderer.ts:45-47); GVT_FORMAT_RGBA8_ACES = the WebGL final pass
    without bloom (bloom.glsl.ts:106-124). Checked against numpy on the same HDR frame (<= 1 code value: powf/rounding)."""
    from gravitas_b200 import _lib
    cam, phys = _setup(renderer, 128, 72, max_steps=160)
    hdr = np.array(renderer.render(cam, phys)).astype(np.float32)
    for fmt in (_lib.FORMAT_RGBA8_REINHARD, _lib.FORMAT_RGBA8_ACES):
        renderer.params.c.output_format = fmt
        got = np.array(renderer.render(cam, phys))
        assert got.dtype == np.uint8 and got.shape == (72, 128, 4) and np.all(got[..., 3] == 255)
        assert np.array_equal(renderer.read_frame(fmt), got)
        c = hdr[..., :3].astype(np.float64)
        if fmt == _lib.FORMAT_RGBA8_REINHARD:
            ref = c / (c + 1.0)
        else:
            ref = np.clip((c * (2.51 * c + 0.03)) / (c * (2.43 * c + 0.59) + 0.14), 0.0, 1.0) ** 0.4545
        ref8 = np.rint(np.clip(ref, 0, 1) * 255.0)
        assert np.abs(got[..., :3].astype(np.float64) - ref8).max() <= 1
        assert got[..., :3].max() > 0
    renderer.params.c.output_format = _lib.FORMAT_RGBA32F
