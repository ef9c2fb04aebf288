"""TAA resolve vs the numpy restatement of ataa.wgsl.ts / reprojection.glsl.ts, in two builds of the kernel:

* the PRECISE build (k_taa_resolve_precise: every operation an IEEE round-to-nearest f32 operation in shader order,
  unfolded reprojection chain) is held to the north_star tolerance: <= 1e-6 relative per component with a floor of
  1e-3 x frame peak for near-zero components -- this is the test of the resolve's SEMANTICS;
* the production build (k_taa_resolve: warp-shuffle 3x3 moments, MUFU.SQRT / MUFU.RCP / MUFU.RSQ, host-folded
  reprojection matrices) is held to 1e-3 of the frame scale against the same oracle, and its distance from the precise
  build is MEASURED and printed: the pass is f32 and two of its steps are ill-conditioned by construction -- sigma =
  sqrt(E[x^2] - E[x]^2) cancels catastrophically on flat neighbourhoods (sqrt(eps_f32) ~ 3.5e-4 relative), and the
  bilinear history fetch turns a 1e-7 relative difference in the reprojected uv into ~1e-5 px x texel contrast. The
  reference stores these frames as RGBA16F (2^-11 ~ 5e-4 relative; reprojection.ts:120-140)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-6


def rel_err(got, ref):
    ref = np.asarray(ref, np.float64)
    peak = max(float(np.abs(ref[..., :3]).max()), 1e-300)
    return np.abs(np.asarray(got, np.float64) - ref) / np.maximum(np.abs(ref), 1e-3 * peak)


def check_precise(got_precise, ref, what):
    e = rel_err(got_precise, ref)
    print(f"{what}: precise build vs numpy: max rel err {e.max():.3e}, bit-identical pixels {float((got_precise == ref).all(-1).mean()):.4f}")
    assert e.max() <= TOL, (what, float(e.max()))


def report_fast(got_fast, got_precise, what):
    e = rel_err(got_fast, got_precise)
    print(f"{what}: production build vs precise build (the measured cost of MUFU + reassociation): "
          f"median {np.median(e):.2e} p99 {np.percentile(e, 99):.2e} max {e.max():.2e}")


def cams(W, H, d_az=0.005):
    from gravitas_b200 import camera
    c0, vp0 = camera.default_camera(W, H, azimuth=math.pi)
    c1, vp1 = camera.default_camera(W, H, azimuth=math.pi + d_az, prev_view_proj=vp0)   # orbit step, config 5
    return c0, c1


@pytest.mark.parametrize("W,H", [(64, 40), (157, 83), (30, 1), (31, 33), (1, 1)])
def test_taa_kernel_matches_numpy(renderer, W, H):
    import taa_oracle
    rng = np.random.default_rng(7)
    cur = rng.random((H, W, 4), dtype=np.float32) * 3.0
    hist = rng.random((H, W, 4), dtype=np.float32) * 3.0
    cur[..., 3] = hist[..., 3] = 1.0
    _, cam = cams(W, H)
    got = renderer.taa_resolve(cam, cur, hist)
    ref = taa_oracle.taa_resolve(cam, cur, hist)
    got_p = renderer.taa_resolve(cam, cur, hist, precise=True)
    check_precise(got_p, ref, f"ataa {W}x{H} white noise")
    report_fast(got, got_p, f"ataa {W}x{H} white noise")
    np.testing.assert_allclose(got, ref, rtol=1e-3, atol=3e-3)       # production build; white noise in [0,3): worst case for both effects
    assert np.all(got[..., 3] == 1.0) and np.all(got_p[..., 3] == 1.0)
    # smooth input: tight
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    smooth = np.stack([1 + np.sin(xx / 9.0) * np.cos(yy / 7.0), 2 + np.cos(xx / 5.0), 1.5 + np.sin(yy / 11.0),
                       np.ones_like(xx)], -1).astype(np.float32)
    sm_hist = np.roll(smooth, 1, axis=1) * np.float32(1.05)
    got_s = renderer.taa_resolve(cam, smooth, sm_hist)
    ref_s = taa_oracle.taa_resolve(cam, smooth, sm_hist)
    check_precise(renderer.taa_resolve(cam, smooth, sm_hist, precise=True), ref_s, f"ataa {W}x{H} smooth")
    np.testing.assert_allclose(got_s, ref_s, rtol=2e-4, atol=2e-3)   # production build


def test_taa_static_scene_converges_and_clamps(renderer):
    import taa_oracle
    W, H = 96, 54
    cam, _ = cams(W, H)                       # prev_view_proj == view_proj: static camera
    rng = np.random.default_rng(3)
    base = rng.random((H, W, 4), dtype=np.float32)
    base[..., 3] = 1.0
    flat = np.full((H, W, 4), 0.5, np.float32)
    out = renderer.taa_resolve(cam, flat, np.zeros_like(flat))
    # flat neighbourhood -> sigma = 0 -> history clamped onto the current colour -> output == current
    # (up to the sqrt(eps_f32) noise of sigma: m2/9 - mean^2 does not cancel exactly because 1/9 is inexact)
    np.testing.assert_allclose(out[..., :3], 0.5, atol=0.5 * 2 * 0.92 * 3.5e-4)
    out2 = renderer.taa_resolve(cam, base, base)
    ref2 = taa_oracle.taa_resolve(cam, base, base)
    np.testing.assert_allclose(out2, ref2, rtol=1e-3, atol=1e-3)     # production build
    check_precise(renderer.taa_resolve(cam, base, base, precise=True), ref2, "ataa static camera")
    # the flat-neighbourhood case in the precise build: sigma is whatever IEEE arithmetic gives, exactly as numpy
    check_precise(renderer.taa_resolve(cam, flat, np.zeros_like(flat), precise=True),
                  taa_oracle.taa_resolve(cam, flat, np.zeros_like(flat)), "ataa flat frame")


def test_render_with_taa_flag_equals_trace_then_resolve(renderer, oracle):
    """gvt_render_frame with GVT_FLAG_TAA == trace kernel output pushed through the TAA oracle, over 3 frames of an
    orbiting, jittered camera (config 5 in miniature); history starts zeroed like a fresh WebGPU texture."""
    import taa_oracle
    from gravitas_b200 import camera, renderer as R, _lib
    W, H, steps = 120, 68, 96
    spin = float(np.float32(0.999))
    renderer.init_pipelines(mass=1.0, spin=spin, spec_w=64, spec_h=16, max_temp=1e7)
    renderer.resize(W, H)
    renderer.reset_history()
    import gravitas_b200 as g
    plain = g.KerrRenderer(device=0)      # a second renderer supplies the un-resolved frames (a non-TAA render on the
    plain.init()                          # first one would overwrite the frame that becomes its TAA history)
    plain.init_pipelines(mass=1.0, spin=spin, spec_w=64, spec_h=16, max_temp=1e7)
    for precise in (False, True):
        renderer.reset_history()
        hist = np.zeros((H, W, 4), np.float32)
        prev_vp = None
        for k in range(3):
            cam, vp = camera.default_camera(W, H, azimuth=math.pi + 0.005 * k, prev_view_proj=prev_vp)
            phys = R.pack_physics(1.0, spin, W, H, frame_index=k)
            plain.params = R.RenderParams(max_steps=steps, flags=_lib.FLAG_JITTER)
            cur = np.array(plain.render(cam, phys))
            renderer.params = R.RenderParams(max_steps=steps, flags=_lib.FLAG_JITTER | _lib.FLAG_TAA |
                                             (_lib.FLAG_TAA_PRECISE if precise else 0))
            got = np.array(renderer.render(cam, phys))
            assert renderer.last_stats.kernel_launches == 2 and renderer.last_stats.taa_ms > 0
            ref = taa_oracle.taa_resolve(cam, cur, hist)
            if precise:
                check_precise(got, ref, f"render + TAA frame {k}")
            else:
                scale = float(np.abs(ref[..., :3]).max())
                np.testing.assert_allclose(got, ref, rtol=1e-3, atol=1e-3 * scale)   # production build
            hist = got
            prev_vp = vp
    plain.cleanup()


@pytest.mark.parametrize("W,H,moving", [(64, 40, False), (157, 83, False), (31, 33, True), (1, 1, False)])
def test_webgl_reprojection_variant(renderer, W, H, moving):
    """reprojection.glsl.ts:70-115 (the WebGL2 pipeline's resolve): +-1.5 sigma, same-texel history, variance-guided alpha."""
    import taa_oracle
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    cur = np.stack([1 + 0.5 * np.sin(xx / 9.0) * np.cos(yy / 7.0), 0.8 + 0.3 * np.cos(xx / 5.0), 0.6 + 0.2 * np.sin(yy / 11.0),
                    np.ones_like(xx)], -1).astype(np.float32) + rng.random((H, W, 4), dtype=np.float32) * 0.05
    cur[..., 3] = 1.0
    hist = (cur * np.float32(1.2) + np.float32(0.1)).astype(np.float32)
    got = renderer.taa_resolve_webgl(cur, hist, blend=0.75, camera_moving=moving)
    ref = taa_oracle.taa_resolve_webgl(cur, hist, 0.75, moving)
    got_p = renderer.taa_resolve_webgl(cur, hist, blend=0.75, camera_moving=moving, precise=True)
    check_precise(got_p, ref, f"reprojection.glsl {W}x{H} moving={moving}")
    report_fast(got, got_p, f"reprojection.glsl {W}x{H}")
    np.testing.assert_allclose(got, ref, rtol=1e-3, atol=1e-3)       # production build
    if moving:
        np.testing.assert_allclose(got[..., :3], cur[..., :3], rtol=1e-5, atol=1e-6)    # alpha = 0: current frame only


def test_render_with_webgl_taa_flag(renderer):
    import taa_oracle
    import gravitas_b200 as g
    from gravitas_b200 import camera, renderer as R, _lib
    W, H, steps = 96, 54, 64
    spin = float(np.float32(0.999))
    renderer.init_pipelines(mass=1.0, spin=spin, spec_w=64, spec_h=16, max_temp=1e7)
    renderer.resize(W, H); renderer.reset_history()
    plain = g.KerrRenderer(device=0); plain.init()
    plain.init_pipelines(mass=1.0, spin=spin, spec_w=64, spec_h=16, max_temp=1e7)
    hist = np.zeros((H, W, 4), np.float32)
    cam, _ = camera.default_camera(W, H)
    for k in range(3):
        phys = R.pack_physics(1.0, spin, W, H, frame_index=k)
        plain.params = R.RenderParams(max_steps=steps, flags=_lib.FLAG_JITTER)
        cur = np.array(plain.render(cam, phys))
        renderer.params = R.RenderParams(max_steps=steps, flags=_lib.FLAG_JITTER | _lib.FLAG_TAA | _lib.FLAG_TAA_WEBGL,
                                         taa_blend=0.75, taa_camera_moving=0)
        got = np.array(renderer.render(cam, phys))
        ref = taa_oracle.taa_resolve_webgl(cur, hist, 0.75, False)
        scale = float(np.abs(ref[..., :3]).max())
        np.testing.assert_allclose(got, ref, rtol=1e-3, atol=1e-3 * scale)
        hist = got
    plain.cleanup()
