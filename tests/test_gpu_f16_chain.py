"""The RGBA16F-native frame chain (gvt_render_set_frame_format): trace output, finished frame and TAA history stored as half4 —
the reference's own texture format (rendering/reprojection.ts:120-140, webgpu/renderer.ts:161-180) — with f32 arithmetic in
the resolve and bloom passes. RGBA32F stays the parity format; here the f16 chain is checked against the SAME oracles with the
texture rounding applied where the reference's textures would apply it (on store)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SPIN32 = float(np.float32(0.999))


@pytest.fixture()
def f16_renderer(renderer):
    from gravitas_b200 import _lib
    renderer.set_frame_format(_lib.FORMAT_RGBA16F)
    try:
        yield renderer
    finally:
        renderer.set_frame_format(_lib.FORMAT_RGBA32F)


def rel_err(got, ref):
    ref = np.asarray(ref, np.float64)
    peak = max(float(np.abs(ref[..., :3]).max()), 1e-300)
    return np.abs(np.asarray(got, np.float64) - ref) / np.maximum(np.abs(ref), 1e-3 * peak)


def test_trace_into_rgba16f_frames_is_the_rounded_f32_frame(renderer):
    """The producing kernel's half4 store is the round-to-nearest-even of the very float4 it stores in RGBA32F mode; every
    read-back format of the f16 chain derives from that."""
    from gravitas_b200 import camera, renderer as R, _lib
    W, H, steps = 157, 83, 160
    renderer.init_pipelines(mass=1.0, spin=SPIN32, spec_w=64, spec_h=16, max_temp=1e7)
    cam, _ = camera.default_camera(W, H)
    phys = R.pack_physics(1.0, SPIN32, W, H)
    renderer.params = R.RenderParams(max_steps=steps)
    f32 = np.array(renderer.render(cam, phys))
    renderer.set_frame_format(_lib.FORMAT_RGBA16F)
    try:
        renderer.params = R.RenderParams(max_steps=steps, output_format=_lib.FORMAT_RGBA16F)
        f16 = np.array(renderer.render(cam, phys))                       # delivered straight from the half4 frame, no conversion pass
        assert f16.dtype == np.float16 and renderer.last_stats.kernel_launches == 1
        assert renderer.last_stats.d2h_bytes == W * H * 8 + 64
        assert np.array_equal(f16.view(np.uint16), f32.astype(np.float16).view(np.uint16))
        assert np.array_equal(renderer.read_frame(_lib.FORMAT_RGBA32F), f16.astype(np.float32))
        assert np.array_equal(renderer.read_frame(_lib.FORMAT_RGBA16F).view(np.uint16), f16.view(np.uint16))
        renderer.params = R.RenderParams(max_steps=steps, output_format=_lib.FORMAT_RGBA32F)
        widened = np.array(renderer.render(cam, phys))                   # RGBA32F host frame from the f16 chain: widened on the way out
        assert np.array_equal(widened, f16.astype(np.float32))
        u8 = renderer.read_frame(_lib.FORMAT_RGBA8_REINHARD)
        v = f16.astype(np.float32)
        want = np.rint(np.clip(v / (v + 1.0), 0, 1) * 255.0)
        want[..., 3] = 255
        assert np.abs(u8.astype(np.float32) - want).max() <= 1.0
    finally:
        renderer.set_frame_format(_lib.FORMAT_RGBA32F)


@pytest.mark.parametrize("W,H", [(64, 40), (157, 83), (31, 33), (1, 1)])
def test_taa_on_rgba16f_textures(renderer, W, H):
    """Both resolves on half4 textures: the PRECISE build is the numpy restatement evaluated on the widened texels and rounded
    once on store -- bit for bit; the production build stays within its 1e-3 bar of it."""
    import taa_oracle
    from gravitas_b200 import camera
    rng = np.random.default_rng(11)
    cur = (rng.random((H, W, 4), dtype=np.float32) * 3.0).astype(np.float16)
    hist = (rng.random((H, W, 4), dtype=np.float32) * 3.0).astype(np.float16)
    cur[..., 3] = hist[..., 3] = 1.0
    c0, vp0 = camera.default_camera(W, H, azimuth=math.pi)
    cam, _ = camera.default_camera(W, H, azimuth=math.pi + 0.005, prev_view_proj=vp0)
    for webgl in (False, True):
        ref32 = (taa_oracle.taa_resolve_webgl(cur.astype(np.float32), hist.astype(np.float32), 0.75, False) if webgl
                 else taa_oracle.taa_resolve(cam, cur.astype(np.float32), hist.astype(np.float32)))
        ref = ref32.astype(np.float16)
        got_p = renderer.taa_resolve_webgl(cur, hist, 0.75, False, precise=True) if webgl else renderer.taa_resolve(cam, cur, hist, precise=True)
        got = renderer.taa_resolve_webgl(cur, hist, 0.75, False) if webgl else renderer.taa_resolve(cam, cur, hist)
        assert got_p.dtype == np.float16 and got.dtype == np.float16
        same = float((got_p.view(np.uint16) == ref.view(np.uint16)).all(-1).mean())
        d = np.abs(got.astype(np.float32) - got_p.astype(np.float32))
        print(f"f16 TAA {'webgl' if webgl else 'ataa'} {W}x{H}: precise build bit-identical pixels {same:.4f}; production vs precise max {d.max():.3e}")
        assert np.array_equal(got_p.view(np.uint16), ref.view(np.uint16))
        np.testing.assert_allclose(got.astype(np.float32), ref32, rtol=2e-3, atol=6e-3)      # production build + one f16 ulp (2^-11)


def test_render_taa_bloom_chain_in_rgba16f(f16_renderer, oracle):
    """Three frames of the orbiting, jittered camera through trace -> TAA (precise build) -> bloom (precise build) with every
    frame buffer RGBA16F, against the oracles chained the same way (cur_k = f16 of the f32 trace; out_k = f16 of the resolve
    of the widened textures)."""
    import taa_oracle
    import bloom_oracle
    import gravitas_b200 as g
    from gravitas_b200 import camera, renderer as R, _lib
    r = f16_renderer
    W, H, steps = 120, 68, 96
    r.init_pipelines(mass=1.0, spin=SPIN32, spec_w=64, spec_h=16, max_temp=1e7)
    r.resize(W, H)
    r.reset_history()
    plain = g.KerrRenderer(device=0)
    plain.init()
    plain.init_pipelines(mass=1.0, spin=SPIN32, spec_w=64, spec_h=16, max_temp=1e7)
    try:
        hist = np.zeros((H, W, 4), np.float16)
        prev_vp = None
        for k in range(3):
            cam, vp = camera.default_camera(W, H, azimuth=math.pi + 0.005 * k, prev_view_proj=prev_vp)
            phys = R.pack_physics(1.0, SPIN32, W, H, frame_index=k)
            plain.params = R.RenderParams(max_steps=steps, flags=_lib.FLAG_JITTER)
            cur16 = np.array(plain.render(cam, phys)).astype(np.float16)
            r.params = R.RenderParams(max_steps=steps, flags=_lib.FLAG_JITTER | _lib.FLAG_TAA | _lib.FLAG_TAA_PRECISE,
                                      output_format=_lib.FORMAT_RGBA16F)
            got = np.array(r.render(cam, phys))
            assert got.dtype == np.float16 and r.last_stats.kernel_launches == 2
            ref = taa_oracle.taa_resolve(cam, cur16.astype(np.float32), hist.astype(np.float32)).astype(np.float16)
            assert np.array_equal(got.view(np.uint16), ref.view(np.uint16)), f"frame {k}"
            # the production resolve on the same f16 chain: within its bar
            hist, prev_vp = got, vp
        scene = got.astype(np.float32)
        scale = float(scene[..., :3].max())
        out = r.bloom(enabled=True, intensity=0.5, threshold=0.05 * scale, blur_passes=2, precise=True)
        ref = bloom_oracle.apply_bloom(scene, True, 0.5, 0.05 * scale, 2)
        e = rel_err(out, ref)
        print(f"bloom on the RGBA16F scene: precise build vs numpy max rel err {e.max():.3e}")
        assert e.max() <= 1e-6
        fast = r.bloom(enabled=True, intensity=0.5, threshold=0.05 * scale, blur_passes=2)
        assert np.abs(fast - ref).max() <= 2e-3
    finally:
        plain.cleanup()
