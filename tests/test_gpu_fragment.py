"""k_fragment_glsl (the reference's production WebGL2 fragment shader as one fused CUDA kernel) vs the C++ restatement
of the GLSL in oracle/glsl_fragment_oracle.hpp, through the C ABI (gvt_render_fragment_glsl).

Tolerances. The f64 instantiation must agree with the f64 oracle to rounding on every pixel whose march is not
sitting on a discontinuity of the shader (floor() star cells, `if (x > threshold)` gates, integer step counts): the
test requires identical per-pixel step counts / horizon flags on >= 99.9 % of pixels and 1e-9 absolute agreement of
the (tone-mapped, [0,1]) colours on those. The f32 instantiation is the shader's own arithmetic; CUDA's and glibc's
single-precision sin/exp/pow/log differ by an ulp or two and a 300-step march amplifies that, so f32 is checked
statistically against the f32 oracle (median, 99th percentile) and against the f64 oracle as an accuracy statement."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wgl(built):
    from gravitas_b200 import webgl
    r = webgl.WebGLRenderer(device=0, noise_seed=11)
    assert r.init(), r.error
    r.debug = True      # GVT_FLAG_DEBUG_COUNTS: the tests read per-pixel step counts
    yield r
    r.cleanup()


def run_both(wgl, oracle, u, precision):
    from gravitas_b200 import _lib
    wgl.precision = precision
    wgl.resize(int(u.resolution[0]), int(u.resolution[1]))
    got = np.array(wgl.render({}, (u.mouse[0], u.mouse[1]), uniforms=u)).astype(np.float64)
    steps, hit = wgl.debug_counts()
    ref = oracle.fragment_glsl(bytes(u), wgl.noise_r, wgl.blue_r, precision=1 if precision == _lib.PRECISION_F32 else 0)
    return got, steps, hit, ref


FEATURE_SETS = {
    "high-quality": dict(),
    "balanced": dict(rayTracingQuality="medium", dopplerBeaming=False, photonSphereGlow=False, relativisticJets=False),
    "no-lensing": dict(gravitationalLensing=False),
    "redshift-overlay": dict(gravitationalRedshift=True),
    "shadow-guide": dict(kerrShadow=True, relativisticJets=False),
    "low": dict(rayTracingQuality="low"),
}


@pytest.mark.parametrize("name", list(FEATURE_SETS))
def test_f64_kernel_matches_f64_oracle(wgl, oracle, name):
    from gravitas_b200 import webgl, _lib
    W, H = 160, 90
    feats = dict(webgl.DEFAULT_FEATURES, **FEATURE_SETS[name])
    params = dict(mass=1.0, spin=0.9, zoom=30.0, lensing=1.0)
    u = webgl.make_uniforms(W, H, params, mouse=(0.5, 0.5 + 7.0 / 180.0), time=1.37, features=feats)
    if name == "shadow-guide":
        import gravitas_b200 as g
        e = g.PhysicsEngine(1.0, 0.9)
        curve = e.compute_shadow_curve(math.radians(97.0), 32)
        u = webgl.make_uniforms(W, H, params, mouse=(0.5, 0.5 + 7.0 / 180.0), time=1.37, features=feats,
                                shadow_curve=curve, shadow_count=curve.size // 2)
    got, steps, hit, ref = run_both(wgl, oracle, u, _lib.PRECISION_F64)
    assert np.all(np.isfinite(got)) and np.all(got[..., 3] == 1.0)
    same = (steps == ref["steps"]) & (hit == ref["hit"])
    assert same.mean() >= 0.999, f"{name}: step/horizon census differs on {(~same).sum()} of {same.size} pixels"
    err = np.abs(got[..., :3] - ref["rgba"][..., :3])[same]
    # frame values are stored as float32: 6e-8 relative on [0, ~2] values
    assert np.percentile(err, 99.9) <= 5e-7, f"{name}: p99.9 abs err {np.percentile(err, 99.9):.3e}"
    assert err.max() <= 2e-3, f"{name}: max abs err {err.max():.3e}"       # a star-cell / threshold flip on a rare pixel
    assert wgl.last_stats.steps_committed == int(steps.sum())
    if name not in ("low",):
        assert ref["rgba"][..., :3].max() > 0.2 and int(steps.sum()) > W * H * 10   # a real picture, a real march


def test_f32_kernel_statistics(wgl, oracle):
    from gravitas_b200 import webgl, _lib
    W, H = 192, 108
    u = webgl.make_uniforms(W, H, dict(mass=1.0, spin=0.9, zoom=30.0, lensing=1.0), mouse=(0.47, 0.5 + 7.0 / 180.0), time=0.5)
    got, steps, hit, ref = run_both(wgl, oracle, u, _lib.PRECISION_F32)
    err = np.abs(got[..., :3] - ref["rgba"][..., :3])
    assert np.median(err) <= 2e-6 and np.percentile(err, 99) <= 2e-3, (np.median(err), np.percentile(err, 99))
    assert (steps == ref["steps"]).mean() >= 0.97
    ref64 = oracle.fragment_glsl(bytes(u), wgl.noise_r, wgl.blue_r, precision=0)
    e64 = np.abs(got[..., :3] - ref64["rgba"][..., :3])
    assert np.median(e64) <= 5e-6 and np.percentile(e64, 99) <= 5e-3      # accuracy of the shader's f32 arithmetic


def test_f32_fast_math_build_statistics(wgl, oracle):
    """GVT_PRECISION_F32_FAST: the same kernel source built with MUFU approximations (what a GLSL compiler emits).
    No bit-level claim: compared with the f64 oracle as an accuracy statement on the tone-mapped [0, 1] colours."""
    from gravitas_b200 import webgl, _lib
    W, H = 192, 108
    u = webgl.make_uniforms(W, H, dict(mass=1.0, spin=0.9, zoom=30.0, lensing=1.0), mouse=(0.47, 0.5 + 7.0 / 180.0), time=0.5)
    wgl.precision = _lib.PRECISION_F32_FAST
    wgl.resize(W, H)
    got = np.array(wgl.render({}, (0.47, 0.54), uniforms=u)).astype(np.float64)
    steps, hit = wgl.debug_counts()
    ref64 = oracle.fragment_glsl(bytes(u), wgl.noise_r, wgl.blue_r, precision=0)
    e = np.abs(got[..., :3] - ref64["rgba"][..., :3])
    assert np.all(np.isfinite(got))
    assert np.median(e) <= 2e-5 and np.percentile(e, 99) <= 1e-2, (np.median(e), np.percentile(e, 99))
    assert (steps == ref64["steps"]).mean() >= 0.95 and (hit == ref64["hit"]).mean() >= 0.999
    wgl.precision = _lib.PRECISION_F32


def test_quaternion_camera_mass_scaling_and_debug(wgl, oracle):
    from gravitas_b200 import webgl, _lib
    W, H = 97, 61                                   # ragged: partial tiles on both edges
    q = (0.0, math.sin(0.1), 0.0, math.cos(0.1))    # small yaw
    u = webgl.make_uniforms(W, H, dict(mass=2.5, spin=-0.6, zoom=40.0, lensing=0.7, diskTemp=20000.0), time=3.0,
                            cam_pos=(3.0, 6.0, -70.0), cam_quat=q,
                            features=dict(webgl.DEFAULT_FEATURES, relativisticJets=True))
    got, steps, hit, ref = run_both(wgl, oracle, u, _lib.PRECISION_F64)
    same = (steps == ref["steps"]) & (hit == ref["hit"])
    assert same.mean() >= 0.999
    assert np.percentile(np.abs(got[..., :3] - ref["rgba"][..., :3])[same], 99.9) <= 5e-7
    # u_debug: the uv ramp (fragment.glsl.ts:45-48)
    ud = webgl.make_uniforms(W, H, debug=1.0)
    gd, _, _, rd = run_both(wgl, oracle, ud, _lib.PRECISION_F32)
    np.testing.assert_allclose(gd[..., :3], rd["rgba"][..., :3], atol=1e-6)
    assert abs(gd[H // 2, W // 2, 0] - 0.5) < 0.02 and abs(gd[H // 2, W // 2, 1] - 0.5) < 0.02


def test_webgl_pipeline_with_taa_and_formats(wgl, oracle):
    """render(params, mouse) twice with the TAA resolve on (as the reference with a ReprojectionManager): frame 2 ==
    reprojection.glsl resolve of (shader frame 2, frame 1); RGBA8 read-back of a linear-output frame."""
    import taa_oracle
    from gravitas_b200 import webgl, _lib
    W, H = 128, 72
    wgl.precision = _lib.PRECISION_F32
    wgl.resize(W, H)
    wgl._k.reset_history()
    wgl.taa = True
    wgl.time = 0.0
    params = dict(mass=1.0, spin=0.9, zoom=30.0, lensing=1.0, features=dict(webgl.DEFAULT_FEATURES, bloom=False))
    mouse = {"x": 0.5, "y": 0.54}
    try:
        f1 = np.array(wgl.render(params, mouse))
        assert wgl.last_stats.kernel_launches == 2 and wgl.last_stats.taa_ms > 0
        f2 = np.array(wgl.render(params, mouse))
        # un-resolved frame 2 from a second renderer with the same textures and time
        plain = webgl.WebGLRenderer(device=0, noise_seed=11)
        assert plain.init()
        plain.resize(W, H)
        u2 = webgl.make_uniforms(W, H, params, (0.5, 0.54), time=0.02, features=params["features"], has_post=True)
        cur2 = np.array(plain.render({}, (0.5, 0.54), uniforms=u2))
        plain.cleanup()
        ref2 = taa_oracle.taa_resolve_webgl(cur2, f1, 0.75, False)
        scale = float(np.abs(ref2[..., :3]).max())
        np.testing.assert_allclose(f2, ref2, rtol=1e-3, atol=1e-3 * scale)
    finally:
        wgl.taa = False
    # 8-bit display output of a tone-mapped frame
    u = webgl.make_uniforms(W, H, params, (0.5, 0.54), time=0.3)
    f32 = np.array(wgl.render({}, (0.5, 0.54), uniforms=u))
    f16 = wgl.read_frame(_lib.FORMAT_RGBA16F)
    np.testing.assert_allclose(f16.astype(np.float32), f32, rtol=1e-3, atol=1e-3)
    u8 = wgl.read_frame(_lib.FORMAT_RGBA8_UNORM)                       # already ACES + gamma: quantise only
    assert np.abs(u8.astype(np.float32) - np.clip(f32, 0, 1) * 255.0).max() <= 0.5 + 1e-3
    got8 = np.array(wgl.render({}, (0.5, 0.54), uniforms=u, output_format=_lib.FORMAT_RGBA8_UNORM))
    assert got8.dtype == np.uint8 and np.array_equal(got8, u8)


def test_fragment_validation(wgl):
    from gravitas_b200 import webgl, _lib
    import ctypes as C
    u = webgl.make_uniforms(64, 36)
    u.struct_size = 12
    with pytest.raises(_lib.GravitasError):
        wgl.render({}, (0.5, 0.5), uniforms=u)
    u = webgl.make_uniforms(64, 36)
    with pytest.raises(_lib.GravitasError):          # TAA without the WebGL variant has no camera to reproject with
        wgl.render({}, (0.5, 0.5), uniforms=u, flags=_lib.FLAG_TAA)
    with pytest.raises(_lib.GravitasError):
        _lib.check(_lib.lib().gvt_render_set_noise_textures(wgl._k._h, None, None, 256))


@pytest.mark.parametrize("case", ["hq", "guide"])
def test_f64_kernel_matches_committed_fixture(built, case):
    """The CUDA kernel against tests/golden/glsl_fragment_48x27.npz (oracle output committed with its generator)."""
    import os
    from gravitas_b200 import webgl, _lib
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "glsl_fragment_48x27.npz")))
    r = webgl.WebGLRenderer(device=0, noise_seed=int(g["noise_seed"]))
    assert r.init(), r.error
    r.debug = True
    try:
        u = _lib.GvtGlslUniforms.from_buffer_copy(g[f"{case}_uniforms"].tobytes())
        r.precision = _lib.PRECISION_F64
        r.resize(48, 27)
        got = np.array(r.render({}, (u.mouse[0], u.mouse[1]), uniforms=u)).astype(np.float64)
        steps, hit = r.debug_counts()
        assert np.array_equal(steps, g[f"{case}_steps"]) and np.array_equal(hit, g[f"{case}_hit"])
        np.testing.assert_allclose(got, g[f"{case}_rgba"], rtol=0, atol=2e-7)   # float32 frame buffer
    finally:
        r.cleanup()


@pytest.mark.parametrize("W,H,passes", [(128, 72, 2), (157, 83, 1), (64, 36, 0), (5, 3, 2)])
def test_bloom_matches_numpy_restatement(wgl, W, H, passes):
    """gvt_render_bloom (bright pass, RGBA16F quarter-res Gaussian ping-pong, combine + ACES + gamma) vs the numpy
    restatement of bloom.glsl.ts / bloom.ts. Tolerance: intermediates are RGBA16F, so an FMA-vs-separate rounding
    difference can flip a half-float ulp (2^-11 relative) in a blurred texel; the output is display-referred [0, 1]."""
    import bloom_oracle
    from gravitas_b200 import webgl, _lib
    wgl.precision = _lib.PRECISION_F32
    wgl.taa = False
    wgl.resize(W, H)
    feats = dict(webgl.PRESETS["ultra-quality"], bloom=True)
    u = webgl.make_uniforms(W, H, dict(mass=1.0, spin=0.9, zoom=30.0, lensing=1.0), (0.5, 0.54), time=0.4, features=feats, has_post=True)
    scene = np.array(wgl.render({}, (0.5, 0.54), uniforms=u))            # linear HDR (ENABLE_LINEAR_OUTPUT); stays the frame
    got_plain = wgl._k.bloom(enabled=False)                               # drawTextureToScreen: ACES + gamma only
    ref_plain = bloom_oracle.apply_bloom(scene, enabled=False)
    np.testing.assert_allclose(got_plain, ref_plain, atol=5e-6)          # production build: MUFU lg2/ex2 gamma vs numpy's powf
    got = wgl._k.bloom(enabled=True, intensity=0.5, threshold=0.05, blur_passes=passes)   # low threshold: most of the disk blooms
    ref = bloom_oracle.apply_bloom(scene, True, 0.5, 0.05, passes)
    # The PRECISE build (IEEE f32 operations in GLSL order, libm powf) is held to 1e-6 relative (floor 1e-3 x peak; the
    # frame is display-referred, peak <= 1): the RGBA16F intermediates then round identically in kernel and oracle.
    for enabled, r_ in ((False, ref_plain), (True, ref)):
        got_p = wgl._k.bloom(enabled=enabled, intensity=0.5, threshold=0.05, blur_passes=passes, precise=True)
        peak = float(np.abs(r_[..., :3]).max())
        e = np.abs(got_p.astype(np.float64) - r_) / np.maximum(np.abs(r_), 1e-3 * max(peak, 1e-30))
        print(f"bloom {W}x{H} passes {passes} enabled {enabled}: precise build vs numpy max rel err {e.max():.3e}")
        assert e.max() <= 1e-6, (enabled, float(e.max()))
        if enabled:
            d = np.abs(got - got_p)
            print(f"   production vs precise build (FMA contraction + MUFU; an RGBA16F ulp flip is 2^-11 relative): "
                  f"median {np.median(d):.2e} max {d.max():.2e}")
    # the production build only: an FMA-vs-separate rounding difference can flip a half-float ulp in a blurred texel
    assert np.abs(got - ref).max() <= 2e-3 and np.median(np.abs(got - ref)) <= 2e-6, (np.abs(got - ref).max(), passes)
    assert np.abs(ref - ref_plain).max() > 0.01 or W < 16                 # the bloom actually contributes
    u8 = wgl._k.bloom(enabled=True, intensity=0.5, threshold=0.05, blur_passes=passes, fmt=_lib.FORMAT_RGBA8_UNORM)
    assert u8.dtype == np.uint8 and np.abs(u8.astype(np.float32) - got * 255.0).max() <= 0.5 + 1e-3
    # present(): the renderer's post tail picks bloom or the plain draw from features.bloom
    out = wgl.present(dict(features=dict(feats, bloom=False)), fmt=_lib.FORMAT_RGBA32F)
    np.testing.assert_allclose(out, got_plain, atol=0)


@pytest.mark.parametrize("W,H", [(1, 1), (7, 3), (33, 2), (8, 4)])
def test_fragment_degenerate_sizes_and_budgets(wgl, oracle, W, H):
    """Edge cases: frames smaller than one 8x4 warp tile, a zero / one-step ray budget, a camera beyond MAX_DIST (every
    ray leaves on its first test), a camera inside 1.5 r+ (the shader's kamikaze clamp), a near-extremal and a
    retrograde spin — all against the f64 oracle."""
    from gravitas_b200 import webgl, _lib
    feats = dict(webgl.PRESETS["ultra-quality"], bloom=False)
    cases = [dict(params=dict(spin=0.9, zoom=30.0, lensing=1.0), steps=None),
             dict(params=dict(spin=0.9, zoom=30.0, lensing=1.0), steps=0),
             dict(params=dict(spin=0.9, zoom=30.0, lensing=1.0), steps=1),
             dict(params=dict(spin=0.5, zoom=6000.0, lensing=1.0), steps=None),      # ro at 12000 > MAX_DIST
             dict(params=dict(spin=0.5, zoom=1.0, lensing=1.0), steps=None),         # ro at 2 M < 1.5 r+ ~ 2.8 M
             dict(params=dict(spin=0.999, zoom=20.0, lensing=2.0, mass=0.5), steps=None),
             dict(params=dict(spin=-0.99, zoom=20.0, lensing=0.3, mass=3.0), steps=None)]
    for c in cases:
        u = webgl.make_uniforms(W, H, c["params"], mouse=(0.37, 0.58), time=2.0, features=feats)
        if c["steps"] is not None:
            u.max_ray_steps = c["steps"]
        got, steps, hit, ref = run_both(wgl, oracle, u, _lib.PRECISION_F64)
        assert np.all(np.isfinite(got)), c
        assert np.array_equal(steps, ref["steps"]) and np.array_equal(hit, ref["hit"]), c
        np.testing.assert_allclose(got[..., :3], ref["rgba"][..., :3], rtol=0, atol=2e-7)
        if c["steps"] is not None:
            assert int(steps.max()) <= c["steps"]
    if (W, H) == (8, 4):
        assert wgl.last_stats.n_horizon + wgl.last_stats.n_escape == W * H


def test_both_seams_webgl_pipeline(built, oracle):
    """Seam A feeds Seam B exactly as in the reference's WebGL canvas: PhysicsEngine.tick_sab publishes the Bardeen
    curve in the SAB PHYSICS block -> WebGLRenderer.render() uploads it as u_shadowCurve / u_shadowCount
    (webgl/renderer.ts:277-303) -> the shader's Kerr-shadow guide draws it; frame checked against the oracle fed the
    same uniform block, then TAA + bloom tail."""
    import bloom_oracle
    import gravitas_b200 as g
    from gravitas_b200 import webgl, _lib
    W, H = 144, 81
    eng = g.PhysicsEngine(1.0, 0.9)
    eng.set_camera_state(0.0, 60.0 * math.cos(math.radians(97.0)), -60.0 * math.sin(math.radians(97.0)))
    r = webgl.WebGLRenderer(device=0, noise_seed=2)
    assert r.init(), r.error
    try:
        r.physics_bridge = eng
        r.precision = _lib.PRECISION_F64
        r.resize(W, H)
        params = dict(mass=1.0, spin=0.9, zoom=30.0, lensing=1.0,
                      features=dict(webgl.PRESETS["high-quality"], kerrShadow=True, bloom=False))
        u = r.uniforms(params, (0.5, 0.5 + 7.0 / 180.0))
        sab = eng.get_sab_ptr()
        assert u.shadow_count == sab[128 + 15] == 64.0
        np.testing.assert_array_equal(np.array(u.shadow_curve[:]), sab[128 + 16:128 + 16 + 128])
        got = np.array(r.render(params, (0.5, 0.5 + 7.0 / 180.0), uniforms=u)).astype(np.float64)
        ref = oracle.fragment_glsl(bytes(u), r.noise_r, r.blue_r, precision=0)
        np.testing.assert_allclose(got[..., :3], ref["rgba"][..., :3], rtol=0, atol=2e-7)
        u_off = _lib.GvtGlslUniforms.from_buffer_copy(bytes(u))
        u_off.show_kerr_shadow = 0.0
        plain = oracle.fragment_glsl(bytes(u_off), r.noise_r, r.blue_r, precision=0)
        drawn = np.abs(ref["rgba"][..., :3] - plain["rgba"][..., :3]).max(-1) > 1e-3
        # the guide is 0.045 M thick and a pixel spans ~0.5 M here, so it touches only a few pixel centres
        assert 1 <= drawn.sum() < W * H // 10
        # linear-HDR frame -> bloom + final pass, as render() does when features.bloom (renderer.ts:366-399)
        u2 = r.uniforms(dict(params, features=dict(params["features"], bloom=True)), (0.5, 0.54), has_post=True)
        lin = np.array(r.render(params, (0.5, 0.54), uniforms=u2))
        out = r.present(dict(features=dict(bloom=True)), fmt=_lib.FORMAT_RGBA32F)
        refb = bloom_oracle.apply_bloom(lin, True, 0.5, 0.8, 2)
        assert np.abs(out - refb).max() <= 2e-3
    finally:
        r.cleanup()


def test_fragment_shader_full_4k_every_pixel(wgl, oracle):
    """Every pixel of a full 3840x2160 frame of the fragment shader (ultra-quality preset, all features) in f64 against
    the f64 oracle (~12 s of oracle time on 16 host threads): identical step counts and horizon flags everywhere except
    at most a handful of discontinuity-sitting pixels, colours to the float32 frame-buffer resolution."""
    from gravitas_b200 import webgl, _lib
    W, H = 3840, 2160
    feats = dict(webgl.PRESETS["ultra-quality"], bloom=False)
    u = webgl.make_uniforms(W, H, dict(mass=1.0, spin=0.9, zoom=30.0, lensing=1.0), mouse=(0.5, 0.5 + 7.0 / 180.0), time=0.25,
                            features=feats)
    got, steps, hit, ref = run_both(wgl, oracle, u, _lib.PRECISION_F64)
    same = (steps == ref["steps"]) & (hit == ref["hit"])
    n_bad = int((~same).sum())
    assert n_bad <= 8, f"step/horizon census differs on {n_bad} of {W * H} pixels"
    err = np.abs(got[..., :3] - ref["rgba"][..., :3])
    print(f"[4K fragment f64] census differs on {n_bad} pixels; max colour err {err.max():.3e} (same-census pixels "
          f"{err[same].max():.3e}); pixels > 1e-6: {int((err.max(-1) > 1e-6).sum())}; steps {int(steps.sum())}; "
          f"oracle {ref['seconds']:.1f} s")
    assert err[same].max() <= 1e-6, err[same].max()
    # a star cell / threshold can flip on a pixel whose step count is unchanged: allow a handful of larger outliers
    assert int((err.max(-1) > 1e-6).sum()) <= n_bad + 8
    assert wgl.last_stats.steps_committed == int(ref["total_steps"]) or n_bad > 0
    wgl.resize(64, 36)


def test_fragment_randomised_uniforms(wgl, oracle):
    """Seeded random uniform blocks (mass, spin of either sign, zoom, both cameras, time, disk parameters, lensing strength,
    random feature subsets) on small frames: the f64 kernel reproduces the f64 oracle's per-pixel step counts / horizon
    flags and colours on every one of them."""
    from gravitas_b200 import webgl, _lib
    rng = np.random.default_rng(20260117)
    W, H = 40, 24
    keys = ["gravitationalLensing", "accretionDisk", "dopplerBeaming", "backgroundStars", "photonSphereGlow",
            "relativisticJets", "gravitationalRedshift", "kerrShadow"]
    n_bad = 0
    for case in range(24):
        feats = dict(webgl.PRESETS["ultra-quality"], bloom=False)
        for k in keys:
            feats[k] = bool(rng.random() < (0.8 if k in ("gravitationalLensing", "accretionDisk") else 0.5))
        feats["rayTracingQuality"] = str(rng.choice(["medium", "high", "ultra"]))
        mass = float(rng.uniform(0.3, 4.0))
        params = dict(mass=mass, spin=float(rng.uniform(-0.99, 0.99)), zoom=float(rng.uniform(6.0, 60.0)) * mass ** 0.5,
                      lensing=float(rng.uniform(0.2, 1.6)), diskDensity=float(rng.uniform(0.5, 5.0)),
                      diskTemp=float(rng.uniform(2000.0, 60000.0)), diskSize=float(rng.uniform(8.0, 80.0)),
                      diskScaleHeight=float(rng.uniform(0.02, 0.3)))
        kw = {}
        if rng.random() < 0.4:      # quaternion camera somewhere around the hole, looking roughly at it
            d = float(rng.uniform(15.0, 80.0)) * mass
            az, el = float(rng.uniform(0, 2 * math.pi)), float(rng.uniform(-0.6, 0.6))
            kw["cam_pos"] = (d * math.cos(el) * math.sin(az), d * math.sin(el), -d * math.cos(el) * math.cos(az))
            half = -0.5 * az
            kw["cam_quat"] = (0.0, math.sin(half), 0.0, math.cos(half))
        u = webgl.make_uniforms(W, H, params, mouse=(float(rng.random()), float(rng.uniform(0.15, 0.85))),
                                time=float(rng.uniform(0.0, 50.0)), features=feats, **kw)
        got, steps, hit, ref = run_both(wgl, oracle, u, _lib.PRECISION_F64)
        assert np.all(np.isfinite(got)), (case, params)
        same = (steps == ref["steps"]) & (hit == ref["hit"])
        n_bad += int((~same).sum())
        err = np.abs(got[..., :3] - ref["rgba"][..., :3])
        # the display-referred [0,1] output agrees to the float32 frame buffer; linear-HDR values (none here) would scale
        assert err[same].max() <= 3e-7 * max(1.0, float(np.abs(ref["rgba"][..., :3]).max())), (case, err[same].max(), params, feats)
    assert n_bad <= 2, n_bad
