"""WebGLRenderer — the WebGL2 pipeline's renderer API (src/rendering/webgl/renderer.ts:36-482) over
gvt_render_fragment_glsl: ``init(canvas)`` -> ``resize(w, h)`` -> ``render(params, mouse)`` -> ``cleanup()``.

``render`` runs the production fragment shader (src/shaders/blackhole/fragment.glsl.ts) as one fused CUDA kernel and,
like the reference when its ReprojectionManager exists, the WebGL2 TAA resolve (reprojection.glsl.ts); bloom is out of
scope, so the frame is returned instead of being composited to a canvas. The two noise textures the reference fills
with ``Math.random()`` (utils/webgl-utils.ts:259-303) are an explicit, seedable input here."""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import GvtFrameStats, GvtGlslUniforms, check, lib
from .renderer import KerrRenderer, _stats

# configs/simulation.config.ts:205-211
RAY_TRACING_STEPS = {"off": 0, "low": 32, "medium": 64, "high": 128, "ultra": 256}
# configs/simulation.config.ts:21-76 PERFORMANCE_PRESETS
PRESETS = {
    "maximum-performance": dict(gravitationalLensing=False, rayTracingQuality="off", accretionDisk=False, dopplerBeaming=False,
                                backgroundStars=False, photonSphereGlow=False, bloom=False, relativisticJets=False,
                                gravitationalRedshift=False, kerrShadow=False),
    "balanced": dict(gravitationalLensing=True, rayTracingQuality="medium", accretionDisk=True, dopplerBeaming=False,
                     backgroundStars=True, photonSphereGlow=False, bloom=False, relativisticJets=False,
                     gravitationalRedshift=False, kerrShadow=False),
    "high-quality": dict(gravitationalLensing=True, rayTracingQuality="high", accretionDisk=True, dopplerBeaming=True,
                         backgroundStars=True, photonSphereGlow=True, bloom=True, relativisticJets=True,
                         gravitationalRedshift=False, kerrShadow=False),
    "ultra-quality": dict(gravitationalLensing=True, rayTracingQuality="ultra", accretionDisk=True, dopplerBeaming=True,
                          backgroundStars=True, photonSphereGlow=True, bloom=True, relativisticJets=True,
                          gravitationalRedshift=False, kerrShadow=False),
}
DEFAULT_FEATURES = PRESETS["high-quality"]      # configs/simulation.config.ts:77-78,200-202
# types/simulation.ts:31-44 with the defaults of configs/simulation.config.ts:82-181
DEFAULT_PARAMS = dict(mass=1.0, spin=0.5, diskDensity=4.0, diskTemp=9500.0, lensing=0.7, paused=False, zoom=30.0,
                      diskSize=50.0, diskScaleHeight=0.2)


def feature_bits(features, has_post=False):
    """shaders/manager.ts:55-82: feature toggles -> #defines (here: GVT_GLSL_* bits)."""
    f = dict(DEFAULT_FEATURES, **(features or {}))
    bits = 0
    if f["gravitationalLensing"]: bits |= _lib.GLSL_LENSING
    if f["accretionDisk"]: bits |= _lib.GLSL_DISK
    if f["dopplerBeaming"]: bits |= _lib.GLSL_DOPPLER
    if f["backgroundStars"]: bits |= _lib.GLSL_STARS
    if f["photonSphereGlow"]: bits |= _lib.GLSL_PHOTON_GLOW
    if f["relativisticJets"] and f["accretionDisk"]: bits |= _lib.GLSL_JETS
    if f["gravitationalRedshift"]: bits |= _lib.GLSL_REDSHIFT
    if f["rayTracingQuality"] in ("low", "off"): bits |= _lib.GLSL_QUALITY_LOW
    if has_post: bits |= _lib.GLSL_LINEAR_OUTPUT
    return bits


def make_uniforms(width, height, params=None, mouse=(0.5, 0.5), time=0.0, features=None, shadow_curve=None,
                  shadow_count=None, has_post=False, cam_pos=(0.0, 0.0, 0.0), cam_quat=(0.0, 0.0, 0.0, 1.0), debug=0.0):
    """The uniform block exactly as WebGLRenderer.render uploads it (webgl/renderer.ts:277-358)."""
    p = dict(DEFAULT_PARAMS, **(params or {}))
    f = dict(DEFAULT_FEATURES, **(features or {}))
    u = GvtGlslUniforms()
    u.struct_size = C.sizeof(u)
    u.features = feature_bits(f, has_post)
    u.resolution[0], u.resolution[1] = float(width), float(height)
    u.time = time
    u.mass = p["mass"]
    u.spin = p["spin"] * p["mass"]                       # renderer.ts:326
    u.zoom = p["zoom"] * 2.0                             # renderer.ts:327
    u.disk_size, u.disk_scale_height = p["diskSize"], p["diskScaleHeight"]
    u.max_ray_steps = RAY_TRACING_STEPS.get(f["rayTracingQuality"], 250)
    u.show_redshift = 1.0 if f["gravitationalRedshift"] else 0.0
    u.show_kerr_shadow = 1.0 if f["kerrShadow"] else 0.0
    u.lensing_strength = p["lensing"]
    u.mouse[0], u.mouse[1] = mouse
    u.disk_density = p["diskDensity"]
    u.disk_temp = p["diskTemp"] * math.pow(p["mass"], -0.25)   # renderer.ts:351-354
    u.debug = debug
    for i in range(3):
        u.cam_pos[i] = cam_pos[i]
    for i in range(4):
        u.cam_quat[i] = cam_quat[i]
    if shadow_curve is None or (shadow_count is not None and shadow_count <= 0):
        # default Schwarzschild fallback when telemetry is missing (renderer.ts:289-298)
        b = 3.0 * math.sqrt(3.0) * p["mass"]
        shadow_count = 64
        shadow_curve = np.zeros(128, np.float32)
        for i in range(64):
            phi = (i / 64) * math.pi * 2.0
            shadow_curve[2 * i], shadow_curve[2 * i + 1] = math.cos(phi) * b, math.sin(phi) * b
    sc = np.zeros(128, np.float32)
    flat = np.asarray(shadow_curve, np.float32).ravel()[:128]
    sc[:flat.size] = flat
    for i in range(128):
        u.shadow_curve[i] = sc[i]
    u.shadow_count = float(shadow_count if shadow_count is not None else flat.size // 2)
    return u


def random_noise_textures(seed=0, size=256):
    """createNoiseTexture / createBlueNoiseTexture (webgl-utils.ts:259-303): floor(random * 255) per RGBA byte."""
    rng = np.random.default_rng(seed)
    noise = np.floor(rng.random((size, size, 4)) * 255).astype(np.uint8)
    blue = np.floor(rng.random((size, size, 4)) * 255).astype(np.uint8)
    return noise, blue


class WebGLRenderer:
    def __init__(self, device=0, rank=0, world_size=1, nccl_id=None, noise_seed=0):
        self._k = KerrRenderer(device=device, rank=rank, world_size=world_size, nccl_id=nccl_id)
        self.error = None                       # webgl/renderer.ts:38
        self.on_metrics_update = None           # webgl/renderer.ts:39
        self.width = self.height = 0
        self.time = 0.0
        self.precision = _lib.PRECISION_F32
        self.taa = False                        # the reference resolves whenever its ReprojectionManager exists
        self.last_mouse = (0.0, 0.0)
        self.last_stats = None
        self._noise_seed = noise_seed
        self.debug = False                      # record per-pixel steps / horizon flags (parity hook, debug_counts())
        self.physics_bridge = None              # optional PhysicsEngine: supplies the SAB shadow curve (renderer.ts:277-287)

    def init(self, canvas=None):                # renderer.ts:58-161
        try:
            self._k.init()
            noise, blue = random_noise_textures(self._noise_seed)
            self.set_noise_textures(noise, blue)
        except _lib.GravitasError as e:
            self.error = str(e)
            return False
        return True

    def set_noise_textures(self, noise_rgba8, blue_rgba8):
        noise_rgba8 = np.ascontiguousarray(noise_rgba8, np.uint8)
        blue_rgba8 = np.ascontiguousarray(blue_rgba8, np.uint8)
        if noise_rgba8.shape != (256, 256, 4) or blue_rgba8.shape != (256, 256, 4):
            raise ValueError("noise textures are (256, 256, 4) uint8")
        pu8 = C.POINTER(C.c_uint8)
        check(lib().gvt_render_set_noise_textures(self._k._h, noise_rgba8.ctypes.data_as(pu8), blue_rgba8.ctypes.data_as(pu8), 256))
        self.noise_r, self.blue_r = noise_rgba8[..., 0].copy(), blue_rgba8[..., 0].copy()

    def resize(self, width, height):            # renderer.ts:163-171
        if self.width == width and self.height == height:
            return
        self._k.resize(width, height)
        self.width, self.height = int(width), int(height)

    def uniforms(self, params, mouse, has_post=False):
        curve, count = None, None
        if self.physics_bridge is not None:      # renderer.ts:277-287: telemetry.physics[15] > 0 -> SAB curve
            self.physics_bridge.tick_sab(0.016)
            phys = self.physics_bridge.get_sab_ptr()[128:256 + 16]
            if phys[15] > 0.0:
                curve, count = phys[16:16 + 128].copy(), float(phys[15])
        return make_uniforms(self.width, self.height, params, (mouse["x"], mouse["y"]) if isinstance(mouse, dict) else mouse,
                             self.time, (params or {}).get("features"), curve, count, has_post)

    def render(self, params, mouse, readback=True, flags=0, output_format=_lib.FORMAT_RGBA32F, uniforms=None):
        """render(params: SimulationParams, mouse: {x, y}) (renderer.ts:173). Returns the (H, W, 4) frame, row 0 =
        gl_FragCoord.y 0.5 (bottom-up, as gl.readPixels delivers it)."""
        params = params or {}
        m = (mouse["x"], mouse["y"]) if isinstance(mouse, dict) else tuple(mouse)
        moving = abs(m[0] - self.last_mouse[0]) > 1e-4 or abs(m[1] - self.last_mouse[1]) > 1e-4   # renderer.ts:178-187
        self.last_mouse = m
        if not params.get("paused", False):
            self.time += 0.01                    # renderer.ts:192
        u = uniforms if uniforms is not None else self.uniforms(params, m, has_post=self.taa)
        W, H = int(u.resolution[0]), int(u.resolution[1])
        st = GvtFrameStats()
        if self.taa:
            flags |= _lib.FLAG_TAA | _lib.FLAG_TAA_WEBGL
        if self.debug:
            flags |= _lib.FLAG_DEBUG_COUNTS
        host = None
        if readback:
            buf = self._k.pinned_frame(W, H, output_format)
            host = buf.ptr
        check(lib().gvt_render_fragment_glsl(self._k._h, C.byref(u), self.precision, flags, output_format, 0.75,
                                             1 if moving else 0, host, C.byref(st)))
        self.width, self.height = W, H
        self._k.width, self._k.height = W, H
        self.last_stats = _stats(st)
        if self.on_metrics_update is not None:
            self.on_metrics_update(self.last_stats)
        if not readback:
            return None
        return buf.array(np.dtype(_lib.FORMAT_DTYPE[output_format]), (H, W, 4))

    def present(self, params=None, fmt=_lib.FORMAT_RGBA8_UNORM, readback=True):
        """The post-processing tail of render() (renderer.ts:366-414) once a linear-HDR frame exists: bloom when
        ``features.bloom`` (BloomManager.applyBloomToTexture), else the plain ACES + gamma draw (drawTextureToScreen)."""
        f = dict(DEFAULT_FEATURES, **((params or {}).get("features") or {}))
        return self._k.bloom(enabled=bool(f["bloom"]), fmt=fmt, readback=readback)

    def debug_counts(self):
        """Parity hook: per-pixel march steps and horizon flags of the last frame."""
        steps = np.zeros((self.height, self.width), np.uint32)
        hit = np.zeros((self.height, self.width), np.uint32)
        pu = C.POINTER(C.c_uint32)
        check(lib().gvt_render_fragment_glsl_debug(self._k._h, steps.ctypes.data_as(pu), hit.ctypes.data_as(pu)))
        return steps, hit

    def read_frame(self, fmt=_lib.FORMAT_RGBA32F):
        return self._k.read_frame(fmt)

    def cleanup(self):                          # renderer.ts:471-481
        self._k.cleanup()
