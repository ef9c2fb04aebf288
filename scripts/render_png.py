"""Render one frame of each seam-B path to PNG (no imaging library needed): the Kerr-Schild geodesic trace
(KerrRenderer, config-3 scheme, ACES 8-bit output) and the production WebGL2 fragment shader (WebGLRenderer,
ultra-quality preset + bloom). Usage: python scripts/render_png.py [out_dir] [width] [height]"""
import math, os, struct, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blackhole-simulation_b200"))
import numpy as np
import gravitas_b200 as g
from gravitas_b200 import _lib, camera, renderer as R, webgl


def write_png(path, rgb8):
    h, w, _ = rgb8.shape
    raw = b"".join(b"\x00" + rgb8[y].tobytes() for y in range(h))
    chunk = lambda t, d: struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 9)) + chunk(b"IEND", b""))


out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 960
H = int(sys.argv[3]) if len(sys.argv) > 3 else 540
os.makedirs(out, exist_ok=True)
spin = float(np.float32(0.999))
r = g.KerrRenderer()
r.init()
r.init_pipelines(mass=1.0, spin=spin, spec_w=256, spec_h=32, max_temp=1e7)
r.params = R.RenderParams(max_steps=512)
cam, _ = camera.default_camera(W, H)
hdr = np.array(r.render(cam, R.pack_physics(1.0, spin, W, H)))[..., :3]
# the frame is linear radiance in the spectral LUT's units (spectrum.rs:86 scales texels by 1e-14 g^4): expose to the
# 99.5th percentile of the lit pixels, then Reinhard + gamma as the WebGPU blit does (webgpu/renderer.ts:45-47)
lit = hdr[hdr.sum(-1) > 0]
v = hdr * (4.0 / max(float(np.percentile(lit, 99.5)), 1e-30)) if lit.size else hdr
img = (np.clip((v / (1.0 + v)) ** (1.0 / 2.2), 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)
write_png(os.path.join(out, "kerr_trace_a0999.png"), np.ascontiguousarray(img))
print(f"kerr trace {W}x{H}x512 f64: {r.last_stats.trace_ms:.2f} ms")
r.cleanup()
w = webgl.WebGLRenderer(noise_seed=11)
assert w.init(), w.error
w.precision = _lib.PRECISION_F32_FAST
w.resize(W, H)
params = dict(mass=1.0, spin=0.9, zoom=30.0, lensing=1.0, features=dict(webgl.PRESETS["ultra-quality"], bloom=True))
u = w.uniforms(params, (0.5, 0.5 + 7.0 / 180.0), has_post=True)
u.time = 1.0
w.render(params, (0.5, 0.54), uniforms=u, readback=False)
img = w.present(params, fmt=_lib.FORMAT_RGBA8_UNORM)
write_png(os.path.join(out, "webgl_fragment_a09.png"), np.ascontiguousarray(img[::-1, :, :3]))   # rows are bottom-up
print(f"fragment shader {W}x{H} ultra + bloom: {w.last_stats.trace_ms:.2f} ms + {w._k.last_bloom_ms:.3f} ms")
w.cleanup()
