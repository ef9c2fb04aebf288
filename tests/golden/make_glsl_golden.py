"""Generates tests/golden/glsl_fragment_48x27.npz: the f64 GLSL-shader oracle (oracle/glsl_fragment_oracle.hpp) on a
48x27 frame for two feature sets, with the exact uniform bytes and noise textures used. Regression pin of the oracle
itself (the reference's shader cannot run here: GLSL needs a browser + GPU), and a full-feature fixture the CUDA kernel
is checked against on the GPU box. Run from the repo root: python tests/golden/make_glsl_golden.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "blackhole-simulation_b200"))
import numpy as np
import oracle as O
from gravitas_b200 import webgl

W, H = 48, 27
noise, blue = webgl.random_noise_textures(seed=5)
nr, br = noise[..., 0].copy(), blue[..., 0].copy()
# the textures are regenerated from the seed (numpy PCG64 streams are stable across versions); a CRC pins them
import zlib
out = {"noise_seed": np.array(5), "noise_crc": np.array([zlib.crc32(nr.tobytes()), zlib.crc32(br.tobytes())], np.uint64)}
cases = {
    "hq": webgl.make_uniforms(W, H, dict(mass=1.0, spin=0.9, zoom=30.0, lensing=1.0), mouse=(0.5, 0.5 + 7.0 / 180.0), time=1.25,
                              features=dict(webgl.PRESETS["ultra-quality"], bloom=False)),
    "guide": webgl.make_uniforms(W, H, dict(mass=2.0, spin=-0.5, zoom=45.0, lensing=0.7), mouse=(0.31, 0.62), time=4.0,
                                 features=dict(webgl.PRESETS["balanced"], kerrShadow=True)),
}
for name, u in cases.items():
    r = O.fragment_glsl(bytes(u), nr, br, precision=0)
    out[f"{name}_uniforms"] = np.frombuffer(bytes(u), np.uint8).copy()
    out[f"{name}_rgba"] = r["rgba"]
    out[f"{name}_steps"] = r["steps"]
    out[f"{name}_hit"] = r["hit"]
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "glsl_fragment_48x27.npz"), **out)
print({k: (v.shape, v.dtype) for k, v in out.items()})
