"""A few 4K frames of the WebGL2 fragment-shader kernel (k_fragment_glsl) — the command ncu wraps, and the parity
statistics quoted in DESIGN.md. Args: [frames] [preset] [precision fast|f32|f64] [parity]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blackhole-simulation_b200"))
import numpy as np
import gravitas_b200 as g
from gravitas_b200 import webgl, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
preset = sys.argv[2] if len(sys.argv) > 2 else "ultra-quality"
prec = {"f64": _lib.PRECISION_F64, "f32": _lib.PRECISION_F32, "fast": _lib.PRECISION_F32_FAST}[sys.argv[3] if len(sys.argv) > 3 else "fast"]
W, H = 3840, 2160
r = webgl.WebGLRenderer(noise_seed=11)
assert r.init(), r.error
r.precision = prec
r.resize(W, H)
feats = dict(webgl.PRESETS[preset], bloom=False)
params = dict(mass=1.0, spin=0.9, zoom=30.0, lensing=1.0, features=feats)
ts = []
for k in range(n):
    r.render(params, {"x": 0.5, "y": 0.5 + 7.0 / 180.0}, readback=False)
    s = r.last_stats
    ts.append(s.trace_ms)
    print(f"frame {k}: shader {s.trace_ms:.3f} ms, steps {s.steps_committed} ({s.steps_committed / (W * H):.1f}/pixel), "
          f"{s.steps_committed / s.trace_ms * 1e3:.4e} steps/s, horizon pixels {s.n_horizon}")
print(f"{preset} {['f64', 'f32', 'f32-fast'][prec]} 4K: median {sorted(ts)[len(ts) // 2]:.3f} ms")
if len(sys.argv) > 4 and sys.argv[4] == "parity":
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    Wp, Hp = 480, 270
    r.resize(Wp, Hp)
    r.debug = True
    for pr, nm in ((_lib.PRECISION_F64, "f64"), (_lib.PRECISION_F32, "f32"), (_lib.PRECISION_F32_FAST, "f32-fast")):
        r.precision = pr
        u = webgl.make_uniforms(Wp, Hp, params, (0.5, 0.5 + 7.0 / 180.0), time=0.77, features=feats)
        got = np.array(r.render({}, (0.5, 0.54), uniforms=u)).astype(np.float64)
        steps, hit = r.debug_counts()
        ref = O.fragment_glsl(bytes(u), r.noise_r, r.blue_r, precision=0 if pr == 0 else 1)
        same = (steps == ref["steps"]) & (hit == ref["hit"])
        e = np.abs(got[..., :3] - ref["rgba"][..., :3])
        print(f"parity {nm} {Wp}x{Hp}: identical step/horizon census on {same.mean() * 100:.3f} % of pixels; abs colour err "
              f"median {np.median(e):.2e} p99 {np.percentile(e, 99):.2e} p99.9 {np.percentile(e, 99.9):.2e} max {e.max():.2e}; "
              f"oracle {ref['seconds']:.2f} s ({ref['total_steps'] / ref['seconds']:.3e} steps/s on {O.lib().orc_num_threads()} threads)")
