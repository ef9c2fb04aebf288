"""GPU-box diagnostic: per-pixel comparison of gvt_trace_states with the oracle; dumps the worst pixels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "blackhole-simulation_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import gravitas_b200 as g
from gravitas_b200 import camera, renderer as R, _lib
import oracle as O

W, H, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
method = int(sys.argv[4]) if len(sys.argv) > 4 else 2
r = g.KerrRenderer(); r.init(); r.init_pipelines(mass=1.0, spin=0.9990000128746033, spec_w=64, spec_h=16, max_temp=1e7)
r.params = R.RenderParams(method=method, max_steps=steps, step_rule=1 if method else 0)
cam, _ = camera.default_camera(W, H)
phys = R.pack_physics(1.0, 0.9990000128746033, W, H)
spec = O.spectrum_lut(64, 16, 1e7); td = O.disk_lut(1.0, 0.9990000128746033)
opts = O.Options.default(method=method, step_rule=1 if method else 0, max_steps=steps)
rp, keep = O.make_render_params(W, H, 1.0, 0.9990000128746033, opts, spectrum=spec, spec_w=64, spec_h=16, tdisk=td)
ref = O.render(cam, rp)
got = r.trace_states(cam, phys)
peak = np.abs(ref["rgba"][..., :3]).max()
e = (np.abs(got["rgba"] - ref["rgba"]) / np.maximum(np.abs(ref["rgba"]), 1e-3 * peak)).max(-1)
ex = (np.abs(got["xp"] - ref["xp"]) / np.maximum(np.abs(ref["xp"]), 1.0)).max(-1)
print("rgba err pct", [f"{np.percentile(e, p):.2e}" for p in (50, 90, 99, 99.9, 100)])
print("state err pct", [f"{np.percentile(ex, p):.2e}" for p in (50, 90, 99, 99.9, 100)])
print("steps diff", (got["steps"] != ref["steps"]).sum(), "term diff", (got["term"] != ref["term"]).sum())
idx = np.argsort(e.ravel())[::-1][:25]
for k in idx:
    y, x = divmod(int(k), W)
    print(x, y, f"e={e[y, x]:.2e} ex={ex[y, x]:.2e} term={ref['term'][y, x]} steps={ref['steps'][y, x]}/{got['steps'][y, x]} "
          f"cross={ref['crossings'][y, x]} drift={ref['drift'][y, x]:.1e}/{got['drift'][y, x]:.1e}")
    print("    ref rgb", ref["rgba"][y, x, :3], "gpu", got["rgba"][y, x, :3])
    print("    ref xp ", ref["xp"][y, x], "\n    gpu xp ", got["xp"][y, x])
idx = np.argsort(ex.ravel())[::-1][:8]
print("worst states")
for k in idx:
    y, x = divmod(int(k), W)
    print(x, y, f"ex={ex[y, x]:.2e} e={e[y, x]:.2e} term={ref['term'][y, x]} steps={ref['steps'][y, x]} drift={ref['drift'][y, x]:.1e}")
