"""GPU-box check of the lean reciprocal / trig accuracy through the parity path: one RHS-heavy frame vs oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "blackhole-simulation_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import gravitas_b200 as g
from gravitas_b200 import camera, renderer as R
import oracle as O
spin = float(np.float32(0.999)); W, H, steps = 160, 90, 512
r = g.KerrRenderer(); r.init(); r.init_pipelines(mass=1.0, spin=spin, spec_w=64, spec_h=16, max_temp=1e7)
r.params = R.RenderParams(max_steps=steps)
cam, _ = camera.default_camera(W, H); phys = R.pack_physics(1.0, spin, W, H)
spec = O.spectrum_lut(64, 16, 1e7); td = O.disk_lut(1.0, spin)
opts = O.Options.default(method=2, step_rule=1, max_steps=steps)
rp, keep = O.make_render_params(W, H, 1.0, spin, opts, spectrum=spec, spec_w=64, spec_h=16, tdisk=td)
ref = O.render(cam, rp); got = r.trace_states(cam, phys)
peak = np.abs(ref["rgba"][..., :3]).max()
e = (np.abs(got["rgba"] - ref["rgba"]) / np.maximum(np.abs(ref["rgba"]), 1e-3 * peak)).max(-1)
ex = (np.abs(got["xp"] - ref["xp"]) / np.maximum(np.abs(ref["xp"]), 1.0)).max(-1)
print("rgba err pct", [f"{np.percentile(e, p):.2e}" for p in (50, 90, 99, 99.9, 100)])
print("state err pct", [f"{np.percentile(ex, p):.2e}" for p in (50, 90, 99, 99.9, 100)])
