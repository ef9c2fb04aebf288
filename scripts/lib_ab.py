"""Frame times of the headline configuration for whichever build GRAVITAS_B200_LIB selects (A/B of two builds: run it once per
library, interleaved). Prints best-of-N kernel times: f64 budget / f64 natural / mixed budget / f32 budget."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blackhole-simulation_b200"))
import numpy as np
import gravitas_b200 as g
from gravitas_b200 import camera, renderer as R, _lib
spin = float(np.float32(0.999))
W, H = 3840, 2160
r = g.KerrRenderer(); r.init(); r.init_pipelines(mass=1.0, spin=spin, spec_w=256, spec_h=32, max_temp=1e7)
cam, _ = camera.default_camera(W, H)
phys = R.pack_physics(1.0, spin, W, H)
out = []
for prec, flags, name in ((0, _lib.FLAG_BUDGET, "f64 budget"), (0, 0, "f64 natural"), (3, _lib.FLAG_BUDGET, "mixed budget"), (1, _lib.FLAG_BUDGET, "f32 budget")):
    r.params = R.RenderParams(method=_lib.METHOD_SYMPLECTIC, precision=prec, max_steps=512, step_rule=1, flags=flags)
    ts = sorted(r.render_rows(cam, phys, 0, H).trace_ms for _ in range(6))
    out.append(f"{name} {ts[0]:.3f} (median {ts[3]:.3f})")
print(os.path.basename(_lib._SO), "|", " | ".join(out), flush=True)
