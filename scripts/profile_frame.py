"""Render a few headline frames (config 3, budget accounting) — the command ncu wraps. Args: [frames] [precision] [method]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blackhole-simulation_b200"))
import gravitas_b200 as g
from gravitas_b200 import camera, renderer as R, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
method = int(sys.argv[3]) if len(sys.argv) > 3 else 2
W, H = 3840, 2160
spin = 0.9990000128746033
r = g.KerrRenderer(); r.init(); r.init_pipelines(mass=1.0, spin=spin, spec_w=256, spec_h=32, max_temp=1e7)
r.params = R.RenderParams(method=method, precision=prec, max_steps=512 if method else 1024, step_rule=1 if method else 0,
                          flags=_lib.FLAG_BUDGET if method else 0)
cam, _ = camera.default_camera(W, H)
phys = R.pack_physics(1.0, spin, W, H)
for i in range(n):
    r.render(cam, phys, readback=False)
    s = r.last_stats
    print(f"frame {i}: trace {s.trace_ms:.3f} ms, steps {s.steps_executed}, {s.steps_executed / s.trace_ms * 1e3:.4e} steps/s, rhs {s.rhs_evals}")
