"""A few 4K frames with jitter + TAA (config-5 style) — the command ncu wraps to capture k_taa_resolve."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blackhole-simulation_b200"))
import gravitas_b200 as g
from gravitas_b200 import camera, renderer as R, _lib
W, H = 3840, 2160
spin = 0.9990000128746033
r = g.KerrRenderer(); r.init(); r.init_pipelines(mass=1.0, spin=spin, spec_w=256, spec_h=32, max_temp=1e7)
r.params = R.RenderParams(max_steps=64, flags=_lib.FLAG_TAA | _lib.FLAG_JITTER)
prev = None
for k in range(4):
    cam, vp = camera.default_camera(W, H, azimuth=math.pi + 0.005 * k, prev_view_proj=prev)
    r.render(cam, R.pack_physics(1.0, spin, W, H, frame_index=k), readback=False)
    prev = vp
    print(f"frame {k}: trace {r.last_stats.trace_ms:.3f} ms taa {r.last_stats.taa_ms:.3f} ms")
