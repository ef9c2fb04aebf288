"""A few 4K frames with jitter + TAA (config-5 style) — the command ncu wraps to capture k_taa_resolve.
Args: [frames] [webgl]; prints per-frame trace / TAA times and the median TAA time."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blackhole-simulation_b200"))
import gravitas_b200 as g
from gravitas_b200 import camera, renderer as R, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
webgl = len(sys.argv) > 2 and sys.argv[2] == "webgl"
W, H = 3840, 2160
spin = 0.9990000128746033
r = g.KerrRenderer(); r.init(); r.init_pipelines(mass=1.0, spin=spin, spec_w=256, spec_h=32, max_temp=1e7)
r.params = R.RenderParams(max_steps=64, flags=_lib.FLAG_TAA | _lib.FLAG_JITTER | (_lib.FLAG_TAA_WEBGL if webgl else 0))
prev = None
ts = []
for k in range(n):
    cam, vp = camera.default_camera(W, H, azimuth=math.pi + 0.005 * k, prev_view_proj=prev)
    r.render(cam, R.pack_physics(1.0, spin, W, H, frame_index=k), readback=False)
    prev = vp
    ts.append(r.last_stats.taa_ms)
    if n <= 8:
        print(f"frame {k}: trace {r.last_stats.trace_ms:.3f} ms taa {r.last_stats.taa_ms:.3f} ms")
ts = sorted(ts[1:])
print(f"taa median {ts[len(ts) // 2] * 1e3:.1f} us  min {ts[0] * 1e3:.1f} us  ({os.environ.get('GRAVITAS_B200_LIB', 'default lib')})")
