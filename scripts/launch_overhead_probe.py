"""What does a trace launch cost beyond its warps' own busy span? Event-timed trace_ms of (a) a near-empty launch (8 rows,
16 steps), (b) the 270-row block of one rank of eight, next to the in-kernel first-start -> last-end span of the same launch
(GVT_TIMELINE_DUMP)."""
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
def span(path):
    t = np.fromfile(path, np.uint64).reshape(-1, 3); t = t[t[:, 1] > 0]
    return (t[:, 1].max() - t[:, 0].min()) * 1e-6
for steps, rows in ((16, 8), (16, 270), (512, 8), (512, 270)):
    r.params = R.RenderParams(method=_lib.METHOD_SYMPLECTIC, precision=0, max_steps=steps, step_rule=1, flags=_lib.FLAG_BUDGET)
    ev = sorted(r.render_rows(cam, phys, 945, 945 + rows).trace_ms for _ in range(9))
    os.environ["GVT_TIMELINE_DUMP"] = "/tmp/tl.bin"
    st = r.render_rows(cam, phys, 945, 945 + rows)
    os.environ["GVT_TIMELINE_DUMP"] = ""
    print(f"steps {steps} rows {rows}: event-timed min {ev[0]:.4f} median {ev[4]:.4f} ms; with timeline {st.trace_ms:.4f} ms, in-kernel span {span('/tmp/tl.bin'):.4f} ms", flush=True)
