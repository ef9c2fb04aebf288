"""Per-warp timeline of one trace-kernel launch (GVT_TIMELINE_DUMP diagnostics): when does each warp start, when does it end, how
many tiles did it take -- for the full 4K frame and for the row block one rank of eight traces. Args: [rows] [precision]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blackhole-simulation_b200"))
import numpy as np
import gravitas_b200 as g
from gravitas_b200 import camera, renderer as R, _lib
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 270
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
spin = float(np.float32(0.999))
W, H = 3840, 2160
r = g.KerrRenderer(); r.init(); r.init_pipelines(mass=1.0, spin=spin, spec_w=256, spec_h=32, max_temp=1e7)
cam, _ = camera.default_camera(W, H)
phys = R.pack_physics(1.0, spin, W, H)
r.params = R.RenderParams(method=_lib.METHOD_SYMPLECTIC, precision=prec, max_steps=512, step_rule=1, flags=_lib.FLAG_BUDGET)
for _ in range(3):
    r.render_rows(cam, phys, 945, 945 + rows)
path = "/tmp/timeline.bin"
os.environ["GVT_TIMELINE_DUMP"] = path
st = r.render_rows(cam, phys, 945, 945 + rows)
os.environ["GVT_TIMELINE_DUMP"] = ""
t = np.fromfile(path, np.uint64).reshape(-1, 3)
t = t[t[:, 1] > 0]
t0 = t[:, 0].min()
start, end, n = (t[:, 0] - t0) * 1e-6, (t[:, 1] - t0) * 1e-6, t[:, 2]
print(f"rows {rows}: event-timed kernel {st.trace_ms:.3f} ms; {len(t)} warps; tiles/warp min {n.min()} mean {n.mean():.2f} max {n.max()}")
print(f"  warp start: min 0, p50 {np.percentile(start, 50):.4f}, max {start.max():.4f} ms")
print(f"  warp end:   min {end.min():.4f}, p10 {np.percentile(end, 10):.4f}, p50 {np.percentile(end, 50):.4f}, p90 {np.percentile(end, 90):.4f}, max {end.max():.4f} ms")
busy = (end - start).sum() / (len(t) * end.max())
print(f"  warp-busy fraction of the launch: {busy:.4f}; mean per-tile time {((end - start) / np.maximum(n, 1)).mean():.4f} ms")
sm = (np.arange(len(t)) // 16)
per_sm_end = np.array([end[sm == k].max() for k in range(sm.max() + 1)])
print(f"  per-SM last-warp end: min {per_sm_end.min():.4f} p50 {np.percentile(per_sm_end, 50):.4f} max {per_sm_end.max():.4f} ms")
for k in sorted(set(n)):
    sel = n == k
    print(f"  warps with {k} tiles: {sel.sum()}, end p50 {np.percentile(end[sel], 50):.4f}, per-tile {((end - start)[sel] / max(k, 1)).mean():.4f} ms")
