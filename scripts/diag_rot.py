"""Which rays does the rotated-trig zone treat differently from the plain zones? Full headline frame, per-pixel steps / term."""
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
r.params = R.RenderParams(method=_lib.METHOD_SYMPLECTIC, max_steps=512, step_rule=1)
res = {}
for tag, env in (("rot", None), ("plain", "1")):
    if env: os.environ["GVT_NO_ROT"] = env
    else: os.environ.pop("GVT_NO_ROT", None)
    parts = []
    for y0 in range(0, H, 540):
        d = r.trace_states(cam, phys, y0=y0, y1=y0 + 540)
        parts.append((d["steps"].copy(), d["term"].copy(), d["xp"].copy(), d["rgba"].copy()))
    res[tag] = [np.concatenate([p[i] for p in parts]) for i in range(4)]
ds = res["rot"][0] != res["plain"][0]; dt = res["rot"][1] != res["plain"][1]
print("steps differ:", int(ds.sum()), "term differ:", int(dt.sum()))
for (yy, xx) in zip(*np.nonzero(ds | dt)):
    a, b = res["rot"], res["plain"]
    print(f"px ({xx},{yy}): steps {a[0][yy,xx]}/{b[0][yy,xx]} term {a[1][yy,xx]}/{b[1][yy,xx]}  final r {a[2][yy,xx,1]:.4g}/{b[2][yy,xx,1]:.4g} th {a[2][yy,xx,2]:.4g}/{b[2][yy,xx,2]:.4g} pth {a[2][yy,xx,6]:.4g}/{b[2][yy,xx,6]:.4g} pph {a[2][yy,xx,7]:.5g}  rgb {a[3][yy,xx,0]:.4g}/{b[3][yy,xx,0]:.4g}")
e = np.abs(res["rot"][3] - res["plain"][3]).max()
print("max abs rgba diff over the frame:", e, "of peak", res["plain"][3][..., :3].max())
