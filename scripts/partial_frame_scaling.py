"""How much of a rank's frame time at N GPUs is fixed cost? The N row blocks of the SAME 4K frame are traced one by one on ONE
GPU (gvt_render_rows: exactly the launches the N ranks make, minus the exchange); their times against the full-frame launch
separate per-launch cost (launch, LUT staging, tile-queue tail) and block imbalance from the collective. Args: [precision]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blackhole-simulation_b200"))
import numpy as np
import gravitas_b200 as g
from gravitas_b200 import camera, renderer as R, _lib
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 0
spin = float(np.float32(0.999))
W, H = 3840, 2160
r = g.KerrRenderer(); r.init(); r.init_pipelines(mass=1.0, spin=spin, spec_w=256, spec_h=32, max_temp=1e7)
cam, _ = camera.default_camera(W, H)
phys = R.pack_physics(1.0, spin, W, H)
for flags, name in ((_lib.FLAG_BUDGET, "budget"), (0, "natural")):
    r.params = R.RenderParams(method=_lib.METHOD_SYMPLECTIC, precision=prec, max_steps=512, step_rule=1, flags=flags)
    full = min(r.render_rows(cam, phys, 0, H).trace_ms for _ in range(4))
    for n in (2, 4, 8):
        rows = H // n
        ts = []
        for k in range(n):
            ts.append(min(r.render_rows(cam, phys, k * rows, (k + 1) * rows).trace_ms for _ in range(3)))
        print(f"precision {prec} {name}: full {full:.3f} ms; N={n}: blocks " + " ".join(f"{t:.3f}" for t in ts) +
              f" | sum {sum(ts):.3f} (x{sum(ts) / full:.4f}), slowest x N = {max(ts) * n:.3f} -> kernel-side efficiency {full / (max(ts) * n):.4f}, "
              f"mean-block efficiency {full / sum(ts):.4f}", flush=True)
