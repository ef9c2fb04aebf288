"""GPU-box diagnostic: single-step GPU vs oracle along the oracle trajectory of chosen pixels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "blackhole-simulation_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import gravitas_b200 as g
from gravitas_b200 import camera, renderer as R, _lib
import oracle as O

W, H = 160, 90
cam, _ = camera.default_camera(W, H)
rp, keep = O.make_render_params(W, H, 1.0, 0.999, O.Options.default())
e = g.PhysicsEngine(1.0, 0.999)
rh = e.compute_horizon()
for (px, py) in [(75, 45), (79, 24), (20, 20)]:
    s = O.camera_ray(cam, rp, px, py)
    s = O.renormalize(1.0, 0.999, 1, s)
    states, hs = [], []
    for k in range(100):
        if s[1] < rh * 1.001: break
        h = min(max((s[1] - rh) * 0.15, 0.05), 1.0)
        states.append(s.copy()); hs.append(h)
        s = O.step_symplectic(1.0, 0.999, 1, s, h)
        if k % 10 == 0: s = O.renormalize(1.0, 0.999, 1, s)
    states = np.array(states)
    print("pixel", px, py, "n", len(states))
    worst = 0
    for k, (st, h) in enumerate(zip(states, hs)):
        # one GPU step with constant step h, renorm interval huge -> but integrate() renormalises at start and at steps%interval==0 (step 0)
        p = R.RenderParams(method=2, coords=1, step_rule=0, max_steps=1, initial_step=h, renormalize_interval=10)
        got = e.integrate_rays(st[None, :], p)["xp"][0]
        opts = O.Options.default(method=2, step_rule=0, max_steps=1, initial_step=h)
        ref = O.integrate(1.0, 0.999, 1, opts, st)["xp"][0]
        d = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3)
        worst = max(worst, d.max())
        if d.max() > 1e-13:
            print(f"  step {k} r={st[1]:.5f} th={st[2]:.5f} h={h:.3f} maxrel={d.max():.2e} comp={d.argmax()} ref={ref[d.argmax()]:.15e} gpu={got[d.argmax()]:.15e}")
    print("  worst single-step rel diff", worst)
