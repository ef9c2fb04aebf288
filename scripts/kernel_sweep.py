"""Kernel tuning sweep on one B200 (run through gpurun): for each precision mode of k_trace_tile time the headline frame
(config 3: 3840x2160x512, a*=0.999, budget accounting, bench LUT 256x32) and compare a strided lattice of it with the
CPU oracle (per-pixel RGBA, termination, step count). Args: [lattice stride] [modes, e.g. 0,3,1] [frames]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blackhole-simulation_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import gravitas_b200 as g
from gravitas_b200 import camera, renderer as R, _lib
import oracle as O

stride = int(sys.argv[1]) if len(sys.argv) > 1 else 16
modes = [int(m) for m in (sys.argv[2] if len(sys.argv) > 2 else "0,3,1").split(",")]
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 4
W, H, STEPS = 3840, 2160, 512
spin = float(np.float32(0.999))
r = g.KerrRenderer(); r.init(); r.init_pipelines(mass=1.0, spin=spin, spec_w=256, spec_h=32, max_temp=1e7)
cam, _ = camera.default_camera(W, H)
phys = R.pack_physics(1.0, spin, W, H)
lat = dict(x0=stride // 2, xs=stride, y0=stride // 2, ys=stride)
spec, td = O.spectrum_lut(256, 32, 1e7), O.disk_lut(1.0, spin)
opts = O.Options.default(method=O.METHOD_SYMPLECTIC, step_rule=1, max_steps=STEPS)
rp, keep = O.make_render_params(W, H, 1.0, spin, opts, spectrum=spec, spec_w=256, spec_h=32, tdisk=td)
t0 = time.time()
ref = O.render(cam, rp, **lat)
peak = float(np.abs(ref["rgba"][..., :3]).max())
print(f"oracle lattice {ref['rgba'].shape[:2]} in {time.time() - t0:.1f} s; lib = {_lib._SO}", flush=True)
names = {0: "f64", 1: "f32", 3: "mixed"}
out = {}
for m in modes:
    r.params = R.RenderParams(method=_lib.METHOD_SYMPLECTIC, precision=m, max_steps=STEPS, step_rule=1, flags=_lib.FLAG_BUDGET)
    ms = []
    for i in range(frames):
        r.render(cam, phys, readback=False)
        ms.append(r.last_stats.trace_ms)
    st = r.last_stats
    r.params = R.RenderParams(method=_lib.METHOD_SYMPLECTIC, precision=m, max_steps=STEPS, step_rule=1, flags=0)
    r.render(cam, phys, readback=False)
    nat = r.last_stats
    frame = np.array(r.render(cam, phys))[lat["y0"]::lat["ys"], lat["x0"]::lat["xs"]].astype(np.float64)
    ef = (np.abs(frame - ref["rgba"]) / np.maximum(np.abs(ref["rgba"]), 1e-3 * peak)).max(-1)
    got = r.trace_states(cam, phys, **lat)
    e = (np.abs(got["rgba"] - ref["rgba"]) / np.maximum(np.abs(ref["rgba"]), 1e-3 * peak)).max(-1)
    lit = ref["rgba"][..., :3].sum(-1) > 0
    res = {"trace_ms_frames": [round(x, 3) for x in ms], "trace_ms_best": min(ms[1:] or ms), "steps_executed": int(st.steps_executed),
           "steps_per_s": st.steps_executed / (min(ms[1:] or ms) * 1e-3), "natural_ms": nat.trace_ms,
           "natural_steps": int(nat.steps_committed),
           "census": [int(nat.n_horizon), int(nat.n_escape), int(nat.n_maxsteps), int(nat.n_disk)],
           "frame_gt_1e-6": int((ef > 1e-6).sum()), "frame_max_rel": float(ef.max()),
           "lattice_px": int(e.size), "gt_1e-6": int((e > 1e-6).sum()), "gt_1e-5": int((e > 1e-5).sum()), "max_rel": float(e.max()),
           "median_lit": float(np.median(e[lit])), "p999": float(np.percentile(e, 99.9)),
           "term_diff": int((got["term"] != ref["term"]).sum()), "steps_diff": int((got["steps"] != ref["steps"]).sum())}
    out[names.get(m, str(m))] = res
    print(names.get(m, m), json.dumps(res), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
tag = os.environ.get("SWEEP_TAG", "sweep")
with open(os.path.join(ROOT, "gpurun_out", f"{tag}.json"), "w") as f:
    json.dump(out, f, indent=1)
