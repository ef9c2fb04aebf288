"""Both seams together, the way the reference app drives them: the worker ticks the PhysicsEngine and publishes the
SAB (workers/physics.worker.ts:111-176), the main thread reads the CAMERA / PHYSICS blocks through the seqlock
(engine/physics-bridge.ts:148-188), builds CameraUniforms (components/canvas/WebGPUCanvas.tsx:119-178) and calls
renderer.render(camera, physics) (rendering/webgpu/renderer.ts:280)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_engine_tick_to_rendered_frames(built, renderer, oracle):
    from gravitas_b200 import camera, renderer as R, _lib
    W, H, steps = 128, 72, 96
    spin = 0.9
    eng = built.PhysicsEngine(1.0, spin)
    sab = np.zeros(2 * 1024 * 1024 // 4, np.float32)           # SharedArrayBuffer(2 MiB) viewed as f32 (physics-bridge.ts:60)
    eng.attach_sab(sab)
    eng.set_camera_state(0.0, -3.7, 29.8)                      # position only (lib.rs:120-122)
    eng.set_auto_spin(True)
    renderer.init_pipelines(mass=1.0, spin=float(np.float32(spin)), spec_w=64, spec_h=16, max_temp=1e7)
    renderer.resize(W, H)
    renderer.reset_history()
    renderer.params = R.RenderParams(max_steps=steps, flags=_lib.FLAG_TAA | _lib.FLAG_JITTER)
    prev_vp, frames, seq_prev = None, [], float(sab[built.OFFSETS["TELEMETRY"]])
    for k in range(4):
        sab[built.OFFSETS["CONTROL"] + 1] = 0.3                # mouse_dx written by the main thread (physics-bridge.ts:153)
        sab[built.OFFSETS["CONTROL"] + 4] = 0.016
        eng.tick_sab(0.0)                                      # dt from CONTROL[4]
        cam_blk = sab[built.OFFSETS["CAMERA"]:built.OFFSETS["CAMERA"] + 16]
        phys_blk = sab[built.OFFSETS["PHYSICS"]:built.OFFSETS["PHYSICS"] + 16]
        assert np.all(np.isfinite(cam_blk)) and sab[built.OFFSETS["CONTROL"] + 1] == 0.0    # input consumed
        assert phys_blk[0] == np.float32(eng.compute_horizon()) and phys_blk[3] == np.float32(spin)
        eye = tuple(float(v) for v in cam_blk[0:3])
        cam, vp = camera.camera_uniforms(eye, W, H, prev_view_proj=prev_vp)
        phys = R.pack_physics(float(phys_blk[2]), float(phys_blk[3]), W, H, frame_index=k)
        frame = np.array(renderer.render(cam, phys))
        assert frame.shape == (H, W, 4) and np.isfinite(frame).all() and (frame[..., :3].sum() > 0)
        frames.append((eye, frame))
        prev_vp = vp
    # the camera orbits (yaw from mouse_dx + auto-spin) at constant radius, and TAA accumulates: frames differ and brighten
    r = [math.sqrt(sum(c * c for c in e)) for e, _ in frames]
    assert max(r) - min(r) < 1e-3 and frames[0][0] != frames[-1][0]
    assert frames[1][1][..., :3].sum() > frames[0][1][..., :3].sum()      # history was zero on the first frame (0.92 blend)
    # the un-resolved render of the last camera matches the oracle pixel for pixel
    renderer.params = R.RenderParams(max_steps=steps)
    cam, _ = camera.camera_uniforms(frames[-1][0], W, H)
    phys = R.pack_physics(1.0, spin, W, H)
    got = np.array(renderer.render(cam, phys))
    s32 = float(np.float32(spin))
    spec, td = oracle.spectrum_lut(64, 16, 1e7), oracle.disk_lut(1.0, s32)
    opts = oracle.Options.default(method=2, step_rule=1, max_steps=steps)
    rp, keep = oracle.make_render_params(W, H, 1.0, s32, opts, spectrum=spec, spec_w=64, spec_h=16, tdisk=td)
    ref = oracle.render(cam, rp, want=("rgba",))
    peak = float(ref["rgba"][..., :3].max())
    assert (np.abs(got - ref["rgba"]) <= 1e-6 * np.maximum(np.abs(ref["rgba"]), 1e-3 * peak) + 1e-30).all()


def test_bench_own_arm_json_contract():
    """`python bench.py` (product arm, one GPU): one JSON line with the contract's keys; e2e measured through host
    buffers; roofline / clocks / gpu_launches present; side kernels reported without error."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "3", "--warmup", "3", "--no-cpu"],
                       capture_output=True, text=True, timeout=900, cwd=root)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [l for l in p.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["value"] > 5e10 and abs(d["value"] * d["ms_per_step"] * 1e-3 - 3840 * 2160 * 512) < 1e-3 * 3840 * 2160 * 512
    assert d["gpu_launches"] == 3 and "workload" in d["config"] and "model" not in d["config"]
    e = d["e2e"]
    assert e["value"] > 5e10 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] >= 3840 * 2160 * 16
    r = d["roofline"]
    assert 0.5 < r["frac"] < 1.0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # traffic is read from the committed ncu export of THIS build (keyed by the kernel-source hash); a build newer than the
    # last capture must say so instead of reporting a stale number
    assert (r["traffic"] is not None and r["traffic"] > 0) or (r["traffic"] is None and r["traffic_note"])
    assert d["clocks"]["sm_mhz"] > 0 and isinstance(d["clocks"]["reasons"], list)
    x = d["extra"]
    assert "side_kernels_error" not in x and 0.2 < x["taa_resolve"]["frac"] < 1.0 and x["webgl_fragment_shader"]["ms"] > 0
