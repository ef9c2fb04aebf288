#!/usr/bin/env python
"""bench.py — geodesic steps/s on BASELINE.json's headline workload (config 3: Kerr a*=0.999, 3840x2160, 512
fixed steps, f64 + Planckian redshift LUT), one process per GPU.

  python bench.py [--gpus N --steps K --warmup W]            own arm (sm_100a kernels through the C ABI)
  python bench.py --impl reference [...]                      the reference's CPU path on the host cores

A "step" of the bench = one 4K frame traced (the hot path over one batch of synthetic input: the SURVEY §8(d)
camera). `value` = geodesic steps/s with the frame resident in HBM (no read-back in the timed region), budget
accounting (every pixel executes exactly 512 step computations => 3840*2160*512 steps per frame); `e2e` = the
same metric through the public render(camera, physics) call with HOST buffers: camera/physics uniforms go H2D and
the finished RGBA32F frame comes back D2H into pinned memory inside the timed region, every frame.
Timing: CUDA events on the library's launch stream (first op -> last op of each frame), summed over the K timed
frames, max over ranks; wall-clock is reported beside it as a cross-check.

The reference is Rust (gravitas-core) and cannot be built in this image (no cargo/rustc/wasm-pack/node), so the
reference arm and `cpu_baseline` time oracle/ — the C++ restatement of gravitas-core's integrate() — kind "port".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "blackhole-simulation_b200"))

W, H, STEPS, MASS = 3840, 2160, 512, 1.0
SPIN = 0.9990000128746033   # 0.999 as the f32 the PhysicsParams uniform carries (types/webgpu.ts:42-64)
SPEC_W, SPEC_H, TMAX = 256, 32, 1e7          # 128 KB RGBA32F spectral LUT: shared-memory resident
# Algorithmic flop per geodesic step, SURVEY.md §8(d) (CSE'd count; add=mul=div=sqrt=1, FMA=2, sincos/pow = 0):
FLOP_PER_STEP = {"symplectic": 330.0, "rk4": 440.0, "rkf45": 900.0}
# Profile-derived figures (DRAM traffic per launch, executed instruction mix, pipe activity) are READ from the committed,
# machine-readable ncu export of the dominant kernel -- never pasted -- and only reported when the export was captured on
# the build that is running (gvt_build_info() source hash); otherwise they are null and marked stale.
PROFILE_EXPORT = os.path.join(ROOT, "profiles", "r02_k_trace_tile_f64_budget_4k512.json")
PROFILE_EXPORT_MIXED = os.path.join(ROOT, "profiles", "r02_k_trace_tile_mixed_budget_4k512.json")
METRIC = "geodesic steps/s at 3840x2160x512, a=0.999; % of FP32 roofline"


def workload_config(n_gpus, peer_store=False, interleave=False):
    return {
        "workload": "config 3: Kerr a*=0.999 (Kerr-Schild), 3840x2160, 512 fixed implicit-midpoint steps/pixel "
                    "(step rule compute.wgsl.ts:213), f64, thin-disk g-factor + Planckian redshift LUT 256x32, "
                    "camera r0=30 polar 97deg azimuth pi fov 60deg; budget accounting (W*H*512 steps/frame)",
        "width": W, "height": H, "steps_per_pixel": STEPS, "spin": SPIN, "integrator": "implicit-midpoint",
        "precision": "f64", "shard": ((f"rows dealt round-robin to {n_gpus} ranks + " if interleave else f"row-block x{n_gpus} + ") +
                                      ("NVLink peer stores fused into the trace kernel" if peer_store else "1 ncclAllGather"))
        if n_gpus > 1 else "single GPU",
        "l2": "inputs are ~131 KB (LUT + camera block), compute-bound and L2-insensitive; each frame writes a "
              "132.7 MB RGBA32F frame (> 126 MB L2), so no explicit L2 flush between iterations",
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist_  # plumbing only: rendezvous, id broadcast, barrier, max-over-ranks
        dist_.init_process_group("gloo")
        dist = dist_
    return rank, world, local, dist


def barrier(dist):
    if dist is not None:
        dist.barrier()


def allreduce_max(dist, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def allreduce_sum(dist, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0])


def load_profile_export(path, build_info):
    """The committed ncu export of a kernel, or None if absent / captured on another build of the library."""
    try:
        with open(path) as f:
            ex = json.load(f)
    except (OSError, ValueError):
        return None, "no committed ncu export"
    if ex.get("build_info") != build_info:
        return None, f"committed ncu export is stale (captured on '{ex.get('build_info')}', running '{build_info}')"
    return ex, None


def cpu_sample(target_seconds=15.0, threads=None):
    """Time the oracle (C++ restatement of gravitas-core integrate(), built -O3 -march=native on this host, contraction
    off) on a strided lattice of the SAME 4K frame, sized to ~target_seconds of CPU work. threads=None: all host threads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from gravitas_b200 import camera
    spec = O.spectrum_lut(SPEC_W, SPEC_H, TMAX)
    td = O.disk_lut(MASS, SPIN)
    opts = O.Options.default(method=O.METHOD_SYMPLECTIC, step_rule=1, max_steps=STEPS)
    rp, keep = O.make_render_params(W, H, MASS, SPIN, opts, spectrum=spec, spec_w=SPEC_W, spec_h=SPEC_H, tdisk=td)
    cam, _ = camera.default_camera(W, H)
    probe = O.render(cam, rp, x0=5, xs=32, y0=3, ys=32, want=(), native=True, threads=threads)   # 1/1024 of the frame
    cores = threads or O.lib(True).orc_num_threads()
    rate = probe["total_steps"] / max(probe["seconds"], 1e-9)
    frac = min(1.0, target_seconds * rate / (W * H * STEPS * 0.96))
    stride = max(1, int(round((1.0 / frac) ** 0.5)))
    res = O.render(cam, rp, x0=stride // 2, xs=stride, y0=stride // 2, ys=stride, want=(), native=True, threads=threads)
    frame_factor = W * H / res["n"]          # sampled rays -> whole frame
    return {"value": res["total_steps"] / res["seconds"], "unit": "steps/s", "cores": cores, "kind": "port",
            "sample": f"every {stride}th pixel in x and y of the same 3840x2160x512 frame ({res['n']} rays, "
                      f"{res['total_steps']} accepted steps, {res['seconds']:.2f} s on {cores} thread(s)); natural termination; "
                      f"g++ -O3 -march=native -ffp-contract=off",
            "seconds": res["seconds"], "steps": res["total_steps"], "frame_factor": frame_factor,
            "frame_ms_extrapolated": 1e3 * res["seconds"] * frame_factor}


def cpu_config1():
    """BASELINE configs[0], the reference's own CPU-runnable case, in full: Schwarzschild a = 0, 256 x 256 camera rays, 128
    adaptive RKF45 steps, Boyer-Lindquist (gravitas-wasm/src/lib.rs:444-452 options), all host threads."""
    import numpy as np
    import oracle as O
    from gravitas_b200 import camera
    Wx = Hx = 256
    cam, _ = camera.default_camera(Wx, Hx)
    opts = O.Options.default(max_steps=128)
    rp_o, keep = O.make_render_params(Wx, Hx, 1.0, 0.0, opts, coords=0)
    rays = np.array([O.camera_ray(cam, rp_o, x, y) for y in range(Hx) for x in range(Wx)])
    O.integrate(1.0, 0.0, 0, opts, rays[:4096], native=True)           # warm the thread pool / caches
    t0 = time.perf_counter()
    ref = O.integrate(1.0, 0.0, 0, opts, rays, native=True)
    dt = time.perf_counter() - t0
    steps = float(ref["steps"].sum())
    return {"workload": "config 1: Schwarzschild a=0, 256x256 rays, 128 adaptive RKF45 steps, Boyer-Lindquist", "ms": 1e3 * dt,
            "steps_per_s": steps / dt, "rhs_evals_per_s": float(ref["rhs"].sum()) / dt, "cores": O.lib(True).orc_num_threads()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # K "steps", each a bounded sample of the frame; W warm-up samples
    # ~90 s of CPU work in total (the probe-based sizing runs ~1.7x over its target on many-core boxes)
    per = max(1.0, min(10.0, 55.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_sample(per)
    tot_steps, tot_s, last = 0, 0.0, None
    for _ in range(args.steps):
        last = cpu_sample(per)
        tot_steps += last["steps"]; tot_s += last["seconds"]
    v = tot_steps / tot_s
    cb = {"value": v, "unit": "steps/s", "cores": last["cores"], "kind": "port", "sample": last["sample"]}
    # a bench "step" is one 4K frame; each timed sample covered 1/frame_factor of it, so the per-frame time is the
    # sample time x frame_factor (rays are sampled on a uniform lattice of the same frame)
    frame_ms = 1e3 * (tot_s / args.steps) * last["frame_factor"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": frame_ms,
        "ms_per_step_note": f"whole-frame time extrapolated from the timed sample: sample seconds x {last['frame_factor']:.1f} "
                            f"(= 3840*2160 / sampled rays); the sample itself took {1e3 * tot_s / args.steps:.0f} ms per bench step",
        "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": cb, "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = Rust gravitas-core; not buildable here (no cargo/rustc), so this arm times oracle/ "
                "(C++ restatement of geodesic::integrate + the composite shading), all host threads",
    }))


def timed_frames(r, cam, phys, n, dist, readback, out=None):
    """n frames; returns (sum of event-timed ms [max over ranks], wall ms, per-frame stats list)."""
    barrier(dist)
    t0 = time.perf_counter()
    ev_ms, stats = 0.0, []
    for _ in range(n):
        r.render(cam, phys, out=out, readback=readback)
        ev_ms += r.last_stats.total_ms
        stats.append(r.last_stats)
    wall = (time.perf_counter() - t0) * 1e3
    barrier(dist)
    return allreduce_max(dist, ev_ms), allreduce_max(dist, wall), stats


def side_kernels(r, local, camera, R, _lib):
    """Beside the headline (one GPU): the two other kernels of the path, a few 4K frames each."""
    side = {}
    import math
    hbm_gbs, hbm_src = 6650.0, "fallback of /opt/skills/guides/B200_PROFILING.md"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_gbs, hbm_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    # TAA resolve (config-5 style frames: jitter + TAA, short march so the resolve is not lost in event noise)
    r.params = R.RenderParams(method=_lib.METHOD_SYMPLECTIC, precision=_lib.PRECISION_F32, max_steps=32,
                              step_rule=_lib.STEP_WGSL, flags=_lib.FLAG_TAA | _lib.FLAG_JITTER)
    prev, taa_ms = None, []
    for k in range(6):
        c5, vp = camera.default_camera(W, H, azimuth=math.pi + 0.005 * k, prev_view_proj=prev)
        r.render(c5, R.pack_physics(MASS, SPIN, W, H, frame_index=k), readback=False)
        prev = vp
        if k >= 2:
            taa_ms.append(r.last_stats.taa_ms)
    t = sorted(taa_ms)[len(taa_ms) // 2]
    taa_bytes = 48 * W * H                      # 16 B current + 16 B history + 16 B store per pixel
    side["taa_resolve"] = {"kernel": "k_taa_resolve<0>", "ms": t, "bound": "hbm", "algorithmic_bytes": taa_bytes,
                           "achieved_gbs": taa_bytes / (t * 1e-3) * 1e-9, "peak_gbs": hbm_gbs, "peak_source": hbm_src,
                           "frac": taa_bytes / (t * 1e-3) * 1e-9 / hbm_gbs}
    # the same resolve on the RGBA16F-native frame chain (the reference's texture format): 8 B current + 8 B history + 8 B store
    r.set_frame_format(_lib.FORMAT_RGBA16F)
    try:
        prev, taa16 = None, []
        for k in range(6):
            c5, vp = camera.default_camera(W, H, azimuth=math.pi + 0.005 * k, prev_view_proj=prev)
            r.render(c5, R.pack_physics(MASS, SPIN, W, H, frame_index=k), readback=False)
            prev = vp
            if k >= 2:
                taa16.append(r.last_stats.taa_ms)
        t16 = sorted(taa16)[len(taa16) // 2]
        b16 = 24 * W * H
        side["taa_resolve_rgba16f"] = {"kernel": "k_taa_resolve<0, F16>", "ms": t16, "bound": "hbm", "algorithmic_bytes": b16,
                                       "achieved_gbs": b16 / (t16 * 1e-3) * 1e-9, "peak_gbs": hbm_gbs, "frac": b16 / (t16 * 1e-3) * 1e-9 / hbm_gbs,
                                       "pixels_per_s": W * H / (t16 * 1e-3)}
    finally:
        r.set_frame_format(_lib.FORMAT_RGBA32F)
    # the whole WebGL2 fragment shader (k_fragment_glsl, MUFU build), ultra-quality preset, + bloom / final pass
    from gravitas_b200 import webgl
    wr = webgl.WebGLRenderer(device=local, noise_seed=11)
    if wr.init():
        wr.precision = _lib.PRECISION_F32_FAST
        wr.resize(W, H)
        sp = dict(mass=1.0, spin=0.9, zoom=30.0, lensing=1.0, features=dict(webgl.PRESETS["ultra-quality"], bloom=False))
        ms = []
        for k in range(5):
            wr.render(sp, {"x": 0.5, "y": 0.5 + 7.0 / 180.0}, readback=False)
            if k >= 2:
                ms.append(wr.last_stats.trace_ms)
        st = wr.last_stats
        m = sorted(ms)[len(ms) // 2]
        wr._k.bloom(enabled=True, readback=False); wr._k.bloom(enabled=True, readback=False)
        side["webgl_fragment_shader"] = {"kernel": "k_fragment_glsl<float> (MUFU build)", "preset": "ultra-quality, 256-step budget",
                                         "ms": m, "fps": 1e3 / m, "march_steps_per_s": st.steps_committed / (m * 1e-3),
                                         "mean_steps_per_pixel": st.steps_committed / (W * H),
                                         "bloom_final_pass_ms": wr._k.last_bloom_ms}
        wr.cleanup()
    return side


def run_own(args):
    import gravitas_b200 as g
    from gravitas_b200 import _lib, camera, renderer as R

    rank, world, local, dist = dist_setup(args.gpus)
    nccl_id = None
    if world > 1:
        objs = [g.KerrRenderer.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(objs, src=0)
        nccl_id = objs[0]
    r = g.KerrRenderer(device=local, rank=rank, world_size=world, nccl_id=nccl_id)
    r.init()   # raises GravitasError if no sm_100 device / library missing: there is no fallback
    r.init_pipelines(max_steps=STEPS, mass=MASS, spin=SPIN, spec_w=SPEC_W, spec_h=SPEC_H, max_temp=TMAX)
    cam, _ = camera.default_camera(W, H)
    phys = R.pack_physics(MASS, SPIN, W, H)
    if world > 1:
        # one host frame for the box: every rank page-locks the same shared-memory segment and copies its own row
        # block into it (GVT_FLAG_D2H_OWN_ROWS); rank 0 is the consumer
        names = [None]
        if rank == 0:
            pinned = R.SharedFrame(W, H)
            names = [pinned.name]
        dist.broadcast_object_list(names, src=0)
        if rank != 0:
            pinned = R.SharedFrame(W, H, name=names[0], create=False)
    else:
        pinned = r.pinned_frame(W, H)
    info = r.device_info()

    peer = 0
    if world > 1 and not args.no_peer_store:
        # both gathers are GPU paths of the library; if CUDA IPC is not permitted on this box every rank falls back to
        # the ncclAllGather together (the choice is all-reduced so that no rank is left in the other protocol)
        r.resize(W, H)
        ok = 1.0
        try:
            r.connect_peers(dist)
        except g.GravitasError as ex:
            ok = 0.0
            print(f"[bench] rank {rank}: peer-store unavailable ({ex}); using ncclAllGather", file=sys.stderr)
        if allreduce_sum(dist, ok) == world:
            peer = _lib.FLAG_PEER_STORE
            if args.shard == "interleaved":
                # rows dealt round-robin: the zones of the march (near the hole / saturated step rule / polar tiles) cost
                # differently per step, and a contiguous row block holds an uneven share of them
                peer |= _lib.FLAG_ROW_INTERLEAVE

    def params(flags=0, **kw):
        r.params = R.RenderParams(method=_lib.METHOD_SYMPLECTIC, precision=_lib.PRECISION_F64, max_steps=STEPS,
                                  step_rule=_lib.STEP_WGSL, flags=flags | peer, **kw)

    # roofline denominators (MEASURED_PEAKS.json carries HBM and bf16 only): in-run DFMA / FFMA micro-benchmarks
    peak64, _ = r.measure_fma_peak(_lib.PRECISION_F64)
    peak32, _ = r.measure_fma_peak(_lib.PRECISION_F32)

    sampler = ClockSampler(local)
    params(flags=_lib.FLAG_BUDGET)
    for _ in range(args.warmup):
        r.render(cam, phys, readback=False)
    if rank == 0:
        sampler.start()
    # ---- value: frame resident in HBM, budget accounting ----
    ev_ms, wall_ms, stats = timed_frames(r, cam, phys, args.steps, dist, readback=False)
    total_steps = allreduce_sum(dist, float(sum(s.steps_executed for s in stats)))
    launches = sum(s.kernel_launches for s in stats)
    trace_ms = sum(s.trace_ms for s in stats) / len(stats)
    gather_ms = sum(s.gather_ms for s in stats) / len(stats)
    value = total_steps / (ev_ms * 1e-3)
    # ---- e2e: host buffers every frame ----
    if world > 1:
        r.params.c.flags |= _lib.FLAG_D2H_OWN_ROWS
    e_ms, e_wall, e_stats = timed_frames(r, cam, phys, args.steps, dist, readback=True, out=pinned)
    e_steps = allreduce_sum(dist, float(sum(s.steps_executed for s in e_stats)))
    e2e = {"value": e_steps / (e_ms * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": int(e_stats[0].h2d_bytes),
           "d2h_bytes_per_step": int(e_stats[0].d2h_bytes), "ms_per_step": e_ms / args.steps,
           "wall_ms_per_step": e_wall / args.steps,
           "host_frame": "pinned buffer" if world == 1 else
                         f"one shared-memory host frame page-locked by all {world} ranks; each rank copies its own row block "
                         "(bytes above are per rank; the box moves one 132.7 MB frame per step in total)"}
    clocks = sampler.stop() if rank == 0 else None
    # ---- beside the headline: natural termination, and the f32 instantiation (config-2 arithmetic) ----
    params(flags=0)
    n_ms, _, n_stats = timed_frames(r, cam, phys, max(2, args.steps // 3), dist, readback=False)
    n_steps = allreduce_sum(dist, float(sum(s.steps_committed for s in n_stats)))
    s0 = n_stats[0]
    r.params = R.RenderParams(method=_lib.METHOD_SYMPLECTIC, precision=_lib.PRECISION_F32, max_steps=STEPS,
                              step_rule=_lib.STEP_WGSL, flags=_lib.FLAG_BUDGET | peer)
    r.render(cam, phys, readback=False)
    f_ms, _, f_stats = timed_frames(r, cam, phys, max(2, args.steps // 3), dist, readback=False)
    f_steps = allreduce_sum(dist, float(sum(s.steps_executed for s in f_stats)))
    # GVT_PRECISION_MIXED: f64 state and corrector, f32 predictors beyond 35 M (every pixel of this frame within 1e-6 of the
    # all-f64 oracle: tests/test_gpu_parity.py::test_headline_frame_every_pixel). Same frame, same accounting.
    r.params = R.RenderParams(method=_lib.METHOD_SYMPLECTIC, precision=_lib.PRECISION_MIXED, max_steps=STEPS,
                              step_rule=_lib.STEP_WGSL, flags=_lib.FLAG_BUDGET | peer)
    r.render(cam, phys, readback=False)
    m_ms, _, m_stats = timed_frames(r, cam, phys, max(2, args.steps // 3), dist, readback=False)
    m_steps = allreduce_sum(dist, float(sum(s.steps_executed for s in m_stats)))
    m_trace_ms = sum(s.trace_ms for s in m_stats) / len(m_stats)
    # N > 1: the exchange north_star names -- one ncclAllGather after the trace kernel -- beside the default fused gather
    ag = None
    if world > 1 and peer:
        r.params = R.RenderParams(method=_lib.METHOD_SYMPLECTIC, precision=_lib.PRECISION_F64, max_steps=STEPS,
                                  step_rule=_lib.STEP_WGSL, flags=_lib.FLAG_BUDGET)
        r.render(cam, phys, readback=False)
        a_ms, _, a_stats = timed_frames(r, cam, phys, max(2, args.steps // 2), dist, readback=False)
        a_steps = allreduce_sum(dist, float(sum(s.steps_executed for s in a_stats)))
        ag = {"steps_per_s": a_steps / (a_ms * 1e-3), "ms_per_frame": a_ms / len(a_stats),
              "all_gather_ms": sum(s.gather_ms for s in a_stats) / len(a_stats),
              "trace_kernel_ms_rank0": sum(s.trace_ms for s in a_stats) / len(a_stats)}

    # ---- beside the headline (one GPU only): the two other kernels of the path, a few frames each ----
    side = {}
    if world == 1:
        try:
            side = side_kernels(r, local, camera, R, _lib)
        except Exception as ex:   # the side measurements must never cost the headline line
            side = {"side_kernels_error": repr(ex)}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        full = cpu_sample(15.0)
        cpu = {k: full[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cpu["frame_ms_extrapolated"] = full["frame_ms_extrapolated"]
        one = cpu_sample(5.0, threads=1)                      # the single-thread figure beside the all-core one
        cpu["single_thread"] = {"value": one["value"], "unit": "steps/s", "sample": one["sample"]}
        cpu["config1"] = cpu_config1()                        # the reference's own CPU-runnable case, in full
    build_info = _lib.lib().gvt_build_info().decode()
    prof, prof_note = load_profile_export(PROFILE_EXPORT, build_info)
    prof_m, _ = load_profile_export(PROFILE_EXPORT_MIXED, build_info)

    if rank == 0:
        fl = FLOP_PER_STEP["symplectic"]
        # roofline of the dominant kernel (k_trace_tile): per-launch algorithmic flops / per-launch event time.
        # At N>1 each rank's launch covers its own row block.
        steps_per_launch = stats[0].steps_executed
        ach = steps_per_launch * fl / (trace_ms * 1e-3) * 1e-12
        out = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ev_ms / args.steps, "wall_ms_per_step": wall_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, bool(peer), bool(peer & _lib.FLAG_ROW_INTERLEAVE)), "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {
                "bound": "fp64", "kernel": "k_trace_tile<double,symplectic,budget>", "achieved": ach, "peak": peak64,
                "unit": "TFLOP/s", "frac": ach / peak64,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch on the full frame, read from the committed ncu
                # export of THIS build (null when the export is absent or stale, never a pasted constant)
                "traffic": (prof["dram_bytes_per_launch"] if (prof and world == 1) else None),
                "traffic_note": prof_note if world == 1 else "per-rank launches cover a row block; the export is of the 1-GPU launch",
                # the 330 flop/step above is SURVEY's algorithmic (CSE'd) count; what the kernel EXECUTES per step comes from
                # the SASS mix of the same export (DFMA = 2, DMUL = DADD = 1), scaled by the live kernel time
                "executed_flop_per_step": prof["executed_fp64_flop_per_step"] if prof else None,
                "executed_flop_frac": (steps_per_launch * prof["executed_fp64_flop_per_step"] / (trace_ms * 1e-3) * 1e-12 / peak64)
                                      if (prof and world == 1) else None,
                "fp64_pipe_instructions_per_step": prof["fp64_pipe_instructions_per_step"] if prof else None,
                "pipe_active": (prof["fp64_pipe_active_pct"] / 100.0) if prof else None,
                "profile_export": os.path.relpath(PROFILE_EXPORT, ROOT) if prof else None, "build_info": build_info,
                "peak_source": "in-run DFMA micro-benchmark (gvt_measure_fma_peak; MEASURED_PEAKS.json has no "
                               "FP32/FP64 entry). The path is FMA-pipe bound, not HBM or tensor: ~0 B read and 16 B "
                               "written per pixel per 512 steps",
                "flop_per_step": fl, "steps_per_launch": int(steps_per_launch), "kernel_ms": trace_ms,
                "frac_of_fp32_peak": ach / peak32, "fp32_peak_tflops": peak32, "fp64_peak_tflops": peak64,
                "hbm_bytes_per_launch": int(16 * steps_per_launch // STEPS),   # one float4 store per pixel this rank produced
            },
            "cpu_baseline": cpu,
            "clocks": clocks,
            "extra": {
                "device": info, "trace_kernel_ms": trace_ms, "all_gather_ms": gather_ms,
                "natural_termination": {"steps_per_s": n_steps / (n_ms * 1e-3), "ms_per_frame": n_ms / len(n_stats),
                                        "steps_per_frame_rank0": int(s0.steps_committed),
                                        "census_rank0": {"horizon": int(s0.n_horizon), "escape": int(s0.n_escape),
                                                         "max_steps": int(s0.n_maxsteps), "disk_opaque": int(s0.n_disk)}},
                "f32_budget": {"steps_per_s": f_steps / (f_ms * 1e-3), "ms_per_frame": f_ms / len(f_stats),
                               "frac_of_fp32_peak": (f_steps / (f_ms * 1e-3)) * fl * 1e-12 / peak32 / world,
                               "note": "f32 arithmetic throughout: fast, but NOT within the 1e-6 tolerance of the f64 oracle "
                                       "(15 % of lit pixels outside it on this frame); reported as an accuracy / speed reference only"},
                "mixed_precision": {
                    "what": "GVT_PRECISION_MIXED: f64 state + f64 corrector of the implicit midpoint, f32 predictors (MUFU sin/cos) on "
                            "the outbound leg beyond 35 M; every pixel of this frame within 1e-6 relative of the all-f64 oracle "
                            "(tests/test_gpu_parity.py::test_headline_frame_every_pixel)",
                    "steps_per_s": m_steps / (m_ms * 1e-3), "ms_per_frame": m_ms / len(m_stats), "trace_kernel_ms": m_trace_ms,
                    "frac_of_fp32_peak": (m_steps / (m_ms * 1e-3)) * fl * 1e-12 / peak32 / world,
                    "speedup_vs_f64": (ev_ms / args.steps) / (m_ms / len(m_stats)),
                    "executed_fp64_flop_per_step": prof_m["executed_fp64_flop_per_step"] if prof_m else None,
                    "executed_fp32_flop_per_step": prof_m["executed_fp32_flop_per_step"] if prof_m else None,
                    "dram_bytes_per_launch": prof_m["dram_bytes_per_launch"] if (prof_m and world == 1) else None},
                "allgather": ag,
                **side,
            },
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        pinned.close()
    r.cleanup()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_extra(args):
    """BASELINE configs[3] and configs[4] at full size, reported as extra lines (not the headline metric)."""
    import math
    import gravitas_b200 as g
    from gravitas_b200 import _lib, camera, renderer as R
    rank, world, local, dist = dist_setup(args.gpus)
    nccl_id = None
    if world > 1:
        objs = [g.KerrRenderer.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(objs, src=0)
        nccl_id = objs[0]
    r = g.KerrRenderer(device=local, rank=rank, world_size=world, nccl_id=nccl_id)
    r.init()
    r.init_pipelines(mass=MASS, spin=SPIN, spec_w=SPEC_W, spec_h=SPEC_H, max_temp=TMAX)
    if args.workload == "config1":
        # BASELINE configs[0]: Schwarzschild a=0, 256x256 camera rays, 128 RKF45 steps, Boyer-Lindquist, through the
        # PhysicsEngine seam (batched integrate_ray_relativistic on the GPU) and, beside it, the CPU port
        import time as _t
        import numpy as np
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O
        Wx = Hx = 256
        cam, _ = camera.default_camera(Wx, Hx)
        opts = O.Options.default(max_steps=128)
        rp_o, keep = O.make_render_params(Wx, Hx, 1.0, 0.0, opts, coords=0)
        rays = np.array([O.camera_ray(cam, rp_o, x, y) for y in range(Hx) for x in range(Wx)])
        eng = g.PhysicsEngine(1.0, 0.0)
        prm = R.RenderParams(method=_lib.METHOD_RKF45, coords=_lib.COORDS_BL, step_rule=_lib.STEP_CONSTANT, max_steps=128)
        for _ in range(args.warmup):
            eng.integrate_rays(rays, prm)
        t0 = _t.perf_counter()
        for _ in range(args.steps):
            got = eng.integrate_rays(rays, prm)
        gpu_s = (_t.perf_counter() - t0) / args.steps
        t0 = _t.perf_counter()
        ref = O.integrate(1.0, 0.0, 0, opts, rays)
        cpu_s = _t.perf_counter() - t0
        steps = float(ref["steps"].sum())
        err = np.abs(got["xp"] - ref["xp"]) / np.maximum(np.abs(ref["xp"]), 1.0)
        if rank == 0:
            print(json.dumps({"extra_workload": "config 1: Schwarzschild a=0, 256x256 rays, 128 adaptive RKF45 steps, Boyer-Lindquist "
                              "(every ray ends MaxSteps; parity is on the 128-step state)", "n_gpus": 1,
                              "gpu_e2e_ms_host_arrays_in_out": gpu_s * 1e3, "gpu_steps_per_s_e2e": steps / gpu_s,
                              "cpu_port_ms": cpu_s * 1e3, "cpu_port_steps_per_s": steps / cpu_s,
                              "cpu_threads": O.lib().orc_num_threads(), "state_rel_err_median": float(np.median(err)),
                              "state_rel_err_max": float(err.max())}))
    elif args.workload == "config2":
        # BASELINE configs[1]: Kerr a*=0.999, 1920x1080, 256 fixed steps, f32
        Wx, Hx = 1920, 1080
        r.params = R.RenderParams(method=_lib.METHOD_SYMPLECTIC, precision=_lib.PRECISION_F32, max_steps=256,
                                  step_rule=_lib.STEP_WGSL, flags=_lib.FLAG_BUDGET)
        cam, _ = camera.default_camera(Wx, Hx)
        phys = R.pack_physics(MASS, SPIN, Wx, Hx)
        for _ in range(args.warmup):
            r.render(cam, phys, readback=False)
        ev_ms, wall_ms, stats = timed_frames(r, cam, phys, args.steps, dist, readback=False)
        steps = allreduce_sum(dist, float(sum(s.steps_executed for s in stats)))
        e_ms, _, _ = timed_frames(r, cam, phys, args.steps, dist, readback=True, out=r.pinned_frame(Wx, Hx))
        peak32, _ = r.measure_fma_peak(_lib.PRECISION_F32)
        if rank == 0:
            sps = steps / (ev_ms * 1e-3)
            print(json.dumps({"extra_workload": "config 2: Kerr a*=0.999, 1920x1080, 256 fixed implicit-midpoint steps, f32, budget "
                              "accounting", "n_gpus": world, "frames": args.steps, "ms_per_frame": ev_ms / args.steps,
                              "e2e_ms_per_frame": e_ms / args.steps, "steps_per_s": sps,
                              "frac_of_fp32_peak": sps * FLOP_PER_STEP["symplectic"] * 1e-12 / peak32 / world,
                              "fp32_peak_tflops": peak32, "fps": 1e3 * args.steps / ev_ms}))
    elif args.workload == "glsl":
        # the production WebGL2 shader's march: Cartesian Velocity-Verlet, f32, "ultra" budget of 256
        # steps (simulation.config.ts:205-211), MAX_DIST 10000 (physics.config.ts:60), at 1080p and at 4K
        res = {}
        for (Wx, Hx) in ((1920, 1080), (3840, 2160)):
            r.params = R.RenderParams(method=_lib.METHOD_VERLET_GLSL, precision=_lib.PRECISION_F32, max_steps=256,
                                      step_rule=_lib.STEP_CONSTANT, escape_radius=10000.0)
            cam, _ = camera.default_camera(Wx, Hx)
            phys = R.pack_physics(MASS, SPIN, Wx, Hx)
            for _ in range(args.warmup):
                r.render(cam, phys, readback=False)
            ev_ms, wall_ms, stats = timed_frames(r, cam, phys, args.steps, dist, readback=False)
            steps = allreduce_sum(dist, float(sum(s.steps_committed for s in stats)))
            res[f"{Wx}x{Hx}"] = {"ms_per_frame": ev_ms / args.steps, "fps": 1e3 * args.steps / ev_ms,
                                 "verlet_steps_per_s": steps / (ev_ms * 1e-3), "mean_steps_per_pixel": steps / args.steps / (Wx * Hx)}
        if rank == 0:
            print(json.dumps({"extra_workload": "GLSL-semantics path: Cartesian Velocity-Verlet on the shader's pseudo-Kerr acceleration "
                              "(fragment.glsl.ts:129-221), f32, 256-step budget, natural termination, thin-disk LUT shading",
                              "n_gpus": world, "frames": args.steps, **res}))
    elif args.workload == "webgl":
        # the whole production WebGL2 fragment shader (k_fragment_glsl: march + volumetric disk + jets + stars + glows
        # + ACES), "ultra-quality" preset (256-step budget), 4K, through WebGLRenderer.render(params, mouse); device
        # time per frame for the MUFU build a GLSL compiler would produce and, on one GPU, for the parity builds.
        # At N > 1: row-block shards + ncclAllGather, and row-interleaved shards + fused peer stores.
        from gravitas_b200 import webgl
        wid = None
        if world > 1:
            objs = [g.KerrRenderer.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(objs, src=0)
            wid = objs[0]
        w = webgl.WebGLRenderer(device=local, rank=rank, world_size=world, nccl_id=wid, noise_seed=11)
        assert w.init(), w.error
        Wx, Hx = 3840, 2160
        w.resize(Wx, Hx)
        sp = dict(mass=1.0, spin=0.9, zoom=30.0, lensing=1.0, features=dict(webgl.PRESETS["ultra-quality"], bloom=False))
        res = {}

        def run(nm, pr, flags):
            w.precision = pr
            tot, shader, steps = 0.0, 0.0, 0.0
            barrier(dist)
            for k in range(args.warmup + args.steps):
                w.render(sp, {"x": 0.5, "y": 0.5 + 7.0 / 180.0}, readback=False, flags=flags)
                if k >= args.warmup:
                    tot += w.last_stats.total_ms; shader += w.last_stats.trace_ms; steps += w.last_stats.steps_committed
            tot = allreduce_max(dist, tot) / args.steps
            res[nm] = {"ms_per_frame": tot, "fps": 1e3 / tot, "slowest_rank_shader_ms": allreduce_max(dist, shader) / args.steps,
                       "march_steps_per_s": allreduce_sum(dist, steps) / args.steps / (tot * 1e-3),
                       "mean_steps_per_pixel": allreduce_sum(dist, steps) / args.steps / (Wx * Hx)}

        if world == 1:
            for nm, pr in (("f32_fast", _lib.PRECISION_F32_FAST), ("f32_precise", _lib.PRECISION_F32), ("f64", _lib.PRECISION_F64)):
                run(nm, pr, 0)
            bl = []
            for k in range(args.warmup + args.steps):   # post tail on the last frame: bloom + final pass
                w._k.bloom(enabled=True, readback=False)
                if k >= args.warmup:
                    bl.append(w._k.last_bloom_ms)
            res["bloom_final_pass_ms"] = sum(bl) / len(bl)
        else:
            run("f32_fast_row_blocks_allgather", _lib.PRECISION_F32_FAST, 0)
            w._k.connect_peers(dist)
            run("f32_fast_row_blocks_peer_store", _lib.PRECISION_F32_FAST, _lib.FLAG_PEER_STORE)
            run("f32_fast_row_interleaved_peer_store", _lib.PRECISION_F32_FAST, _lib.FLAG_PEER_STORE | _lib.FLAG_ROW_INTERLEAVE)
            # the reference's default WebGL pipeline resolves every frame (ReprojectionManager): shader -> WebGL TAA
            w.taa = True
            w._k.reset_history()
            run("f32_fast_taa_row_blocks_peer_store", _lib.PRECISION_F32_FAST, _lib.FLAG_PEER_STORE)
            w._k.reset_history()
            run("f32_fast_taa_16row_stripes_peer_store", _lib.PRECISION_F32_FAST, _lib.FLAG_PEER_STORE | _lib.FLAG_ROW_INTERLEAVE)
            w.taa = False
        barrier(dist)
        w.cleanup()
        if rank == 0:
            print(json.dumps({"extra_workload": "WebGL2 production fragment shader (fragment.glsl.ts, all features of the "
                              "ultra-quality preset except bloom), 3840x2160, mouse/zoom camera at 97 deg polar, zoom 60",
                              "n_gpus": world, "frames": args.steps, **res}))
    elif args.workload == "config4":
        Wx, Hx = 7680, 4320
        r.params = R.RenderParams(method=_lib.METHOD_RKF45, max_steps=1024, step_rule=_lib.STEP_CONSTANT)
        cam, _ = camera.default_camera(Wx, Hx)
        phys = R.pack_physics(MASS, SPIN, Wx, Hx)
        for _ in range(args.warmup):
            r.render(cam, phys, readback=False)
        ev_ms, wall_ms, stats = timed_frames(r, cam, phys, args.steps, dist, readback=False)
        steps = allreduce_sum(dist, float(sum(s.steps_committed for s in stats)))
        rhs = allreduce_sum(dist, float(sum(s.rhs_evals for s in stats)))
        inter = None
        if world > 1:
            # the same frames with row-interleaved shards over the fused peer-store gather (natural
            # termination makes rows unequal; interleaving spreads them)
            r.resize(Wx, Hx)
            r.connect_peers(dist)
            r.params = R.RenderParams(method=_lib.METHOD_RKF45, max_steps=1024, step_rule=_lib.STEP_CONSTANT,
                                      flags=_lib.FLAG_PEER_STORE | _lib.FLAG_ROW_INTERLEAVE)
            r.render(cam, phys, readback=False)
            i_ms, _, i_stats = timed_frames(r, cam, phys, args.steps, dist, readback=False)
            t_max = allreduce_max(dist, sum(s.trace_ms for s in i_stats) / len(i_stats))
            inter = {"ms_per_frame": i_ms / args.steps, "slowest_rank_trace_ms": t_max}
        b_max = allreduce_max(dist, sum(s.trace_ms for s in stats) / len(stats))
        if rank == 0:
            print(json.dumps({"extra_workload": "config 4: Kerr a*=0.999, 7680x4320, <=1024 adaptive RKF45 steps (tol 1e-8, escape "
                              "1000), natural termination, row-block shard", "n_gpus": world, "frames": args.steps,
                              "slowest_rank_trace_ms": b_max, "row_interleaved_peer_store": inter,
                              "ms_per_frame": ev_ms / args.steps, "accepted_steps_per_s": steps / (ev_ms * 1e-3),
                              "rhs_evals_per_s": rhs / (ev_ms * 1e-3), "accepted_steps_per_frame": steps / args.steps,
                              "mean_steps_per_pixel": steps / args.steps / (Wx * Hx),
                              "all_gather_ms": sum(s.gather_ms for s in stats) / len(stats)}))
    else:
        Wx, Hx, frames = 3840, 2160, max(args.steps, 8)
        r.params = R.RenderParams(method=_lib.METHOD_SYMPLECTIC, max_steps=STEPS, step_rule=_lib.STEP_WGSL,
                                  flags=_lib.FLAG_TAA | _lib.FLAG_JITTER)
        prev, ev_ms, tot_steps, taa_ms, gather_ms = None, 0.0, 0.0, 0.0, 0.0
        barrier(dist)
        for k in range(args.warmup + frames):
            cam, vp = camera.default_camera(Wx, Hx, azimuth=math.pi + 0.005 * k, prev_view_proj=prev)
            phys = R.pack_physics(MASS, SPIN, Wx, Hx, frame_index=k)
            r.render(cam, phys, readback=False)
            prev = vp
            if k >= args.warmup:
                s = r.last_stats
                ev_ms += s.total_ms; tot_steps += s.steps_committed; taa_ms += s.taa_ms; gather_ms += s.gather_ms
        ev_ms = allreduce_max(dist, ev_ms)
        tot_steps = allreduce_sum(dist, tot_steps)
        if rank == 0:
            print(json.dumps({"extra_workload": "config 5: orbiting camera (azimuth += 0.005/frame), 3840x2160, 512 fixed steps, "
                              "Halton jitter + TAA resolve (ataa.wgsl.ts), natural termination; BASELINE asks for 240 frames, "
                              f"{frames} timed here", "n_gpus": world, "frames": frames, "ms_per_frame": ev_ms / frames,
                              "fps": 1e3 * frames / ev_ms, "steps_per_s": tot_steps / (ev_ms * 1e-3),
                              "taa_ms_per_frame": taa_ms / frames, "all_gather_ms": gather_ms / frames}))
    r.cleanup()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-peer-store", action="store_true",
                    help="N > 1: use the ncclAllGather after the trace kernel instead of the default fused gather (NVLink "
                         "peer stores from the trace kernel + 4-byte all-reduce barriers)")
    ap.add_argument("--shard", default="interleaved", choices=["blocks", "interleaved"],
                    help="N > 1 with the fused gather: rows dealt round-robin to the ranks (default: the zones of the march make "
                         "rows near the hole dearer even under budget accounting; measured on 8 GPUs 6.01 ms/frame, natural "
                         "termination 5.92 ms) or contiguous row blocks (6.12 / 6.22 ms)")
    ap.add_argument("--workload", default="config3", choices=["config1", "config2", "config3", "config4", "config5", "glsl", "webgl"],
                    help="config3 = the headline (default). The others print an 'extra_workload' JSON line for BASELINE "
                         "configs[0] (Schwarzschild 256x256x128 RKF45: GPU batch integrate + the CPU port), configs[1] "
                         "(1080p, 256 steps, f32), configs[3] (8K, 1024 adaptive RKF45), configs[4] (orbit, 4K, TAA)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "own" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "config3":
        run_extra(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
