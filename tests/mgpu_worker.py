"""torchrun worker for the multi-GPU parity test: every rank traces its row block and joins the single
ncclAllGather; rank 0 checks the gathered frame bit-for-bit against a one-GPU render of the whole frame."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blackhole-simulation_b200"))
import numpy as np
import torch.distributed as dist

import gravitas_b200 as g
from gravitas_b200 import _lib, camera, renderer as R, shard

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
local = int(os.environ.get("LOCAL_RANK", rank))
objs = [g.KerrRenderer.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(objs, src=0)
spin = float(np.float32(0.999))
multi = g.KerrRenderer(device=local, rank=rank, world_size=world, nccl_id=objs[0])
multi.init()
multi.init_pipelines(mass=1.0, spin=spin, spec_w=64, spec_h=16, max_temp=1e7)
single = g.KerrRenderer(device=local)
single.init()
single.init_pipelines(mass=1.0, spin=spin, spec_w=64, spec_h=16, max_temp=1e7)

for (W, H, steps) in ((256, 144, 128), (157, 83, 64)):       # divisible and ragged heights
    for taa in (False, True):
        flags = (_lib.FLAG_TAA | _lib.FLAG_JITTER) if taa else 0
        prev = None
        multi.resize(W, H); multi.reset_history(); single.resize(W, H); single.reset_history()
        for k in range(2):
            cam, vp = camera.default_camera(W, H, azimuth=math.pi + 0.005 * k, prev_view_proj=prev)
            phys = R.pack_physics(1.0, spin, W, H, frame_index=k)
            multi.params = R.RenderParams(max_steps=steps, flags=flags)
            single.params = R.RenderParams(max_steps=steps, flags=flags)
            a = np.array(multi.render(cam, phys))
            st = multi.last_stats
            b = np.array(single.render(cam, phys))
            assert (st.rows_begin, st.rows_end) == shard.shard_rows(H, rank, world), (st.rows_begin, st.rows_end)
            assert np.array_equal(a, b), f"rank {rank}: gathered frame differs from the single-GPU frame (W={W} H={H} taa={taa} k={k})"
            tot = np.array([float(st.steps_committed)])
            import torch
            t = torch.tensor(tot); dist.all_reduce(t)
            if not taa:
                assert int(t[0]) == int(single.last_stats.steps_committed)
            prev = vp
# fused gather (GVT_FLAG_PEER_STORE): NVLink peer stores from the producing kernel + a 4-byte all-reduce barrier give
# the same frames as the ncclAllGather, with and without TAA
for (W, H, steps) in ((256, 144, 96), (157, 83, 48)):
    for taa in (False, True):
        flags = _lib.FLAG_PEER_STORE | ((_lib.FLAG_TAA | _lib.FLAG_JITTER) if taa else 0)
        multi.resize(W, H); multi.reset_history(); single.resize(W, H); single.reset_history()
        multi.connect_peers(dist)
        prev = None
        for k in range(3):
            cam, vp = camera.default_camera(W, H, azimuth=math.pi + 0.005 * k, prev_view_proj=prev)
            phys = R.pack_physics(1.0, spin, W, H, frame_index=k)
            multi.params = R.RenderParams(max_steps=steps, flags=flags)
            single.params = R.RenderParams(max_steps=steps, flags=flags & ~_lib.FLAG_PEER_STORE)
            a = np.array(multi.render(cam, phys))
            assert multi.last_stats.gather_ms >= 0
            b = np.array(single.render(cam, phys))
            assert np.array_equal(a, b), f"rank {rank}: peer-store frame differs (W={W} H={H} taa={taa} k={k})"
            prev = vp
dist.barrier()

# one shared host frame assembled by all ranks (GVT_FLAG_D2H_OWN_ROWS): equals the single-GPU frame on every rank
W, H, steps = 256, 144, 64
names = [None]
if rank == 0:
    shared = R.SharedFrame(W, H)
    names = [shared.name]
dist.broadcast_object_list(names, src=0)
if rank != 0:
    shared = R.SharedFrame(W, H, name=names[0], create=False)
cam, _ = camera.default_camera(W, H)
phys = R.pack_physics(1.0, spin, W, H)
multi.params = R.RenderParams(max_steps=steps, flags=_lib.FLAG_D2H_OWN_ROWS)
single.params = R.RenderParams(max_steps=steps)
multi.render(cam, phys, out=shared)
r0, r1 = shard.shard_rows(H, rank, world)
assert multi.last_stats.d2h_bytes == (r1 - r0) * W * 16 + 64, multi.last_stats.d2h_bytes
dist.barrier()
b = np.array(single.render(cam, phys))
assert np.array_equal(np.array(shared.array()), b), f"rank {rank}: shared host frame differs"
dist.barrier()
shared.close()
multi.cleanup(); single.cleanup()

# the WebGL2 fragment-shader path shares the pipeline: sharded rows + all-gather / fused peer stores, with and without
# the WebGL TAA resolve, bit-identical to the single-GPU frames
from gravitas_b200 import webgl
ids = [g.KerrRenderer.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
wm = webgl.WebGLRenderer(device=local, rank=rank, world_size=world, nccl_id=ids[0], noise_seed=3)
ws = webgl.WebGLRenderer(device=local, noise_seed=3)
assert wm.init() and ws.init(), (wm.error, ws.error)
sp = dict(mass=1.0, spin=0.9, zoom=30.0, lensing=1.0, features=dict(webgl.PRESETS["high-quality"], bloom=False))
for (W, H) in ((192, 108), (157, 83)):
    for taa in (False, True):
        for peer in (False, True):
            wm.resize(W, H); ws.resize(W, H)
            wm._k.reset_history(); ws._k.reset_history()
            if peer:
                wm._k.connect_peers(dist)
            wm.taa = ws.taa = taa
            wm.time = ws.time = 0.0
            for k in range(2):
                mouse = {"x": 0.5 + 0.001 * k, "y": 0.54}
                a = np.array(wm.render(sp, mouse, flags=_lib.FLAG_PEER_STORE if peer else 0))
                b = np.array(ws.render(sp, mouse))
                assert (wm.last_stats.rows_begin, wm.last_stats.rows_end) == shard.shard_rows(H, rank, world)
                assert np.array_equal(a, b), f"rank {rank}: fragment-shader frame differs (W={W} H={H} taa={taa} peer={peer} k={k})"
# row-interleaved shards (GVT_FLAG_ROW_INTERLEAVE, peer stores): same frames again, on both producing kernels
for (W, H) in ((192, 108), (157, 83)):
    wm.resize(W, H); ws.resize(W, H)
    wm._k.connect_peers(dist)
    wm.taa = ws.taa = False
    wm.time = ws.time = 0.0
    a = np.array(wm.render(sp, {"x": 0.5, "y": 0.54}, flags=_lib.FLAG_PEER_STORE | _lib.FLAG_ROW_INTERLEAVE))
    b = np.array(ws.render(sp, {"x": 0.5, "y": 0.54}))
    assert (wm.last_stats.rows_begin, wm.last_stats.rows_end) == (rank, H)
    assert np.array_equal(a, b), f"rank {rank}: interleaved fragment-shader frame differs (W={W} H={H})"
    wm._k.reset_history(); ws._k.reset_history()
    wm.taa = ws.taa = True                      # WebGL TAA over 16-row stripes + halo
    wm.time = ws.time = 0.0
    for k in range(3):
        mouse = {"x": 0.5 + 0.001 * k, "y": 0.54}
        a = np.array(wm.render(sp, mouse, flags=_lib.FLAG_PEER_STORE | _lib.FLAG_ROW_INTERLEAVE))
        b = np.array(ws.render(sp, mouse))
        assert np.array_equal(a, b), f"rank {rank}: striped fragment-shader TAA frame differs (W={W} H={H} k={k})"
dist.barrier()
wm.cleanup(); ws.cleanup()
ids = [g.KerrRenderer.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
multi = g.KerrRenderer(device=local, rank=rank, world_size=world, nccl_id=ids[0]); multi.init()
multi.init_pipelines(mass=1.0, spin=spin, spec_w=64, spec_h=16, max_temp=1e7)
single = g.KerrRenderer(device=local); single.init()
single.init_pipelines(mass=1.0, spin=spin, spec_w=64, spec_h=16, max_temp=1e7)
for (W, H, method) in ((256, 144, _lib.METHOD_SYMPLECTIC), (157, 83, _lib.METHOD_RKF45)):
    multi.resize(W, H); single.resize(W, H)
    multi.connect_peers(dist)
    cam, _ = camera.default_camera(W, H)
    phys = R.pack_physics(1.0, spin, W, H)
    kw = dict(method=method, max_steps=96, step_rule=_lib.STEP_WGSL if method == _lib.METHOD_SYMPLECTIC else _lib.STEP_CONSTANT)
    multi.params = R.RenderParams(flags=_lib.FLAG_PEER_STORE | _lib.FLAG_ROW_INTERLEAVE, **kw)
    single.params = R.RenderParams(**kw)
    a = np.array(multi.render(cam, phys))
    b = np.array(single.render(cam, phys))
    assert np.array_equal(a, b), f"rank {rank}: interleaved trace frame differs (W={W} H={H} method={method})"
    t = torch.tensor([float(multi.last_stats.steps_committed)]); dist.all_reduce(t)
    assert int(t[0]) == int(single.last_stats.steps_committed)
    # the flag is refused without the peer-store gather (an all-gather needs contiguous blocks)
    multi.params = R.RenderParams(flags=_lib.FLAG_ROW_INTERLEAVE, **kw)
    try:
        multi.render(cam, phys)
        raise AssertionError("GVT_FLAG_ROW_INTERLEAVE accepted without GVT_FLAG_PEER_STORE")
    except g.GravitasError as e:
        assert e.code == _lib.GVT_ERR_INVALID
    # with TAA the stripes are 16 rows + halo: three frames of an orbiting, jittered camera equal the single-GPU frames
    multi.reset_history(); single.reset_history()
    prev = None
    for k in range(3):
        camk, vp = camera.default_camera(W, H, azimuth=math.pi + 0.005 * k, prev_view_proj=prev)
        physk = R.pack_physics(1.0, spin, W, H, frame_index=k)
        fl = _lib.FLAG_TAA | _lib.FLAG_JITTER
        multi.params = R.RenderParams(flags=fl | _lib.FLAG_PEER_STORE | _lib.FLAG_ROW_INTERLEAVE, **kw)
        single.params = R.RenderParams(flags=fl, **kw)
        a = np.array(multi.render(camk, physk))
        b = np.array(single.render(camk, physk))
        assert np.array_equal(a, b), f"rank {rank}: striped TAA frame differs (W={W} H={H} method={method} k={k})"
        prev = vp
# own-row delivery of interleaved shards into host frames: a page-locked shared frame (the kernel stores straight into
# it) and a pageable private buffer (strided cudaMemcpy2D of this rank's rows)
import ctypes as C
W, H = 157, 83
multi.resize(W, H); single.resize(W, H)
multi.connect_peers(dist)
cam, _ = camera.default_camera(W, H)
phys = R.pack_physics(1.0, spin, W, H)
kw = dict(method=_lib.METHOD_SYMPLECTIC, max_steps=64, step_rule=_lib.STEP_WGSL)
single.params = R.RenderParams(**kw)
b = np.array(single.render(cam, phys))
multi.params = R.RenderParams(flags=_lib.FLAG_PEER_STORE | _lib.FLAG_ROW_INTERLEAVE | _lib.FLAG_D2H_OWN_ROWS, **kw)
names = [None]
if rank == 0:
    shared = R.SharedFrame(W, H)
    names = [shared.name]
dist.broadcast_object_list(names, src=0)
if rank != 0:
    shared = R.SharedFrame(W, H, name=names[0], create=False)
multi.render(cam, phys, out=shared)
dist.barrier()
assert np.array_equal(np.array(shared.array()), b), f"rank {rank}: interleaved shared host frame differs"
dist.barrier()
shared.close()


class Pageable:   # an ordinary (not page-locked) host buffer
    def __init__(self, shape):
        self.a = np.full(shape, -1.0, np.float32)
        self.ptr = C.c_void_p(self.a.ctypes.data)
    def array(self, dtype, shape):
        return self.a


pg = Pageable((H, W, 4))
multi.render(cam, phys, out=pg)
mine = shard.interleaved_rows(H, rank, world)
others = [y for y in range(H) if y not in mine]
assert np.array_equal(pg.a[mine], b[mine]), f"rank {rank}: own interleaved rows differ in the pageable buffer"
assert np.all(pg.a[others] == -1.0)                       # nothing but this rank's rows was touched
assert multi.last_stats.d2h_bytes == len(mine) * W * 16 + 64
dist.barrier()

# BASELINE configs[3] at its stated size through the ranks: 7680x4320, <= 1024 adaptive RKF45 steps, natural termination,
# row-interleaved shards over the fused peer-store gather -- bit-identical to the one-GPU frame, same step census.
# And the headline frame (4K x 512) in both precision modes the bench reports, row blocks + ncclAllGather.
import time
for (W, H, kw, flags, what) in (
        (7680, 4320, dict(method=_lib.METHOD_RKF45, max_steps=1024, step_rule=_lib.STEP_CONSTANT), _lib.FLAG_PEER_STORE | _lib.FLAG_ROW_INTERLEAVE,
         "config 4 (8K, <=1024 RKF45), row-interleaved + peer stores"),
        (3840, 2160, dict(method=_lib.METHOD_SYMPLECTIC, max_steps=512, step_rule=_lib.STEP_WGSL), 0, "config 3 (4K x 512, f64), row blocks + ncclAllGather"),
        (3840, 2160, dict(method=_lib.METHOD_SYMPLECTIC, max_steps=512, step_rule=_lib.STEP_WGSL, precision=_lib.PRECISION_MIXED), _lib.FLAG_PEER_STORE,
         "config 3 (4K x 512, mixed), row blocks + peer stores")):
    multi.resize(W, H); single.resize(W, H)
    if flags & _lib.FLAG_PEER_STORE:
        multi.connect_peers(dist)
    cam, _ = camera.default_camera(W, H)
    phys = R.pack_physics(1.0, spin, W, H)
    multi.params = R.RenderParams(flags=flags, **kw)
    single.params = R.RenderParams(**kw)
    a = np.array(multi.render(cam, phys))
    ms_multi = multi.last_stats.total_ms
    b = np.array(single.render(cam, phys))
    assert np.array_equal(a, b), f"rank {rank}: {what}: frame differs from the single-GPU frame"
    t = torch.tensor([float(multi.last_stats.steps_committed)], dtype=torch.float64); dist.all_reduce(t)   # ~6e9: beyond float32
    assert int(t[0]) == int(single.last_stats.steps_committed), (what, int(t[0]), int(single.last_stats.steps_committed))
    if rank == 0:
        print(f"[mgpu] {what}: {world}-rank frame bit-identical to 1 GPU ({W}x{H}, {int(t[0])} steps; {ms_multi:.1f} ms vs "
              f"{single.last_stats.total_ms:.1f} ms on one GPU)", flush=True)
dist.barrier()

# the RGBA16F-native frame chain across ranks: half4 peer stores / half-sized all-gather, TAA history in f16 on every rank
for flags_extra, name in ((0, "ncclAllGather"), (_lib.FLAG_PEER_STORE, "peer stores")):
    W, H, steps = 256, 144, 96
    for rr in (multi, single):
        rr.set_frame_format(_lib.FORMAT_RGBA16F)
        rr.resize(W, H); rr.reset_history()
    if flags_extra:
        multi.connect_peers(dist)
    prev = None
    for k in range(3):
        cam, vp = camera.default_camera(W, H, azimuth=math.pi + 0.005 * k, prev_view_proj=prev)
        phys = R.pack_physics(1.0, spin, W, H, frame_index=k)
        fl = _lib.FLAG_TAA | _lib.FLAG_JITTER
        multi.params = R.RenderParams(max_steps=steps, flags=fl | flags_extra, output_format=_lib.FORMAT_RGBA16F)
        single.params = R.RenderParams(max_steps=steps, flags=fl, output_format=_lib.FORMAT_RGBA16F)
        a = np.array(multi.render(cam, phys)); b = np.array(single.render(cam, phys))
        assert a.dtype == np.float16 and np.array_equal(a.view(np.uint16), b.view(np.uint16)), f"rank {rank}: RGBA16F chain differs ({name}, k={k})"
        prev = vp
    if rank == 0:
        print(f"[mgpu] RGBA16F frame chain + TAA, {name}: {world}-rank frames bit-identical to 1 GPU", flush=True)
for rr in (multi, single):
    rr.set_frame_format(_lib.FORMAT_RGBA32F)
dist.barrier()
multi.cleanup(); single.cleanup()
dist.destroy_process_group()
print(f"rank {rank} ok")
