"""KerrRenderer — the src/rendering renderer / frame-buffer API over gvt_render_* (Seam B).

Mirrors WebGPURenderer (rendering/webgpu/renderer.ts:82-411): ``init`` -> ``init_pipelines(max_steps)`` ->
``resize(w, h)`` -> ``render(camera, physics)``; ``render`` returns the finished frame buffer instead of presenting
to a canvas."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (GvtCamera, GvtDeviceConfig, GvtFrameStats, GvtPhysicsParams, GvtRenderParams, check, lib)


class RenderParams:
    """Integration options of the trace (IntegrationOptions, geodesic/integrator.rs:24-47 + WGSL overrides)."""

    def __init__(self, **kw):
        self.c = GvtRenderParams()
        check(lib().gvt_render_params_default(C.byref(self.c)))
        for k, v in kw.items():
            if not hasattr(self.c, k):
                raise AttributeError(k)
            setattr(self.c, k, v)

    def __getattr__(self, k):
        return getattr(self.__dict__["c"], k)


class FrameStats(dict):
    __getattr__ = dict.__getitem__


def _stats(s):
    return FrameStats({f[0]: getattr(s, f[0]) for f in GvtFrameStats._fields_})


def pack_camera(cam88):
    cam88 = np.ascontiguousarray(cam88, dtype=np.float32)
    if cam88.size != 88:
        raise ValueError("CameraUniforms is 88 floats (352 bytes, types/webgpu.ts:25)")
    c = GvtCamera()
    C.memmove(C.byref(c), cam88.ctypes.data, 352)
    return c


def pack_physics(mass, spin, width, height, time=0.0, dt=0.016, frame_index=0):
    p = GvtPhysicsParams()
    p.mass, p.spin = mass, spin
    p.resolution[0], p.resolution[1] = float(width), float(height)
    p.time, p.dt, p.frame_index = time, dt, frame_index
    return p


class PinnedBuffer:
    """Page-locked host frame buffer (what an N-API external ArrayBuffer would wrap)."""

    def __init__(self, nbytes):
        self.ptr = C.c_void_p()
        self.nbytes = nbytes
        check(lib().gvt_host_alloc(nbytes, C.byref(self.ptr)))

    def array(self, dtype, shape):
        """Zero-copy view of the page-locked buffer. The view keeps this PinnedBuffer alive (its base object holds a
        reference), so a view held by the caller never outlives the memory it points into; the NEXT frame rendered into
        the same buffer overwrites it, as a canvas would -- copy (np.array(view)) to keep a frame."""
        buf = (C.c_uint8 * self.nbytes).from_address(self.ptr.value)
        buf._owner = self
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr:
            lib().gvt_host_free(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:      # interpreter shutdown: the module globals may already be gone
            pass


class DeviceTarget:
    """A frame target in DEVICE memory: a raw device pointer of the caller -- a buffer of the presenter's
    graphics API mapped through `ExternalBuffer.import_fd`, or any device allocation. Passed as `out=` to the frame calls;
    the producing kernel stores into it (or one device-to-device copy follows) and nothing comes back to the host."""

    def __init__(self, ptr, nbytes=0):
        self.ptr = C.c_void_p(int(ptr))
        self.nbytes = nbytes

    def array(self, dtype, shape):
        return None                      # the frame stays on the device


class ExternalBuffer(DeviceTarget):
    """Memory another API exported as a POSIX fd (VK_KHR_external_memory_fd / GL_EXT_memory_object_fd / a CUDA VMM
    allocation), mapped with cudaImportExternalMemory. On success the driver owns the fd."""

    def __init__(self, handle, ptr, nbytes):
        super().__init__(ptr, nbytes)
        self._h = handle

    @classmethod
    def import_fd(cls, fd, nbytes, device=0, dedicated=False):
        h, p = C.c_void_p(), C.c_void_p()
        check(lib().gvt_external_import_fd(device, fd, nbytes, 1 if dedicated else 0, C.byref(h), C.byref(p)))
        return cls(h, p.value, nbytes)

    def release(self):
        if self._h:
            lib().gvt_external_release(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class SharedFrame:
    """One host frame buffer shared by all ranks of a box (POSIX shared memory, page-locked in every process): with
    GVT_FLAG_D2H_OWN_ROWS each rank copies its row block straight into it, so the consumer (rank 0's host) gets the
    assembled frame for one frame's worth of PCIe traffic in total."""

    def __init__(self, width, height, fmt=_lib.FORMAT_RGBA32F, name=None, create=True):
        from multiprocessing import shared_memory
        self.nbytes = width * height * (16 if fmt == _lib.FORMAT_RGBA32F else 8)
        self.shape, self.dtype = (height, width, 4), (np.float32 if fmt == _lib.FORMAT_RGBA32F else np.float16)
        self.shm = shared_memory.SharedMemory(name=name, create=create, size=self.nbytes)
        self.owner = create
        if not create:   # only the creating rank unlinks; keep the other ranks' resource trackers out of it
            try:
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self._view = np.ndarray(self.shape, dtype=self.dtype, buffer=self.shm.buf)
        self.ptr = C.c_void_p(self._view.ctypes.data)
        check(lib().gvt_host_register(self.ptr, self.nbytes))

    @property
    def name(self):
        return self.shm.name

    def array(self, dtype=None, shape=None):
        return self._view

    def close(self):
        if self.shm is not None:
            lib().gvt_host_unregister(self.ptr)
            self._view = None
            self.shm.close()
            if self.owner:
                self.shm.unlink()
            self.shm = None


class KerrRenderer:
    def __init__(self, device=0, rank=0, world_size=1, nccl_id=None):
        self._h = C.c_void_p()
        self.device, self.rank, self.world_size = device, rank, world_size
        self._nccl_id = nccl_id
        self.params = RenderParams()
        self.width = self.height = 0
        self.last_stats = None
        self._pinned = None

    # WebGPURenderer.init (renderer.ts:82-136): acquire the device
    def init(self):
        cfg = GvtDeviceConfig()
        cfg.struct_size = C.sizeof(cfg)
        cfg.device, cfg.rank, cfg.world_size = self.device, self.rank, self.world_size
        if self.world_size > 1:
            if self._nccl_id is None or len(self._nccl_id) != 128:
                raise ValueError("world_size > 1 needs the 128-byte ncclUniqueId from nccl_unique_id() on rank 0")
            C.memmove(cfg.nccl_id, bytes(self._nccl_id), 128)
        check(lib().gvt_render_create(C.byref(cfg), C.byref(self._h)))
        return True

    @staticmethod
    def nccl_unique_id():
        out = (C.c_uint8 * 128)()
        check(lib().gvt_nccl_unique_id(out))
        return bytes(out)

    # initPipelines(format, maxSteps) (renderer.ts:186-254) + SpectralManager.initialize (spectral.ts:21-61)
    def init_pipelines(self, max_steps=None, mass=1.0, spin=0.999, spec_w=256, spec_h=32, max_temp=1e7):
        if max_steps is not None:
            self.params.c.max_steps = int(max_steps)
        check(lib().gvt_render_init_luts(self._h, mass, spin, spec_w, spec_h, max_temp))

    def set_luts(self, spectrum, spec_w, spec_h, tdisk, rin, rout):
        spectrum = np.ascontiguousarray(spectrum, np.float32)
        tdisk = np.ascontiguousarray(tdisk, np.float32)
        pf = C.POINTER(C.c_float)
        check(lib().gvt_render_set_luts(self._h, spectrum.ctypes.data_as(pf), spec_w, spec_h, tdisk.ctypes.data_as(pf),
                                        tdisk.size, rin, rout))

    def update_settings(self, max_steps):  # renderer.ts:256-263
        self.params.c.max_steps = int(max_steps)

    def get_format(self):  # renderer.ts:265-267 getFormat(): the frame-buffer format delivered to the host
        return {_lib.FORMAT_RGBA32F: "rgba32float", _lib.FORMAT_RGBA16F: "rgba16float"}.get(self.params.c.output_format, "rgba8unorm")

    def resize(self, width, height):  # renderer.ts:269-278
        check(lib().gvt_render_resize(self._h, int(width), int(height)))
        self.width, self.height = int(width), int(height)

    def connect_peers(self, dist):
        """GVT_FLAG_PEER_STORE set-up: exchange the CUDA IPC handles of the two frame buffers between all ranks (through
        the host's process group — plumbing only) and map every peer's pair. Call after resize()."""
        mine = (C.c_uint8 * 128)()
        rc = lib().gvt_render_export_frames(self._h, mine)
        allh = [None] * self.world_size
        dist.all_gather_object(allh, bytes(mine) if rc == 0 else None)   # every rank takes part in the exchange even on error
        check(rc)
        err = None
        for p, hb in enumerate(allh):
            if p != self.rank and err is None:
                if hb is None:
                    err = _lib.GravitasError(_lib.GVT_ERR_CUDA, f"rank {p} could not export its frames")
                    break
                buf = (C.c_uint8 * 128).from_buffer_copy(hb)
                rc = lib().gvt_render_import_peer_frames(self._h, p, buf)
                if rc != 0:
                    msg = lib().gvt_last_error()
                    err = _lib.GravitasError(rc, msg.decode() if msg else "import failed")
        dist.barrier()
        if err is not None:
            raise err

    def pinned_frame(self, width, height, fmt=_lib.FORMAT_RGBA32F):
        nbytes = width * height * _lib.FORMAT_BYTES[fmt]
        if self._pinned is None or self._pinned.nbytes != nbytes:
            self._pinned = PinnedBuffer(nbytes)
        return self._pinned

    def render(self, camera, physics, out=None, readback=True):
        """render(camera: CameraUniforms[88 f32], physics: PhysicsParams) (renderer.ts:280).

        Returns the frame (H, W, 4) float32 (or float16 for RGBA16F) when ``readback``; with ``readback=False`` the
        frame stays in device memory (use ``read_frame``). ``out`` may be a PinnedBuffer to receive the frame."""
        cam = camera if isinstance(camera, GvtCamera) else pack_camera(camera)
        W, H = int(physics.resolution[0]), int(physics.resolution[1])
        st = GvtFrameStats()
        host = None
        if readback:
            buf = out if out is not None else self.pinned_frame(W, H, self.params.c.output_format)
            host = buf.ptr
        check(lib().gvt_render_frame(self._h, C.byref(cam), C.byref(physics), C.byref(self.params.c), host, C.byref(st)))
        self.width, self.height = W, H
        self.last_stats = _stats(st)
        if not readback:
            return None
        return buf.array(np.dtype(_lib.FORMAT_DTYPE[self.params.c.output_format]), (H, W, 4))

    def render_rows(self, camera, physics, row0, row1):
        """One row block of the frame (what one rank of an N-GPU run traces); the frame stays in device memory."""
        cam = camera if isinstance(camera, GvtCamera) else pack_camera(camera)
        st = GvtFrameStats()
        check(lib().gvt_render_rows(self._h, C.byref(cam), C.byref(physics), C.byref(self.params.c), row0, row1, None, C.byref(st)))
        self.width, self.height = int(physics.resolution[0]), int(physics.resolution[1])
        self.last_stats = _stats(st)
        return self.last_stats

    def read_frame(self, fmt=_lib.FORMAT_RGBA32F, out=None):
        if out is not None:                       # a DeviceTarget (or any object with .ptr): the frame goes there
            check(lib().gvt_render_read_frame(self._h, fmt, out.ptr))
            return None
        out = np.zeros((self.height, self.width, 4), np.dtype(_lib.FORMAT_DTYPE[fmt]))
        check(lib().gvt_render_read_frame(self._h, fmt, out.ctypes.data_as(C.c_void_p)))
        return out

    def trace_states(self, camera, physics, x0=0, xs=1, y0=0, y1=None, ys=1):
        """Parity hook: per-pixel final (x,p), termination, steps, max|H|, f64 RGBA over a pixel lattice."""
        cam = camera if isinstance(camera, GvtCamera) else pack_camera(camera)
        W, H = int(physics.resolution[0]), int(physics.resolution[1])
        if y1 is None:
            y1 = H
        nx, ny = max((W - x0 + xs - 1) // xs, 0), max((y1 - y0 + ys - 1) // ys, 0)   # the library validates the lattice
        xp = np.zeros((ny, nx, 8))
        term = np.zeros((ny, nx), np.uint32)
        steps = np.zeros((ny, nx), np.uint32)
        drift = np.zeros((ny, nx))
        rgba = np.zeros((ny, nx, 4))
        pd, pu = C.POINTER(C.c_double), C.POINTER(C.c_uint32)
        check(lib().gvt_trace_states(self._h, C.byref(cam), C.byref(physics), C.byref(self.params.c), x0, xs, y0, y1, ys,
                                     xp.ctypes.data_as(pd), term.ctypes.data_as(pu), steps.ctypes.data_as(pu),
                                     drift.ctypes.data_as(pd), rgba.ctypes.data_as(pd)))
        return dict(xp=xp, term=term, steps=steps, drift=drift, rgba=rgba)

    def _taa(self, cam, cur, hist, webgl, blend, moving, precise):
        """Host-frame entry of both resolves; float16 inputs run the RGBA16F flavour of the kernels (8 B per pixel)."""
        cur = np.ascontiguousarray(cur)
        f16 = cur.dtype == np.float16
        dt = np.float16 if f16 else np.float32
        cur = np.ascontiguousarray(cur, dt)
        hist = np.ascontiguousarray(hist, dt)
        H, W = cur.shape[:2]
        out = np.zeros_like(cur)
        ms = C.c_double(0.0)
        check(lib().gvt_taa_resolve_ex(self._h, C.byref(cam) if cam is not None else None, W, H, cur.ctypes.data_as(C.c_void_p),
                                       hist.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                       _lib.FORMAT_RGBA16F if f16 else _lib.FORMAT_RGBA32F, 1 if webgl else 0, float(blend),
                                       1 if moving else 0, 1 if precise else 0, C.byref(ms)))
        self.last_taa_ms = ms.value
        return out

    def taa_resolve(self, camera, cur, hist, precise=False):
        """ataa.wgsl.ts on host frames (float32, or float16 = RGBA16F textures). precise=True: the validation build."""
        cam = camera if isinstance(camera, GvtCamera) else pack_camera(camera)
        return self._taa(cam, cur, hist, False, 0.0, False, precise)

    def taa_resolve_webgl(self, cur, hist, blend=0.75, camera_moving=False, precise=False):
        """ReprojectionManager.resolve (rendering/reprojection.ts:195-272) semantics on host frames."""
        return self._taa(None, cur, hist, True, blend, camera_moving, precise)

    def set_frame_format(self, fmt):
        """RGBA32F (default, parity format) or RGBA16F (the reference's texture format) for cur / frame / hist."""
        check(lib().gvt_render_set_frame_format(self._h, fmt))

    def bloom(self, enabled=True, intensity=0.5, threshold=0.8, blur_passes=2, fmt=_lib.FORMAT_RGBA32F, readback=True,
              precise=False, out=None):
        """BloomManager.applyBloomToTexture / drawTextureToScreen (rendering/bloom.ts:446-632) on the finished frame:
        returns the display-referred frame (ACES + gamma applied) in ``fmt``. precise=True: the validation build."""
        cfg = _lib.GvtBloomConfig()
        cfg.struct_size = C.sizeof(cfg)
        cfg.enabled, cfg.intensity, cfg.threshold, cfg.blur_passes = 1 if enabled else 0, intensity, threshold, blur_passes
        cfg.precise = 1 if precise else 0
        ms = C.c_double()
        if out is not None:                       # a DeviceTarget: the display-referred frame stays on the device
            check(lib().gvt_render_bloom(self._h, C.byref(cfg), fmt, out.ptr, C.byref(ms)))
            self.last_bloom_ms = ms.value
            return None
        out = np.zeros((self.height, self.width, 4), np.dtype(_lib.FORMAT_DTYPE[fmt])) if readback else None
        check(lib().gvt_render_bloom(self._h, C.byref(cfg), fmt, out.ctypes.data_as(C.c_void_p) if readback else None, C.byref(ms)))
        self.last_bloom_ms = ms.value
        return out

    def reset_history(self):
        check(lib().gvt_render_reset_history(self._h))

    def measure_fma_peak(self, precision):
        t, ms = C.c_double(), C.c_double()
        check(lib().gvt_measure_fma_peak(self._h, precision, C.byref(t), C.byref(ms)))
        return t.value, ms.value

    def device_info(self):
        sm, ma, mi = C.c_int32(), C.c_int32(), C.c_int32()
        name = C.create_string_buffer(256)
        check(lib().gvt_device_info(self._h, C.byref(sm), C.byref(ma), C.byref(mi), name))
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), name=name.value.decode())

    def cleanup(self):  # WebGLRenderer.cleanup (webgl/renderer.ts:471)
        if self._h:
            lib().gvt_render_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.cleanup()
        except Exception:      # interpreter shutdown: the module globals may already be gone
            pass
