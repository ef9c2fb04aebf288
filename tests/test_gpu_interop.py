"""Frame targets in device memory (the hand-off without host memory). Wherever a frame entry point takes
a host buffer it also takes a device pointer; the frame must be the one the host path delivers, bit for bit, and no frame
bytes may be counted as D2H traffic. The external-memory import (cudaImportExternalMemory of a POSIX fd) is exercised
with an allocation exported by CUDA's own virtual-memory API (what a Vulkan / GL presenter would export is the same kind of
fd; neither API exists on the box). torch appears only as the owner of a device allocation."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SPIN = float(np.float32(0.999))
W, H = 160, 96


def _setup(renderer, **kw):
    from gravitas_b200 import camera, renderer as R
    renderer.init_pipelines(mass=1.0, spin=SPIN, spec_w=64, spec_h=16, max_temp=1e7)
    renderer.params = R.RenderParams(max_steps=96, step_rule=1, **kw)
    cam, _ = camera.default_camera(W, H)
    return cam, R.pack_physics(1.0, SPIN, W, H)


def test_device_target_equals_host_frame(renderer):
    import torch
    from gravitas_b200 import renderer as R, _lib
    cam, phys = _setup(renderer)
    host = np.array(renderer.render(cam, phys)).copy()
    assert host[..., :3].max() > 0
    # (1) storage format == output format: the producing kernel stores straight into the device target
    t = torch.full((H, W, 4), -1.0, dtype=torch.float32, device="cuda")
    assert renderer.render(cam, phys, out=R.DeviceTarget(t.data_ptr(), t.numel() * 4)) is None
    assert np.array_equal(t.cpu().numpy(), host)
    assert renderer.last_stats.d2h_bytes == 64 and renderer.last_stats.kernel_launches == 1     # counters only, no copy kernel
    # (2) another output format: one conversion pass + a device-to-device copy
    renderer.params.c.output_format = _lib.FORMAT_RGBA16F
    host16 = np.array(renderer.render(cam, phys)).copy()
    t16 = torch.zeros((H, W, 4), dtype=torch.float16, device="cuda")
    renderer.render(cam, phys, out=R.DeviceTarget(t16.data_ptr()))
    assert np.array_equal(t16.cpu().numpy().view(np.uint16), host16.view(np.uint16))
    assert renderer.last_stats.d2h_bytes == 64
    # (3) the RGBA16F-native frame chain: half4 stores from the producing kernel into the target
    renderer.set_frame_format(_lib.FORMAT_RGBA16F)
    try:
        host16n = np.array(renderer.render(cam, phys)).copy()
        t16.zero_()
        renderer.render(cam, phys, out=R.DeviceTarget(t16.data_ptr()))
        assert np.array_equal(t16.cpu().numpy().view(np.uint16), host16n.view(np.uint16))
        assert renderer.last_stats.kernel_launches == 1
    finally:
        renderer.set_frame_format(_lib.FORMAT_RGBA32F)
    # (4) TAA chain: the resolve kernel is the producing kernel
    renderer.params = R.RenderParams(max_steps=96, step_rule=1, flags=_lib.FLAG_TAA | _lib.FLAG_JITTER)
    renderer.reset_history()
    hosts = [np.array(renderer.render(cam, R.pack_physics(1.0, SPIN, W, H, frame_index=i))).copy() for i in range(3)]
    renderer.reset_history()
    for i in range(3):
        t.fill_(-1.0)
        renderer.render(cam, R.pack_physics(1.0, SPIN, W, H, frame_index=i), out=R.DeviceTarget(t.data_ptr()))
        assert np.array_equal(t.cpu().numpy(), hosts[i]), i
    # (5) read_frame and the bloom / final pass into device memory
    ref = renderer.read_frame()
    t.fill_(-1.0)
    renderer.read_frame(out=R.DeviceTarget(t.data_ptr()))
    assert np.array_equal(t.cpu().numpy(), ref)
    disp = renderer.bloom()
    t.fill_(-1.0)
    renderer.bloom(out=R.DeviceTarget(t.data_ptr()))
    assert np.array_equal(t.cpu().numpy(), disp)


def test_fragment_shader_into_device_target():
    import torch
    import gravitas_b200 as g
    from gravitas_b200 import renderer as R
    w = g.WebGLRenderer(); w.init(); w.resize(W, H)
    params, mouse = dict(mass=1.0, spin=0.9, zoom=30.0), {"x": 0.5, "y": 0.54}
    w.render(params, mouse)
    w.time = 0.0
    host = np.array(w.render(params, mouse)).copy()
    assert host[..., :3].max() > 0
    t = torch.full((H, W, 4), -1.0, dtype=torch.float32, device="cuda")
    ptr = t.data_ptr()
    w.time = 0.0
    keep = w._k.pinned_frame
    w._k.pinned_frame = lambda *_a, **_k: R.DeviceTarget(ptr)                # the frame buffer render() delivers into
    try:
        assert w.render(params, mouse) is None
    finally:
        w._k.pinned_frame = keep
    assert np.array_equal(t.cpu().numpy(), host)
    assert w.last_stats.d2h_bytes == 64
    w.cleanup()


def test_external_import_rejects_bad_handles(built):
    import gravitas_b200 as g
    from gravitas_b200 import _lib
    h, p = C.c_void_p(), C.c_void_p()
    assert built.lib().gvt_external_import_fd(0, -1, 4096, 0, C.byref(h), C.byref(p)) == _lib.GVT_ERR_INVALID
    assert built.lib().gvt_external_import_fd(0, 3, 0, 0, C.byref(h), C.byref(p)) == _lib.GVT_ERR_INVALID
    assert built.lib().gvt_external_import_fd(99, 3, 4096, 0, C.byref(h), C.byref(p)) == _lib.GVT_ERR_INVALID
    r_, w_ = os.pipe()                       # a valid fd that is no exported memory object: the driver refuses it
    try:
        rc = built.lib().gvt_external_import_fd(0, r_, 1 << 20, 0, C.byref(h), C.byref(p))
        assert rc == _lib.GVT_ERR_CUDA and not h.value and b"cudaImportExternalMemory" in built.lib().gvt_last_error()
        s = C.c_void_p()
        rc = built.lib().gvt_external_semaphore_import_fd(0, r_, 0, C.byref(s))
        assert rc == _lib.GVT_ERR_CUDA and not s.value
    finally:
        for fd in (r_, w_):
            try: os.close(fd)
            except OSError: pass
    assert built.lib().gvt_external_release(None) == 0 and built.lib().gvt_external_semaphore_release(None) == 0


def test_external_memory_fd_round_trip(renderer):
    """An allocation created and exported as a POSIX fd by CUDA's virtual-memory API, imported through
    gvt_external_import_fd (cudaImportExternalMemory) and used as the frame target; the exporter's own mapping of the
    same physical memory must then hold the host path's frame."""
    drv = pytest.importorskip("cuda.bindings.driver")
    from gravitas_b200 import renderer as R
    cam, phys = _setup(renderer)
    host = np.array(renderer.render(cam, phys)).copy()

    def ck(res):
        err, rest = res[0], res[1:]
        if err != drv.CUresult.CUDA_SUCCESS:
            raise RuntimeError(str(err))
        return rest[0] if len(rest) == 1 else rest
    ck(drv.cuInit(0))
    prop = drv.CUmemAllocationProp()
    prop.type = drv.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
    prop.location.type = drv.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    prop.location.id = 0
    prop.requestedHandleTypes = drv.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
    gran = ck(drv.cuMemGetAllocationGranularity(prop, drv.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM))
    nbytes = host.nbytes
    size = (nbytes + gran - 1) // gran * gran
    handle = ck(drv.cuMemCreate(size, prop, 0))
    fd = int(ck(drv.cuMemExportToShareableHandle(handle, drv.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0)))
    va = ck(drv.cuMemAddressReserve(size, 0, 0, 0))
    ck(drv.cuMemMap(va, size, 0, handle, 0))
    acc = drv.CUmemAccessDesc()
    acc.location.type = drv.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    acc.location.id = 0
    acc.flags = drv.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
    ck(drv.cuMemSetAccess(va, size, [acc], 1))
    ck(drv.cuMemsetD8(va, 0xFF, size))
    ext = None
    try:
        try:
            ext = R.ExternalBuffer.import_fd(fd, size)
        except Exception as e:               # a driver that maps only graphics-API exports through this path
            os.close(fd)
            pytest.skip(f"cudaImportExternalMemory refuses a CUDA-VMM fd on this driver ({e}); needs a Vulkan / GL exporter")
        assert renderer.render(cam, phys, out=ext) is None
        back = np.empty_like(host)
        ck(drv.cuMemcpyDtoH(back.ctypes.data, va, nbytes))
        assert np.array_equal(back, host)
        assert renderer.last_stats.d2h_bytes == 64
    finally:
        if ext is not None:
            ext.release()
        drv.cuMemUnmap(va, size); drv.cuMemAddressFree(va, size); drv.cuMemRelease(handle)
