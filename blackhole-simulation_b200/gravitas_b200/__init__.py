"""gravitas_b200 — host-side mirror of the reference's two seams over libgravitas_b200.so (C ABI, sm_100a CUDA).

* ``PhysicsEngine``  — same method names / argument meaning as the wasm-bindgen class of
  physics-engine/gravitas-wasm/src/lib.rs:42-465 (Seam A), including the SAB f32-offset protocol.
* ``KerrRenderer``   — the src/rendering renderer API (webgpu/renderer.ts:82-411: init / resize / render(camera,
  physics) / updateSettings), returning the frame buffer (Seam B).
* ``WebGLRenderer``  — the WebGL2 pipeline's renderer API (webgl/renderer.ts:36-482: init / resize / render(params,
  mouse) / cleanup) over the fused fragment-shader kernel.
* ``camera``         — gl-matrix-compatible orbit camera -> CameraUniforms (components/canvas/WebGPUCanvas.tsx:119-178).

The reference host is TypeScript over wasm-bindgen; node is not available in this image, so this ctypes mirror is
the host the tests and the bench drive (the N-API shim + TS drop-ins are in addon/ and INTEGRATION.md).
There is NO CPU fallback: importing works without a GPU (symbols resolve), but any compute call fails loudly
with ``GravitasError`` when no sm_100 device is present or the shared library is missing.
"""
from ._lib import GravitasError, lib, lib_path, OFFSETS  # noqa: F401
from .engine import PhysicsEngine, init_hooks  # noqa: F401
from .renderer import KerrRenderer, RenderParams, FrameStats, DeviceTarget, ExternalBuffer  # noqa: F401
from .webgl import WebGLRenderer  # noqa: F401
from . import camera, shard, webgl  # noqa: F401
