"""The `blackhole-physics` MODULE shape (what `import("blackhole-physics")` resolves to in the reference:
wasm-bindgen `--target web` output of gravitas-wasm) — the Python twin of addon/ts/index.ts, statement for statement, so
that the memory contract the unchanged worker relies on can be replayed in a test without node:

  * ``init()`` returns an object whose ``.memory.buffer`` is ONE shared byte buffer (the wasm linear memory's stand-in);
  * every ``PhysicsEngine(m, a)`` is attached to its own 2048-f32 region of that buffer;
  * ``get_sab_ptr()`` returns that region's BYTE OFFSET inside ``memory.buffer`` (physics.worker.ts:153-163 and
    physics-bridge.ts:107 divide it by 4 / add it to build their Float32Array views).
"""
import types

import numpy as np

from . import engine as _engine

REGION_F32 = 2048          # lib.rs:67
MAX_ENGINES = 64
_memory_buffer = np.zeros((MAX_ENGINES + 1) * REGION_F32 * 4, np.uint8)   # region 0 unused: a valid pointer is never 0
_memory = types.SimpleNamespace(buffer=_memory_buffer)
_next_region = 1


class PhysicsEngine(_engine.PhysicsEngine):
    """addon/ts/index.ts `class PhysicsEngine extends NativePhysicsEngine`."""

    def __init__(self, mass, spin):
        global _next_region
        super().__init__(mass, spin)
        if _next_region > MAX_ENGINES:
            raise RuntimeError("blackhole-physics: more than 64 live PhysicsEngine instances")
        self._byte_offset = _next_region * REGION_F32 * 4
        _next_region += 1
        region = _memory_buffer[self._byte_offset:self._byte_offset + REGION_F32 * 4].view(np.float32)
        self.attach_sab(region)

    def get_sab_ptr(self):  # lib.rs:116-118: a byte offset into memory.buffer
        return self._byte_offset


def init_hooks():
    return None


def init():
    """default export: `await wasmModuleWrap.default()` -> { memory }"""
    return types.SimpleNamespace(memory=_memory)
