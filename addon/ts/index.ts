// Drop-in for the `blackhole-physics` module alias (tsconfig.json:23, vitest.config.ts:13-16): same shape as the
// wasm-bindgen `--target web` output that src/workers/physics.worker.ts:60-64 and src/engine/physics-bridge.ts:86-89
// import — a default async init() resolving to an object with `.memory`, and a named `PhysicsEngine` class — and the
// same MEMORY CONTRACT, so that neither file needs an edit:
//
//   * init() hands out ONE SharedArrayBuffer as `memory.buffer` (the stand-in for the wasm linear memory);
//   * every `new PhysicsEngine(m, a)` is attached to its own 2048-f32 region of that buffer (attach_sab on a Float32Array
//     view of the region), so the native engine writes its SAB block there on every tick_sab();
//   * get_sab_ptr() returns that region's BYTE OFFSET inside memory.buffer — exactly what physics.worker.ts:153-163
//     (`wasmF32.subarray(ptr / 4 + OFFSETS.CAMERA, ...)`) and physics-bridge.ts:107 (`new Float32Array(memory.buffer,
//     ptr + OFFSETS.CONTROL * 4, 16)`) index with. The bridge's fallback path writes its inputs into that CONTROL block,
//     which is the block the engine reads (lib.rs:317-328).
//
// RUNTIME: a native addon needs a Node-API host. The reference runs the engine inside a browser Web Worker; the hosts
// this module targets are an Electron / NW.js renderer (BrowserWindow with nodeIntegration, workers created with
// nodeIntegrationInWorker: true — both the page and the worker can then load it) or plain Node (vitest, SSR, a render
// server). A stock browser tab cannot load it; see INTEGRATION.md 4 for the server-side arrangement.
import { createRequire } from "module";
const addon = createRequire(import.meta.url)("../build/Release/gravitas_b200.node");

const REGION_F32 = 2048;                         // lib.rs:67: the engine-owned buffer is 2048 f32
const MAX_ENGINES = 64;
// region 0 starts at byte 8192 so that a valid get_sab_ptr() is never 0 (a null pointer in the wasm build either)
const memoryBuffer = new SharedArrayBuffer((MAX_ENGINES + 1) * REGION_F32 * 4);
const memory = { buffer: memoryBuffer };
let nextRegion = 1;

type NativeEngine = {
  update_params(mass: number, spin: number): void;
  tick_sab(dtOverride: number): void;
  attach_sab(view: Float32Array): void;                          // lib.rs:74 — the engine then writes that memory in place
  get_sab_ptr(): number;                                         // lib.rs:116 — byte offset of the attached view in its buffer
  get_sab(): Float32Array;                                       // snapshot of the engine-owned 2048-f32 buffer (debugging aid)
  get_sab_layout(): Uint32Array;                                 // [0, 64, 128, 256, 2048]
  set_camera_state(px: number, py: number, pz: number, lx: number, ly: number, lz: number): void;
  set_auto_spin(enabled: boolean): void;
  compute_horizon(): number; compute_isco(): number; compute_photon_sphere(): number;
  compute_dilation(r: number): number; compute_g_factor(r: number, lambda: number): number;
  compute_shadow_curve(thetaObs: number, nPoints: number): Float32Array; compute_shadow_shift(thetaObs: number): Float32Array;
  compute_shadow_radius(): number; compute_disk_flux(r: number): number;
  compute_kretschner(r: number, theta: number): number; compute_light_cone_tilt(r: number, theta: number): number;
  compute_frame_drag_omega(r: number, theta: number): number; compute_flamm_height(r: number): number;
  compute_proper_distance(r1: number, r2: number, nSteps: number): number;
  generate_curvature_field(rMin: number, rMax: number, nRadial: number, nPolar: number): Float32Array;
  generate_tilt_field(rMin: number, rMax: number, nRadial: number, nPolar: number): Float32Array;
  generate_frame_drag_field(rMin: number, rMax: number, nRadial: number, nPolar: number): Float32Array;
  generate_embedding_mesh(rMin: number, rMax: number, nRadial: number, nAngular: number): Float32Array;
  generate_ergosphere_mesh(nPolar: number, nAzimuthal: number): Float32Array;
  generate_disk_lut(): Float32Array; get_disk_lut_ptr(): Float32Array; generate_spectrum_lut(w: number, h: number, maxTemp: number): Float32Array;
  integrate_ray_relativistic(state: number[], steps: number, tol: number, useKerrSchild: boolean): Float64Array;
};
const NativePhysicsEngine: new (mass: number, spin: number) => NativeEngine = addon.PhysicsEngine;

/** Same constructor and methods as the wasm-bindgen class (gravitas-wasm/src/lib.rs:42-465). */
export class PhysicsEngine extends NativePhysicsEngine {
  constructor(mass: number, spin: number) {
    super(mass, spin);
    if (nextRegion > MAX_ENGINES) throw new Error("blackhole-physics: more than 64 live PhysicsEngine instances");
    const region = new Float32Array(memoryBuffer, nextRegion * REGION_F32 * 4, REGION_F32);
    nextRegion += 1;
    this.attach_sab(region);                                      // get_sab_ptr() === region.byteOffset from here on
  }
  free(): void {}                                                 // wasm-bindgen's explicit destructor; GC finalises the native engine
}

export const KerrRenderer = addon.KerrRenderer;
// importExternalFd(fd, bytes, device, dedicated): an opaque frame target in device memory (INTEGRATION.md 6)
export const importExternalFd: (fd: number, bytes: number, device?: number, dedicated?: boolean) => unknown = addon.importExternalFd;
export function init_hooks(): void {}                             // lib.rs:30-33: the wasm panic hook has no native counterpart

export default async function init(): Promise<{ memory: { buffer: SharedArrayBuffer } }> {
  return { memory };                                              // the same object every time, like the wasm instance's exports
}
