// Drop-in for the `blackhole-physics` module alias (tsconfig.json:23, vitest.config.ts:13-16): same shape as the
// wasm-bindgen `--target web` output that src/workers/physics.worker.ts:60-64 and src/engine/physics-bridge.ts:86-89
// import — a default async init() resolving to an object with `.memory`, and a named `PhysicsEngine` class.
import { createRequire } from "module";
const addon = createRequire(import.meta.url)("../build/Release/gravitas_b200.node");

export const PhysicsEngine: new (mass: number, spin: number) => {
  update_params(mass: number, spin: number): void;
  tick_sab(dtOverride: number): void;
  attach_sab(sab: SharedArrayBuffer | ArrayBuffer): void;       // lib.rs:74 — the engine then writes the worker's SAB in place
  get_sab(): Float32Array;                                       // snapshot of the engine-owned 2048-f32 buffer (lib.rs:116)
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
} = addon.PhysicsEngine;

export const KerrRenderer = addon.KerrRenderer;
export function init_hooks(): void {}                           // lib.rs:30-33: the wasm panic hook has no native counterpart

// physics.worker.ts:61,68 and physics-bridge.ts:87-88 only use `.memory.buffer` to build Float32Array views over the
// engine's SAB block; with attach_sab() the engine writes the caller's SharedArrayBuffer directly.
export default async function init(): Promise<{ memory: { buffer: SharedArrayBuffer } }> {
  return { memory: { buffer: new SharedArrayBuffer(2 * 1024 * 1024) } };
}
