// Drop-in for src/rendering/webgl/renderer.ts (WebGLRenderer): same public surface (init / resize / render(params,
// mouse) / cleanup / error / onMetricsUpdate); render() runs the production fragment shader as one CUDA kernel through
// the addon and blits the tone-mapped frame. Callers (components/canvas/WebGLCanvas.tsx) are unchanged.
// RUNTIME: needs a DOM and Node-API together (Electron / NW.js renderer with nodeIntegration), like kerr_b200_renderer.ts.
import { KerrRenderer } from "./index";
import type { SimulationParams } from "@/types/simulation";
import { DEFAULT_FEATURES, getMaxRaySteps } from "@/types/features";
import { SIMULATION_CONFIG } from "@/configs/simulation.config";
import { physicsBridge } from "@/engine/physics-bridge";

const F = { LENSING: 1, DISK: 2, JETS: 4, STARS: 8, PHOTON_GLOW: 16, DOPPLER: 32, REDSHIFT: 64, LINEAR_OUTPUT: 128, QUALITY_LOW: 256 };
const WORDS = 155;                                               // sizeof(GvtGlslUniforms) / 4

export class WebGLB200Renderer {
  public error: string | null = null;                            // renderer.ts:38
  public onMetricsUpdate?: (m: unknown) => void;                 // renderer.ts:39
  private r: InstanceType<typeof KerrRenderer> | null = null;
  private ctx: CanvasRenderingContext2D | null = null;
  private out = new ArrayBuffer(0);
  private width = 0;
  private height = 0;
  private time = 0;
  private lastMouse = { x: 0, y: 0 };

  init(canvas: HTMLCanvasElement): boolean {                      // renderer.ts:58
    try {
      this.r = new KerrRenderer(0);
      const noise = new Uint8Array(256 * 256 * 4), blue = new Uint8Array(256 * 256 * 4);
      for (let i = 0; i < noise.length; i++) { noise[i] = Math.floor(Math.random() * 255); blue[i] = Math.floor(Math.random() * 255); }
      this.r.setNoiseTextures(noise, blue);                       // utils/webgl-utils.ts:259-303
      this.ctx = canvas.getContext("2d");
      return this.ctx !== null;
    } catch (e) { this.error = String(e); return false; }
  }
  resize(width: number, height: number): void {                   // renderer.ts:163
    if (this.width === width && this.height === height) return;
    this.width = width; this.height = height;
    this.out = new ArrayBuffer(width * height * 4);
    this.r?.resize(width, height);
  }
  render(params: SimulationParams, mouse: { x: number; y: number }): void {   // renderer.ts:173
    if (!this.r) return;
    const moving = Math.abs(mouse.x - this.lastMouse.x) > 1e-4 || Math.abs(mouse.y - this.lastMouse.y) > 1e-4;
    this.lastMouse = { ...mouse };
    if (!params.paused) this.time += 0.01;
    const f = params.features || DEFAULT_FEATURES;
    const buf = new ArrayBuffer(WORDS * 4), u = new Float32Array(buf), w = new Uint32Array(buf), iw = new Int32Array(buf);
    let bits = 0;                                                 // shaders/manager.ts:55-82
    if (f.gravitationalLensing) bits |= F.LENSING;
    if (f.accretionDisk) bits |= F.DISK;
    if (f.dopplerBeaming) bits |= F.DOPPLER;
    if (f.backgroundStars) bits |= F.STARS;
    if (f.photonSphereGlow) bits |= F.PHOTON_GLOW;
    if (f.relativisticJets && f.accretionDisk) bits |= F.JETS;
    if (f.gravitationalRedshift) bits |= F.REDSHIFT;
    if (f.rayTracingQuality === "low" || f.rayTracingQuality === "off") bits |= F.QUALITY_LOW;
    w[0] = WORDS * 4; w[1] = bits;
    u[2] = this.width; u[3] = this.height; u[4] = this.time; u[5] = params.mass; u[6] = params.spin * params.mass;   // renderer.ts:319-326
    u[7] = params.diskDensity ?? SIMULATION_CONFIG.diskDensity.default;
    u[8] = (params.diskTemp ?? SIMULATION_CONFIG.diskTemp.default) * Math.pow(params.mass, -0.25);                     // :351-354
    u[9] = mouse.x; u[10] = mouse.y; u[11] = params.zoom * 2.0; u[12] = params.lensing ?? 1.0;
    u[13] = params.diskSize ?? SIMULATION_CONFIG.diskSize.default;
    u[14] = params.diskScaleHeight ?? SIMULATION_CONFIG.diskScaleHeight.default;
    iw[15] = getMaxRaySteps(f.rayTracingQuality);
    u[16] = 0; u[17] = f.gravitationalRedshift ? 1 : 0; u[18] = f.kerrShadow ? 1 : 0;
    // u[20..22] camPos = 0 selects the mouse/zoom camera (renderer.ts:314-315); camQuat = identity
    u[26] = 1;
    let count = 0;                                                // shadow curve from the SAB (renderer.ts:277-298)
    if (physicsBridge && physicsBridge.isReady()) {
      const t = physicsBridge.tick(0.016);
      if (t && t.physics[15] > 0) { count = t.physics[15]; for (let i = 0; i < 128; i++) u[27 + i] = t.physics[16 + i]; }
    }
    if (count <= 0) {
      const b = 3 * Math.sqrt(3) * params.mass; count = 64;
      for (let i = 0; i < 64; i++) { const phi = (i / 64) * Math.PI * 2; u[27 + 2 * i] = Math.cos(phi) * b; u[28 + 2 * i] = Math.sin(phi) * b; }
    }
    u[19] = count;
    let stats;
    if (f.bloom) {                                                // renderer.ts:366-399: linear HDR scene -> bloom -> final pass
      w[1] = bits | F.LINEAR_OUTPUT;
      stats = this.r.renderFragment(u, { format: 0, cameraMoving: moving ? 1 : 0 }, new ArrayBuffer(0));
      this.r.bloom({ enabled: 1, format: 4 }, this.out);          // bloom.ts:32-39 defaults
    } else {
      stats = this.r.renderFragment(u, { format: 4 /* RGBA8_UNORM: the shader already applied ACES + gamma */, cameraMoving: moving ? 1 : 0 }, this.out);
    }
    if (this.onMetricsUpdate) this.onMetricsUpdate(stats);
    if (this.ctx) {                                               // rows arrive bottom-up (gl_FragCoord): flip while blitting
      const img = this.ctx.createImageData(this.width, this.height), src = new Uint8Array(this.out), rb = this.width * 4;
      for (let y = 0; y < this.height; y++) img.data.set(src.subarray((this.height - 1 - y) * rb, (this.height - y) * rb), y * rb);
      this.ctx.putImageData(img, 0, 0);
    }
  }
  cleanup(): void { this.r = null; }                              // renderer.ts:471
}
