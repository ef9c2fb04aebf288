// Drop-in for src/rendering/webgpu/renderer.ts (WebGPURenderer): same public methods, render() fills a frame buffer
// through the addon and blits it. Callers (components/canvas/WebGPUCanvas.tsx:73-194) are unchanged.
// RUNTIME: this file needs BOTH a DOM (the canvas it blits to) and Node-API (the addon, via ./index): an Electron / NW.js
// renderer process with nodeIntegration. In a stock browser the frames would have to come from a render server instead
// (INTEGRATION.md 4); that transport is not part of this repository.
import { KerrRenderer, importExternalFd } from "./index";
import { CameraUniforms, PhysicsParams, writeCameraUniforms, writePhysicsParams } from "@/types/webgpu";

export class KerrB200Renderer {
  private r = new KerrRenderer(0);
  private out = new ArrayBuffer(0);
  private ctx: CanvasRenderingContext2D | null = null;
  private maxSteps = 150;                                       // compute.wgsl.ts:13 default
  private width = 0;
  private height = 0;
  private target: unknown = null;                              // imported external frame target, or null (host frame + blit)

  async init(canvas: HTMLCanvasElement): Promise<boolean> {     // renderer.ts:82
    this.ctx = canvas.getContext("2d");
    return this.ctx !== null;
  }
  async initPipelines(_format: string, maxSteps: number): Promise<void> {   // renderer.ts:186
    this.maxSteps = maxSteps;
    this.r.initLuts(1.0, 0.9, 256, 32, 1e7);
  }
  updateSettings(maxSteps: number): void { this.maxSteps = maxSteps; }      // renderer.ts:256
  getFormat(): string { return "rgba32float"; }                              // renderer.ts:265
  resize(width: number, height: number): void {                              // renderer.ts:269
    this.width = width; this.height = height;
    this.out = new ArrayBuffer(width * height * 16);
    this.r.resize(width, height);
  }
  /** Hand-off without host memory (INTEGRATION.md 6): `fd` is the POSIX fd of a linear RGBA32F buffer the presenter
   *  exported from its graphics API (VK_KHR_external_memory_fd / GL_EXT_memory_object_fd). Frames are then stored into
   *  it by the producing kernel and `render` skips the canvas blit -- the presenter samples the buffer itself. */
  setExternalTarget(fd: number, bytes: number, device = 0): void {
    this.target = importExternalFd(fd, bytes, device, false);
  }
  render(camera: CameraUniforms, physics: PhysicsParams): void {             // renderer.ts:280
    const [w, h] = physics.resolution;
    if (w !== this.width || h !== this.height) this.resize(w, h);
    const cam = new Float32Array(88); writeCameraUniforms(cam, camera);      // types/webgpu.ts:89-116
    const ph = new Float32Array(8); writePhysicsParams(ph, physics);         // types/webgpu.ts:67-87
    if (this.target) { this.r.renderFrame(cam, ph, { maxSteps: this.maxSteps, taa: 1, jitter: 1 }, this.target); return; }
    this.r.renderFrame(cam, ph, { maxSteps: this.maxSteps, taa: 1, jitter: 1 }, this.out);
    this.blit(new Float32Array(this.out), w, h);
  }
  private blit(hdr: Float32Array, w: number, h: number): void {              // Reinhard as renderer.ts:45-47
    if (!this.ctx) return;
    const img = this.ctx.createImageData(w, h);
    for (let i = 0; i < w * h; i++) {
      for (let c = 0; c < 3; c++) {
        const v = hdr[4 * i + c];
        img.data[4 * i + c] = Math.min(255, Math.round(255 * Math.pow(v / (1 + v), 1 / 2.2)));
      }
      img.data[4 * i + 3] = 255;
    }
    this.ctx.putImageData(img, 0, 0);
  }
}
