// gvt_bloom.cu — the WebGL2 pipeline's bloom + final pass (src/rendering/bloom.ts:446-590 applyBloomToTexture,
// src/shaders/postprocess/bloom.glsl.ts) on the finished (TAA-resolved, linear HDR) frame:
//   1. bright pass   full-res RGBA32F frame -> half-res RGBA16F: LINEAR sample at the half-res texel centre,
//                    keep the colour if luminance (0.299, 0.587, 0.114) > threshold, else 0   (bloom.glsl.ts:36-58)
//   2. blur passes   quarter-res RGBA16F ping-pong: 9-tap separable Gaussian, horizontal then vertical, `blurPasses`
//                    times; the first horizontal pass reads the half-res bright texture through the LINEAR sampler
//                    (bloom.glsl.ts:64-90; bloom.ts:508-546)
//   3. combine       full res: scene + bloom (LINEAR upsample) * intensity -> ACES -> pow(1/2.2)   (bloom.glsl.ts:96-128)
// Intermediate textures are RGBA16F with LINEAR / CLAMP_TO_EDGE sampling, as bloom.ts:206-233 creates them; bilinear
// weights are exact f32 (a GPU's texture unit quantises them to 8 bits). All three are HBM-trivial next to the march:
// the bright pass reads the frame once (133 MB at 4K), the combine reads it again and writes the display frame.
#include "gvt_internal.h"
#include <cuda_fp16.h>

namespace gvt {

namespace {

struct Tex16 { const uint2* p; int w, h; };   // RGBA16F
struct Tex32 { const float4* p; int w, h; };  // RGBA32F

__device__ __forceinline__ float4 fetch(const Tex16& t, int x, int y) {
    const uint2 v = __ldg(t.p + (size_t)y * t.w + x);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 fetch(const Tex32& t, int x, int y) { return __ldg(t.p + (size_t)y * t.w + x); }

// texture(sampler, uv) with LINEAR min/mag filter and CLAMP_TO_EDGE
template <class T> __device__ __forceinline__ float4 sample_linear(const T& t, float u, float v) {
    const float x = u * (float)t.w - 0.5f, y = v * (float)t.h - 0.5f;
    const float xf = floorf(x), yf = floorf(y);
    const float fx = x - xf, fy = y - yf;
    const int xi = (int)xf, yi = (int)yf;
    const int x0 = min(max(xi, 0), t.w - 1), x1 = min(max(xi + 1, 0), t.w - 1);
    const int y0 = min(max(yi, 0), t.h - 1), y1 = min(max(yi + 1, 0), t.h - 1);
    const float4 a = fetch(t, x0, y0), b = fetch(t, x1, y0), c = fetch(t, x0, y1), d = fetch(t, x1, y1);
    float4 o;
    { const float tp = a.x + (b.x - a.x) * fx, bt = c.x + (d.x - c.x) * fx; o.x = tp + (bt - tp) * fy; }
    { const float tp = a.y + (b.y - a.y) * fx, bt = c.y + (d.y - c.y) * fx; o.y = tp + (bt - tp) * fy; }
    { const float tp = a.z + (b.z - a.z) * fx, bt = c.z + (d.z - c.z) * fx; o.z = tp + (bt - tp) * fy; }
    { const float tp = a.w + (b.w - a.w) * fx, bt = c.w + (d.w - c.w) * fx; o.w = tp + (bt - tp) * fy; }
    return o;
}
__device__ __forceinline__ uint2 pack_half4(float4 v) {
    const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&lo); o.y = *reinterpret_cast<const uint32_t*>(&hi);
    return o;
}

__global__ void k_bloom_bright(Tex32 scene, uint2* __restrict__ dst, int dw, int dh, float threshold) {
    const int n = dw * dh;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int x = i % dw, y = i / dw;
        const float4 c = sample_linear(scene, ((float)x + 0.5f) / (float)dw, ((float)y + 0.5f) / (float)dh);
        const float lum = c.x * 0.299f + c.y * 0.587f + c.z * 0.114f;
        dst[i] = pack_half4(lum > threshold ? c : make_float4(0.f, 0.f, 0.f, 0.f));
    }
}

__global__ void k_bloom_blur(Tex16 src, uint2* __restrict__ dst, int dw, int dh, float dirx, float diry) {
    const float wgt[5] = {0.227027f, 0.1945946f, 0.1216216f, 0.054054f, 0.016216f};
    const int n = dw * dh;
    const float tx = 1.0f / (float)dw, ty = 1.0f / (float)dh;   // texelSize = 1 / u_resolution (the blur target's size)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int x = i % dw, y = i / dw;
        const float u = ((float)x + 0.5f) / (float)dw, v = ((float)y + 0.5f) / (float)dh;
        float4 c = sample_linear(src, u, v);
        float r = c.x * wgt[0], g = c.y * wgt[0], b = c.z * wgt[0];
#pragma unroll
        for (int k = 1; k < 5; k++) {
            const float ox = dirx * tx * (float)k, oy = diry * ty * (float)k;
            c = sample_linear(src, u + ox, v + oy);
            r += c.x * wgt[k]; g += c.y * wgt[k]; b += c.z * wgt[k];
            c = sample_linear(src, u - ox, v - oy);
            r += c.x * wgt[k]; g += c.y * wgt[k]; b += c.z * wgt[k];
        }
        dst[i] = pack_half4(make_float4(r, g, b, 1.0f));
    }
}

// bloom.glsl.ts:106-124. The combine pass is one read + one write of the frame; IEEE division and libm powf (three of
// each per pixel, with their slow-path branches) made it issue-bound at 2.5 TB/s, so the final pass uses MUFU
// rcp / lg2 / ex2 like the GLSL it restates (~1e-6 on a display-referred value that is quantised to 8 bits next).
__device__ __forceinline__ float aces_gamma_f(float x) {
    const float t = fminf(fmaxf(__fdividef(x * (2.51f * x + 0.03f), x * (2.43f * x + 0.59f) + 0.14f), 0.0f), 1.0f);
    return t > 0.0f ? exp2f(0.4545f * __log2f(t)) : 0.0f;
}

__global__ void k_bloom_combine(Tex32 scene, Tex16 bloom, float4* __restrict__ dst, float intensity, int use_bloom) {
    // one thread per pixel, rows on blockIdx.y (no 64-bit division in the index arithmetic)
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    for (int y = blockIdx.y; y < scene.h && x < scene.w; y += gridDim.y) {
        const size_t i = (size_t)y * scene.w + x;
        const float4 s = __ldg(scene.p + i);                      // the scene is sampled at its own texel centres
        float r = s.x, g = s.y, b = s.z;
        if (use_bloom) {
            const float4 bl = sample_linear(bloom, ((float)x + 0.5f) / (float)scene.w, ((float)y + 0.5f) / (float)scene.h);
            r = r + bl.x * intensity; g = g + bl.y * intensity; b = b + bl.z * intensity;
        }
        dst[i] = make_float4(aces_gamma_f(r), aces_gamma_f(g), aces_gamma_f(b), 1.0f);
    }
}

}  // namespace

// scratch: half = (W/2)*(H/2) uint2, q1 / q2 = (W/4)*(H/4) uint2 each. `display` receives the final frame.
cudaError_t launch_bloom(const float4* frame, int W, int H, uint2* half_tex, uint2* q1, uint2* q2, float4* display,
                         float threshold, float intensity, int blur_passes, int enabled, int sm_count, cudaStream_t stream,
                         int* launches) {
    const int hw = max(1, W / 2), hh = max(1, H / 2), bw = max(1, W / 4), bh = max(1, H / 4);
    const Tex32 scene{frame, W, H};
    const int grid = sm_count * 8;
    Tex16 result{q2, bw, bh};
    if (enabled) {
        k_bloom_bright<<<grid, 256, 0, stream>>>(scene, half_tex, hw, hh, threshold);
        (*launches)++;
        Tex16 src{half_tex, hw, hh};
        for (int i = 0; i < blur_passes; i++) {
            k_bloom_blur<<<grid, 256, 0, stream>>>(src, q1, bw, bh, 1.0f, 0.0f);
            k_bloom_blur<<<grid, 256, 0, stream>>>(Tex16{q1, bw, bh}, q2, bw, bh, 0.0f, 1.0f);
            (*launches) += 2;
            src = Tex16{q2, bw, bh};
        }
        result = src;   // blurPasses = 0: the bright texture itself is combined (bloom.ts:517-546)
    }
    k_bloom_combine<<<dim3((unsigned)((W + 255) / 256), (unsigned)min(H, 65535)), 256, 0, stream>>>(scene, result, display, intensity, enabled);
    (*launches)++;
    return cudaGetLastError();
}

}  // namespace gvt
