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

// PRECISE = true is the validation build: every operation an explicitly rounded IEEE f32 operation (no FMA contraction,
// IEEE division, libm powf) in the order the GLSL text evaluates them, so that the pass can be tested to 1e-6 against
// the numpy restatement of bloom.glsl.ts and the production build's FMA / MUFU shortcuts become a measured difference.
template <bool PRECISE> __device__ __forceinline__ float mul_(float a, float b) { return PRECISE ? __fmul_rn(a, b) : a * b; }
template <bool PRECISE> __device__ __forceinline__ float add_(float a, float b) { return PRECISE ? __fadd_rn(a, b) : a + b; }
template <bool PRECISE> __device__ __forceinline__ float sub_(float a, float b) { return PRECISE ? __fadd_rn(a, -b) : a - b; }
template <bool PRECISE> __device__ __forceinline__ float div_(float a, float b) { return PRECISE ? __fdiv_rn(a, b) : a / b; }
// a + (b - a) t
template <bool PRECISE> __device__ __forceinline__ float lerp_(float a, float b, float t) { return add_<PRECISE>(a, mul_<PRECISE>(sub_<PRECISE>(b, a), t)); }

// texture(sampler, uv) with LINEAR min/mag filter and CLAMP_TO_EDGE
template <bool PRECISE, class T> __device__ __forceinline__ float4 sample_linear(const T& t, float u, float v) {
    const float x = sub_<PRECISE>(mul_<PRECISE>(u, (float)t.w), 0.5f), y = sub_<PRECISE>(mul_<PRECISE>(v, (float)t.h), 0.5f);
    const float xf = floorf(x), yf = floorf(y);
    const float fx = sub_<PRECISE>(x, xf), fy = sub_<PRECISE>(y, yf);
    const int xi = (int)xf, yi = (int)yf;
    const int x0 = min(max(xi, 0), t.w - 1), x1 = min(max(xi + 1, 0), t.w - 1);
    const int y0 = min(max(yi, 0), t.h - 1), y1 = min(max(yi + 1, 0), t.h - 1);
    const float4 a = fetch(t, x0, y0), b = fetch(t, x1, y0), c = fetch(t, x0, y1), d = fetch(t, x1, y1);
    float4 o;
    o.x = lerp_<PRECISE>(lerp_<PRECISE>(a.x, b.x, fx), lerp_<PRECISE>(c.x, d.x, fx), fy);
    o.y = lerp_<PRECISE>(lerp_<PRECISE>(a.y, b.y, fx), lerp_<PRECISE>(c.y, d.y, fx), fy);
    o.z = lerp_<PRECISE>(lerp_<PRECISE>(a.z, b.z, fx), lerp_<PRECISE>(c.z, d.z, fx), fy);
    o.w = lerp_<PRECISE>(lerp_<PRECISE>(a.w, b.w, fx), lerp_<PRECISE>(c.w, d.w, fx), fy);
    return o;
}
__device__ __forceinline__ uint2 pack_half4(float4 v) {
    const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&lo); o.y = *reinterpret_cast<const uint32_t*>(&hi);
    return o;
}
// texel-centre coordinate (i + 0.5) / n
template <bool PRECISE> __device__ __forceinline__ float centre_(int i, int n) { return div_<PRECISE>(add_<PRECISE>((float)i, 0.5f), (float)n); }

// ST = Tex32 (RGBA32F frame chain) or Tex16 (RGBA16F frame chain: the reference's own scene texture format)
template <bool PRECISE, class ST>
__global__ void k_bloom_bright(ST scene, uint2* __restrict__ dst, int dw, int dh, float threshold) {
    const int n = dw * dh;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int x = i % dw, y = i / dw;
        const float4 c = sample_linear<PRECISE>(scene, centre_<PRECISE>(x, dw), centre_<PRECISE>(y, dh));
        const float lum = add_<PRECISE>(add_<PRECISE>(mul_<PRECISE>(c.x, 0.299f), mul_<PRECISE>(c.y, 0.587f)), mul_<PRECISE>(c.z, 0.114f));
        dst[i] = pack_half4(lum > threshold ? c : make_float4(0.f, 0.f, 0.f, 0.f));
    }
}

template <bool PRECISE>
__global__ void k_bloom_blur(Tex16 src, uint2* __restrict__ dst, int dw, int dh, float dirx, float diry) {
    const float wgt[5] = {0.227027f, 0.1945946f, 0.1216216f, 0.054054f, 0.016216f};
    const int n = dw * dh;
    const float tx = div_<PRECISE>(1.0f, (float)dw), ty = div_<PRECISE>(1.0f, (float)dh);   // texelSize = 1 / u_resolution (the blur target's size)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int x = i % dw, y = i / dw;
        const float u = centre_<PRECISE>(x, dw), v = centre_<PRECISE>(y, dh);
        float4 c = sample_linear<PRECISE>(src, u, v);
        float r = mul_<PRECISE>(c.x, wgt[0]), g = mul_<PRECISE>(c.y, wgt[0]), b = mul_<PRECISE>(c.z, wgt[0]);
#pragma unroll
        for (int k = 1; k < 5; k++) {
            const float ox = mul_<PRECISE>(mul_<PRECISE>(dirx, tx), (float)k), oy = mul_<PRECISE>(mul_<PRECISE>(diry, ty), (float)k);
            c = sample_linear<PRECISE>(src, add_<PRECISE>(u, ox), add_<PRECISE>(v, oy));
            r = add_<PRECISE>(r, mul_<PRECISE>(c.x, wgt[k])); g = add_<PRECISE>(g, mul_<PRECISE>(c.y, wgt[k])); b = add_<PRECISE>(b, mul_<PRECISE>(c.z, wgt[k]));
            c = sample_linear<PRECISE>(src, sub_<PRECISE>(u, ox), sub_<PRECISE>(v, oy));
            r = add_<PRECISE>(r, mul_<PRECISE>(c.x, wgt[k])); g = add_<PRECISE>(g, mul_<PRECISE>(c.y, wgt[k])); b = add_<PRECISE>(b, mul_<PRECISE>(c.z, wgt[k]));
        }
        dst[i] = pack_half4(make_float4(r, g, b, 1.0f));
    }
}

// bloom.glsl.ts:106-124. The combine pass is one read + one write of the frame; IEEE division and libm powf (three of
// each per pixel, with their slow-path branches) made it issue-bound at 2.5 TB/s, so the production pass uses MUFU
// rcp / lg2 / ex2 like the GLSL it restates (~1e-6 on a display-referred value that is quantised to 8 bits next);
// the PRECISE build keeps IEEE division and powf.
template <bool PRECISE> __device__ __forceinline__ float aces_gamma_f(float x) {
    if (PRECISE) {
        const float num = __fmul_rn(x, __fadd_rn(__fmul_rn(2.51f, x), 0.03f));
        const float den = __fadd_rn(__fmul_rn(x, __fadd_rn(__fmul_rn(2.43f, x), 0.59f)), 0.14f);
        return powf(fminf(fmaxf(__fdiv_rn(num, den), 0.0f), 1.0f), 0.4545f);
    }
    const float t = fminf(fmaxf(__fdividef(x * (2.51f * x + 0.03f), x * (2.43f * x + 0.59f) + 0.14f), 0.0f), 1.0f);
    return t > 0.0f ? exp2f(0.4545f * __log2f(t)) : 0.0f;
}

template <bool PRECISE, class ST>
__global__ void k_bloom_combine(ST scene, Tex16 bloom, float4* __restrict__ dst, float intensity, int use_bloom) {
    // one thread per pixel, rows on blockIdx.y (no 64-bit division in the index arithmetic)
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    for (int y = blockIdx.y; y < scene.h && x < scene.w; y += gridDim.y) {
        const size_t i = (size_t)y * scene.w + x;
        const float4 s = fetch(scene, x, y);                      // the scene is sampled at its own texel centres
        float r = s.x, g = s.y, b = s.z;
        if (use_bloom) {
            const float4 bl = sample_linear<PRECISE>(bloom, centre_<PRECISE>(x, scene.w), centre_<PRECISE>(y, scene.h));
            r = add_<PRECISE>(r, mul_<PRECISE>(bl.x, intensity)); g = add_<PRECISE>(g, mul_<PRECISE>(bl.y, intensity));
            b = add_<PRECISE>(b, mul_<PRECISE>(bl.z, intensity));
        }
        dst[i] = make_float4(aces_gamma_f<PRECISE>(r), aces_gamma_f<PRECISE>(g), aces_gamma_f<PRECISE>(b), 1.0f);
    }
}

template <bool PRECISE, class ST>
cudaError_t launch_bloom_t(const ST scene, int W, int H, uint2* half_tex, uint2* q1, uint2* q2, float4* display,
                           float threshold, float intensity, int blur_passes, int enabled, int sm_count, cudaStream_t stream,
                           int* launches) {
    const int hw = max(1, W / 2), hh = max(1, H / 2), bw = max(1, W / 4), bh = max(1, H / 4);
    const int grid = sm_count * 8;
    Tex16 result{q2, bw, bh};
    if (enabled) {
        k_bloom_bright<PRECISE, ST><<<grid, 256, 0, stream>>>(scene, half_tex, hw, hh, threshold);
        (*launches)++;
        Tex16 src{half_tex, hw, hh};
        for (int i = 0; i < blur_passes; i++) {
            k_bloom_blur<PRECISE><<<grid, 256, 0, stream>>>(src, q1, bw, bh, 1.0f, 0.0f);
            k_bloom_blur<PRECISE><<<grid, 256, 0, stream>>>(Tex16{q1, bw, bh}, q2, bw, bh, 0.0f, 1.0f);
            (*launches) += 2;
            src = Tex16{q2, bw, bh};
        }
        result = src;   // blurPasses = 0: the bright texture itself is combined (bloom.ts:517-546)
    }
    k_bloom_combine<PRECISE, ST><<<dim3((unsigned)((W + 255) / 256), (unsigned)min(H, 65535)), 256, 0, stream>>>(scene, result, display, intensity, enabled);
    (*launches)++;
    return cudaGetLastError();
}

}  // namespace

// scratch: half = (W/2)*(H/2) uint2, q1 / q2 = (W/4)*(H/4) uint2 each. `display` receives the final frame.
cudaError_t launch_bloom(const float4* frame, int W, int H, uint2* half_tex, uint2* q1, uint2* q2, float4* display,
                         float threshold, float intensity, int blur_passes, int enabled, int sm_count, cudaStream_t stream,
                         int* launches, bool precise, bool scene_f16) {
#define GVT_BLOOM_ARGS W, H, half_tex, q1, q2, display, threshold, intensity, blur_passes, enabled, sm_count, stream, launches
    if (scene_f16) {
        const Tex16 scene{reinterpret_cast<const uint2*>(frame), W, H};
        return precise ? launch_bloom_t<true, Tex16>(scene, GVT_BLOOM_ARGS) : launch_bloom_t<false, Tex16>(scene, GVT_BLOOM_ARGS);
    }
    const Tex32 scene{frame, W, H};
    return precise ? launch_bloom_t<true, Tex32>(scene, GVT_BLOOM_ARGS) : launch_bloom_t<false, Tex32>(scene, GVT_BLOOM_ARGS);
#undef GVT_BLOOM_ARGS
}

}  // namespace gvt
