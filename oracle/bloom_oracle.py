"""CPU ORACLE (test infrastructure) for the WebGL2 bloom + final pass — numpy float32 restatement of
src/shaders/postprocess/bloom.glsl.ts (brightPassShader :36-58, blurShader :64-90, combineShader :96-128) driven as
src/rendering/bloom.ts:446-590 (applyBloomToTexture) drives them: half-res bright pass, quarter-res RGBA16F ping-pong
blur with `blurPasses` (horizontal, vertical) iterations, full-res combine + ACES + gamma. Textures are RGBA16F with
LINEAR / CLAMP_TO_EDGE sampling (bloom.ts:206-233). PARITY STATUS: "parity unpinned" (no reference fixture)."""
import numpy as np

F = np.float32
WEIGHTS = [F(0.227027), F(0.1945946), F(0.1216216), F(0.054054), F(0.016216)]   # bloom.glsl.ts:75


def sample_linear(img, u, v):
    """texture(img, (u, v)) with LINEAR filtering and CLAMP_TO_EDGE; img (h, w, 4) float32, u/v float32 arrays."""
    h, w = img.shape[:2]
    x = (u * F(w) - F(0.5)).astype(F); y = (v * F(h) - F(0.5)).astype(F)
    xf = np.floor(x); yf = np.floor(y)
    fx = (x - xf).astype(F)[..., None]; fy = (y - yf).astype(F)[..., None]
    xi = xf.astype(np.int64); yi = yf.astype(np.int64)
    x0 = np.clip(xi, 0, w - 1); x1 = np.clip(xi + 1, 0, w - 1)
    y0 = np.clip(yi, 0, h - 1); y1 = np.clip(yi + 1, 0, h - 1)
    a, b, c, d = img[y0, x0], img[y0, x1], img[y1, x0], img[y1, x1]
    top = (a + (b - a) * fx).astype(F); bot = (c + (d - c) * fx).astype(F)
    return (top + (bot - top) * fy).astype(F)


def _uv(w, h):
    u = ((np.arange(w, dtype=F) + F(0.5)) / F(w)).astype(F)[None, :].repeat(h, 0)
    v = ((np.arange(h, dtype=F) + F(0.5)) / F(h)).astype(F)[:, None].repeat(w, 1)
    return u, v


def _f16(img):
    return img.astype(np.float16).astype(F)   # an RGBA16F render target


def bright_pass(scene, threshold):
    H, W = scene.shape[:2]
    hw, hh = max(1, W // 2), max(1, H // 2)
    u, v = _uv(hw, hh)
    c = sample_linear(scene, u, v)
    lum = (c[..., 0] * F(0.299) + c[..., 1] * F(0.587) + c[..., 2] * F(0.114)).astype(F)
    return _f16(np.where((lum > F(threshold))[..., None], c, F(0)).astype(F))


def blur(src, bw, bh, dx, dy):
    u, v = _uv(bw, bh)
    tx, ty = F(1.0) / F(bw), F(1.0) / F(bh)
    res = (sample_linear(src, u, v)[..., :3] * WEIGHTS[0]).astype(F)
    for k in range(1, 5):
        ox, oy = F(dx) * tx * F(k), F(dy) * ty * F(k)
        res = (res + sample_linear(src, (u + ox).astype(F), (v + oy).astype(F))[..., :3] * WEIGHTS[k]).astype(F)
        res = (res + sample_linear(src, (u - ox).astype(F), (v - oy).astype(F))[..., :3] * WEIGHTS[k]).astype(F)
    out = np.ones((bh, bw, 4), F)
    out[..., :3] = res
    return _f16(out)


def aces_gamma(x):
    x = x.astype(F)
    t = np.clip((x * (F(2.51) * x + F(0.03))) / (x * (F(2.43) * x + F(0.59)) + F(0.14)), F(0), F(1)).astype(F)
    return np.power(t, F(0.4545)).astype(F)


def apply_bloom(scene, enabled=True, intensity=0.5, threshold=0.8, blur_passes=2):
    """scene: (H, W, 4) float32 linear HDR -> (H, W, 4) float32 display-referred (alpha 1)."""
    scene = np.asarray(scene, F)
    H, W = scene.shape[:2]
    res = scene[..., :3].copy()
    if enabled:
        src = bright_pass(scene, threshold)
        bw, bh = max(1, W // 4), max(1, H // 4)
        for _ in range(blur_passes):
            b1 = blur(src, bw, bh, 1.0, 0.0)
            src = blur(b1, bw, bh, 0.0, 1.0)
        u, v = _uv(W, H)
        res = (res + sample_linear(src, u, v)[..., :3] * F(intensity)).astype(F)
    out = np.ones((H, W, 4), F)
    out[..., :3] = aces_gamma(res)
    return out
