"""CPU ORACLE (test infrastructure) for the TAA resolve pass — numpy float32 restatement of
src/shaders/postprocess/ataa.wgsl.ts:11-83. PARITY STATUS: "parity unpinned" (the reference has no test or
fixture for this pass; it cannot run here — WGSL needs a browser)."""
import numpy as np

F = np.float32


def rgb_to_ycocg(rgb):  # ataa.wgsl.ts:11-16
    r, g, b = rgb[..., 0], rgb[..., 1], rgb[..., 2]
    y = F(0.25) * r + F(0.5) * g + F(0.25) * b
    co = F(0.5) * r - F(0.5) * b
    cg = F(-0.25) * r + F(0.5) * g - F(0.25) * b
    return np.stack([y, co, cg], -1).astype(F)


def ycocg_to_rgb(c):  # ataa.wgsl.ts:18-26
    y, co, cg = c[..., 0], c[..., 1], c[..., 2]
    return np.stack([y + co - cg, y + cg, y - co - cg], -1).astype(F)


def mat_vec(m16, v):  # column-major mat4 (gl-matrix / WGSL) times vec4 field v[..., 4]
    m = np.asarray(m16, F).reshape(4, 4)  # m[col][row]
    return sum(m[c][None, None, :] * v[..., c:c + 1] for c in range(4)).astype(F)


def bilinear_clamp(img, u, v):  # textureSampleLevel + linear sampler, clamp-to-edge (ataa.wgsl.ts:72)
    H, W = img.shape[:2]
    x = np.clip(u * F(W) - F(0.5), F(0), F(W - 1)).astype(F)
    y = np.clip(v * F(H) - F(0.5), F(0), F(H - 1)).astype(F)
    x0 = np.floor(x).astype(np.int64); y0 = np.floor(y).astype(np.int64)
    x1 = np.minimum(x0 + 1, W - 1); y1 = np.minimum(y0 + 1, H - 1)
    fx = (x - x0.astype(F))[..., None]; fy = (y - y0.astype(F))[..., None]
    top = img[y0, x0] + (img[y0, x1] - img[y0, x0]) * fx
    bot = img[y1, x0] + (img[y1, x1] - img[y1, x0]) * fx
    return (top + (bot - top) * fy).astype(F)


def taa_resolve(cam88, cur, hist):
    """cur, hist: (H, W, 4) float32 -> (H, W, 4) float32."""
    cam88 = np.asarray(cam88, F)
    cur = np.asarray(cur, F); hist = np.asarray(hist, F)
    H, W = cur.shape[:2]
    ycc = rgb_to_ycocg(cur[..., :3])
    m1 = np.zeros_like(ycc); m2 = np.zeros_like(ycc)
    ys, xs = np.arange(H), np.arange(W)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            s = ycc[np.clip(ys + dy, 0, H - 1)][:, np.clip(xs + dx, 0, W - 1)]   # ataa.wgsl.ts:43
            m1 += s; m2 += s * s
    mean = m1 / F(9.0)
    std = np.sqrt(np.maximum(m2 / F(9.0) - mean * mean, F(0)))
    lo, hi = mean - F(2.0) * std, mean + F(2.0) * std
    u = ((xs.astype(F) + F(0.5)) / F(W))[None, :].repeat(H, 0)
    v = ((ys.astype(F) + F(0.5)) / F(H))[:, None].repeat(W, 1)
    clip = np.stack([u * F(2) - F(1), -(v * F(2) - F(1)), np.ones_like(u), np.ones_like(u)], -1).astype(F)
    vt = mat_vec(cam88[48:64], clip)                                   # inv_proj
    vd = vt[..., :3] / vt[..., 3:4]
    vd = vd / np.sqrt((vd * vd).sum(-1, keepdims=True))
    wd = mat_vec(cam88[32:48], np.concatenate([vd, np.zeros_like(vd[..., :1])], -1))[..., :3]   # inv_view, w = 0
    wp = cam88[80:83][None, None, :] + wd * F(12.0)                    # reprojectDepth = 12 (ataa.wgsl.ts:64)
    pc = mat_vec(cam88[64:80], np.concatenate([wp, np.ones_like(wp[..., :1])], -1))
    pu = (pc[..., 0] / pc[..., 3]) * F(0.5) + F(0.5)
    pv = (pc[..., 1] / pc[..., 3]) * F(-0.5) + F(0.5)
    h = rgb_to_ycocg(bilinear_clamp(hist[..., :3], pu.astype(F), pv.astype(F)))
    h = np.minimum(np.maximum(h, lo), hi)
    res = ycc * F(1.0 - 0.92) + h * F(0.92)                            # mix(center, history, 0.92)
    out = np.ones((H, W, 4), F)
    out[..., :3] = ycocg_to_rgb(res)
    return out


def taa_resolve_webgl(cur, hist, blend=0.75, camera_moving=False):
    """numpy f32 restatement of src/shaders/postprocess/reprojection.glsl.ts:70-115 at render scale 1 (texel-centre
    fetches with CLAMP_TO_EDGE): mu +- 1.5 sigma clip of the same-texel history, variance-guided blend weight."""
    cur = np.asarray(cur, F); hist = np.asarray(hist, F)
    H, W = cur.shape[:2]
    ycc = rgb_to_ycocg(cur[..., :3])
    m1 = np.zeros_like(ycc); m2 = np.zeros_like(ycc)
    ys, xs = np.arange(H), np.arange(W)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            s = ycc[np.clip(ys + dy, 0, H - 1)][:, np.clip(xs + dx, 0, W - 1)]
            m1 += s; m2 += s * s
    mean = m1 / F(9.0)
    std = np.sqrt(np.maximum(m2 / F(9.0) - mean * mean, F(0)))
    lo, hi = mean - F(1.5) * std, mean + F(1.5) * std
    h = np.minimum(np.maximum(rgb_to_ycocg(hist[..., :3]), lo), hi)
    wv = F(1.0) - np.clip(std[..., 0] * F(4.0), F(0), F(0.55))
    alpha = (np.zeros_like(wv) if camera_moving else F(blend) * wv)[..., None]
    res = ycc * (F(1.0) - alpha) + h * alpha
    out = np.ones((H, W, 4), F)
    out[..., :3] = ycocg_to_rgb(res)
    return out
