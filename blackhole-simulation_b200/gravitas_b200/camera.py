"""gl-matrix-compatible camera maths (f32, column-major) -> the 88-float CameraUniforms block.

Follows components/canvas/WebGPUCanvas.tsx:119-178 (orbit eye, lookAt, perspective(fov 60 deg, near 0.1, far 1000),
inverses, direction, viewProj) and types/webgpu.ts:89-116 (packing). gl-matrix works on Float32Array, i.e. every
stored value is rounded to f32; intermediate arithmetic there is f64 (JS numbers) — mirrored here."""
import math

import numpy as np


def _f32(a):
    return np.asarray(a, dtype=np.float64).astype(np.float32)


def look_at(eye, center, up):  # gl-matrix mat4.lookAt
    ex, ey, ez = [float(v) for v in _f32(eye)]
    cx, cy, cz = [float(v) for v in _f32(center)]
    ux, uy, uz = [float(v) for v in _f32(up)]
    z0, z1, z2 = ex - cx, ey - cy, ez - cz
    ln = 1.0 / math.hypot(z0, z1, z2)
    z0, z1, z2 = z0 * ln, z1 * ln, z2 * ln
    x0, x1, x2 = uy * z2 - uz * z1, uz * z0 - ux * z2, ux * z1 - uy * z0
    ln = math.hypot(x0, x1, x2)
    if ln == 0:
        x0 = x1 = x2 = 0.0
    else:
        ln = 1.0 / ln
        x0, x1, x2 = x0 * ln, x1 * ln, x2 * ln
    y0, y1, y2 = z1 * x2 - z2 * x1, z2 * x0 - z0 * x2, z0 * x1 - z1 * x0
    ln = math.hypot(y0, y1, y2)
    if ln == 0:
        y0 = y1 = y2 = 0.0
    else:
        ln = 1.0 / ln
        y0, y1, y2 = y0 * ln, y1 * ln, y2 * ln
    out = [x0, y0, z0, 0.0, x1, y1, z1, 0.0, x2, y2, z2, 0.0,
           -(x0 * ex + x1 * ey + x2 * ez), -(y0 * ex + y1 * ey + y2 * ez), -(z0 * ex + z1 * ey + z2 * ez), 1.0]
    return _f32(out)


def perspective(fovy, aspect, near, far):  # gl-matrix mat4.perspective (perspectiveNO)
    f = 1.0 / math.tan(fovy / 2.0)
    out = [0.0] * 16
    out[0] = f / aspect
    out[5] = f
    out[11] = -1.0
    nf = 1.0 / (near - far)
    out[10] = (far + near) * nf
    out[14] = 2.0 * far * near * nf
    return _f32(out)


def invert(m):  # gl-matrix mat4.invert (cofactor expansion, f64 intermediates, f32 result)
    a = [float(v) for v in np.asarray(m, dtype=np.float32)]
    a00, a01, a02, a03, a10, a11, a12, a13, a20, a21, a22, a23, a30, a31, a32, a33 = a
    b00 = a00 * a11 - a01 * a10
    b01 = a00 * a12 - a02 * a10
    b02 = a00 * a13 - a03 * a10
    b03 = a01 * a12 - a02 * a11
    b04 = a01 * a13 - a03 * a11
    b05 = a02 * a13 - a03 * a12
    b06 = a20 * a31 - a21 * a30
    b07 = a20 * a32 - a22 * a30
    b08 = a20 * a33 - a23 * a30
    b09 = a21 * a32 - a22 * a31
    b10 = a21 * a33 - a23 * a31
    b11 = a22 * a33 - a23 * a32
    det = b00 * b11 - b01 * b10 + b02 * b09 + b03 * b08 - b04 * b07 + b05 * b06
    if det == 0:
        raise ValueError("singular matrix")
    det = 1.0 / det
    out = [
        (a11 * b11 - a12 * b10 + a13 * b09) * det, (a02 * b10 - a01 * b11 - a03 * b09) * det,
        (a31 * b05 - a32 * b04 + a33 * b03) * det, (a22 * b04 - a21 * b05 - a23 * b03) * det,
        (a12 * b08 - a10 * b11 - a13 * b07) * det, (a00 * b11 - a02 * b08 + a03 * b07) * det,
        (a32 * b02 - a30 * b05 - a33 * b01) * det, (a20 * b05 - a22 * b02 + a23 * b01) * det,
        (a10 * b10 - a11 * b08 + a13 * b06) * det, (a01 * b08 - a00 * b10 - a03 * b06) * det,
        (a30 * b04 - a31 * b02 + a33 * b00) * det, (a21 * b02 - a20 * b04 - a23 * b00) * det,
        (a11 * b07 - a10 * b09 - a12 * b06) * det, (a00 * b09 - a01 * b07 + a02 * b06) * det,
        (a31 * b01 - a30 * b03 - a32 * b00) * det, (a20 * b03 - a21 * b01 + a22 * b00) * det,
    ]
    return _f32(out)


def multiply(a, b):  # gl-matrix mat4.multiply(out, a, b) = a * b (column-major)
    A = np.asarray(a, dtype=np.float64).reshape(4, 4).T
    B = np.asarray(b, dtype=np.float64).reshape(4, 4).T
    return _f32((A @ B).T.reshape(16))


def orbit_eye(r0, polar_deg, azimuth):
    """Y-up orbit position: polar angle from +Y (simulation.config.ts:107), azimuth about Y (useCamera.ts:144)."""
    th = math.radians(polar_deg)
    return (r0 * math.sin(th) * math.cos(azimuth), r0 * math.cos(th), r0 * math.sin(th) * math.sin(azimuth))


def camera_uniforms(eye, width, height, prev_view_proj=None, fov_deg=60.0, near=0.1, far=1000.0):
    """-> (float32[88] CameraUniforms block, float32[16] viewProj for the next frame's prev_view_proj)."""
    view = look_at(eye, (0.0, 0.0, 0.0), (0.0, 1.0, 0.0))
    proj = perspective(math.radians(fov_deg), width / height, near, far)
    inv_view, inv_proj = invert(view), invert(proj)
    view_proj = multiply(proj, view)
    e = _f32(eye)
    d = -np.asarray(e, dtype=np.float64)
    d = _f32(d / np.linalg.norm(d))
    out = np.zeros(88, np.float32)
    out[0:16], out[16:32], out[32:48], out[48:64] = view, proj, inv_view, inv_proj
    out[64:80] = view_proj if prev_view_proj is None else np.asarray(prev_view_proj, np.float32)
    out[80:83] = e
    out[84:87] = d
    return out, view_proj


def default_camera(width, height, r0=30.0, polar_deg=97.0, azimuth=math.pi, prev_view_proj=None):
    """SURVEY §8(d) common setup: r0 = 30 (simulation.config.ts:119), polar 97 deg (:107), azimuth pi (useCamera.ts:144)."""
    return camera_uniforms(orbit_eye(r0, polar_deg, azimuth), width, height, prev_view_proj)
