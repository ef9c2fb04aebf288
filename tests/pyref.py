"""A second, independently written restatement of the reference path in pure Python floats (IEEE f64, CPython's
libm) — TEST INFRASTRUCTURE. It exists to cross-check oracle/ (the C++ restatement) on the pieces that neither the
reference's own tests nor SURVEY §8c's table pin: the fixed-step steppers, camera -> state, the LUT generators and
samplers, the g-factor and the composite RGBA of a pixel. Written from the Rust / WGSL sources directly (file:line
cited), not from the C++. Slow by construction: used on a handful of rays and texels only."""
import math

# ---- metric/kerr.rs ---------------------------------------------------------------------------------------------


def horizon(m, spin):  # metric/mod.rs:75-84
    a = spin * m
    disc = m * m - a * a
    return m if disc < 0.0 else m + math.sqrt(disc)


def isco(m, spin, prograde=True):  # kerr.rs:100-123
    if abs(spin) < 1e-6:
        return m * 6.0
    a2 = spin * spin
    z1 = 1.0 + (1.0 - a2) ** (1.0 / 3.0) * ((1.0 + spin) ** (1.0 / 3.0) + (1.0 - spin) ** (1.0 / 3.0))
    z2 = math.sqrt(3.0 * a2 + z1 * z1)
    sign = -1.0 if prograde else 1.0
    disc = (3.0 - z1) * (3.0 + z1 + 2.0 * z2)
    root = 0.0 if disc < 0.0 else math.sqrt(disc)
    return m * (3.0 + z2 + sign * root)


def contravariant_ks(m, a, r, theta):  # kerr.rs:412-440 -> dict of the non-zero entries by flat index
    r2, a2 = r * r, a * a
    sin2 = max(math.sin(theta) ** 2, 1e-12)
    cos2 = 1.0 - sin2
    sigma = r2 + a2 * cos2
    delta = r2 - 2.0 * m * r + a2
    g = [0.0] * 16
    g[0] = -(1.0 + 2.0 * m * r / sigma)
    g[1] = g[4] = 2.0 * m * r / sigma
    g[5] = delta / sigma
    g[10] = 1.0 / sigma
    g[15] = 1.0 / (sigma * sin2)
    g[7] = g[13] = a / sigma
    return g


def dh_ks(m, a, r, theta, p):  # kerr.rs:442-499
    r2, a2 = r * r, a * a
    st, ct = math.sin(theta), math.cos(theta)
    sin2 = max(st * st, 1e-12)
    cos2 = 1.0 - sin2
    sigma = r2 + a2 * cos2
    sigma2 = sigma * sigma
    delta = r2 - 2.0 * m * r + a2
    ds_dr = 2.0 * r
    ds_dth = -2.0 * a2 * st * ct
    dd_dr = 2.0 * r - 2.0 * m
    dgtt_dr = -(2.0 * m * (sigma - r * ds_dr)) / sigma2
    dgtt_dth = (2.0 * m * r * ds_dth) / sigma2
    dgtr_dr, dgtr_dth = -dgtt_dr, -dgtt_dth
    dgrr_dr = (dd_dr * sigma - delta * ds_dr) / sigma2
    dgrr_dth = -(delta * ds_dth) / sigma2
    dgthth_dr = -ds_dr / sigma2
    dgthth_dth = -ds_dth / sigma2
    dgphph_dr = -ds_dr / (sigma2 * sin2)
    dgphph_dth = -(ds_dth * sin2 + sigma * 2.0 * st * ct) / (sigma2 * sin2 * sin2)
    dgrph_dr = -(a * ds_dr) / sigma2
    dgrph_dth = -(a * ds_dth) / sigma2
    dr = 0.5 * (dgtt_dr * p[0] * p[0] + dgrr_dr * p[1] * p[1] + dgthth_dr * p[2] * p[2] + dgphph_dr * p[3] * p[3]
                + 2.0 * dgtr_dr * p[0] * p[1] + 2.0 * dgrph_dr * p[1] * p[3])
    dth = 0.5 * (dgtt_dth * p[0] * p[0] + dgrr_dth * p[1] * p[1] + dgthth_dth * p[2] * p[2] + dgphph_dth * p[3] * p[3]
                 + 2.0 * dgtr_dth * p[0] * p[1] + 2.0 * dgrph_dth * p[1] * p[3])
    if abs(st) < 1e-10:
        dth = 0.0
    return dr, dth


def rhs(m, a, s):  # geodesic/hamiltonian.rs:13-35 ; s = [t r th ph pt pr pth pph]
    g = contravariant_ks(m, a, s[1], s[2])
    p = s[4:8]
    dt = g[0] * p[0] + g[1] * p[1] + g[3] * p[3]
    dr = g[4] * p[0] + g[5] * p[1] + g[7] * p[3]
    dth = g[10] * p[2]
    dph = g[12] * p[0] + g[13] * p[1] + g[15] * p[3]
    hr, hth = dh_ks(m, a, s[1], s[2], p)
    return [dt, dr, dth, dph, 0.0, -hr, -hth, 0.0]


def hamiltonian(m, a, s):  # invariants/mod.rs:25-37
    g = contravariant_ks(m, a, s[1], s[2])
    p = s[4:8]
    return 0.5 * (g[0] * p[0] * p[0] + g[5] * p[1] * p[1] + g[10] * p[2] * p[2] + g[15] * p[3] * p[3]
                  + 2.0 * g[3] * p[0] * p[3] + 2.0 * g[1] * p[0] * p[1] + 2.0 * g[7] * p[1] * p[3])


def renormalize(m, a, s):  # invariants/renormalization.rs:13-45
    g = contravariant_ks(m, a, s[1], s[2])
    pt, pr, pth, pph = s[4:8]
    A = g[5]
    B = 2.0 * (g[1] * pt + g[7] * pph)
    Cq = g[0] * pt * pt + g[10] * pth * pth + g[15] * pph * pph + 2.0 * g[3] * pt * pph
    out = list(s)
    if abs(A) > 1e-12:
        disc = B * B - 4.0 * A * Cq
        if disc >= 0.0:
            sq = math.sqrt(disc)
            s1, s2 = (-B + sq) / (2.0 * A), (-B - sq) / (2.0 * A)
            out[5] = s1 if abs(s1 - pr) < abs(s2 - pr) else s2
    return out


def step_symplectic(m, a, s, h):  # geodesic/integrator.rs:209-226
    mid = list(s)
    for _ in range(2):
        d = rhs(m, a, mid)
        nxt = [s[i] + d[i] * h for i in range(8)]
        mid = [0.5 * (s[i] + nxt[i]) for i in range(8)]
    d = rhs(m, a, mid)
    return [s[i] + d[i] * h for i in range(8)]


def step_rk4(m, a, s, h):  # geodesic/integrator.rs:193-203
    def add(s0, k, c):
        return [s0[i] + k[i] * c for i in range(8)]
    k1 = rhs(m, a, s)
    k2 = rhs(m, a, add(s, k1, 0.5 * h))
    k3 = rhs(m, a, add(s, k2, 0.5 * h))
    k4 = rhs(m, a, add(s, k3, h))
    return [s[i] + (h / 6.0) * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]) for i in range(8)]


# ---- physics/redshift.rs:65-95 ---------------------------------------------------------------------------------

def g_factor(r, mass, spin, lam):
    a = spin * mass
    r2, a2, m = r * r, a * a, mass
    omega = math.sqrt(m) / (r ** 1.5 + a * math.sqrt(m))
    sigma = r2
    g_tt = -(1.0 - 2.0 * m * r / sigma)
    g_tphi = -(2.0 * m * r * a) / sigma
    g_pp = r2 + a2 + 2.0 * m * r * a2 / sigma
    den = -g_tt - 2.0 * omega * g_tphi - omega * omega * g_pp
    if den <= 0.0:
        return 0.0
    ut = 1.0 / math.sqrt(den)
    f = 1.0 - lam * omega
    if abs(f) < 1e-30:
        return 0.0
    return 1.0 / (ut * f)


# ---- physics/spectrum.rs:12-102 (one texel) ----------------------------------------------------------------------
_H, _C, _KB = 6.62607015e-34, 299792458.0, 1.380649e-23
_C1, _C2 = 2.0 * _H * _C * _C, _H * _C / _KB


def planck(lam, T):
    ex = _C2 / (lam * T)
    if ex > 100.0:
        return 0.0
    return (_C1 / (lam * lam * lam * lam * lam)) / (math.exp(ex) - 1.0)


def spectrum_texel(x, y, W, H, max_temp):
    """-> (r, g, b) as Python floats holding the f32 values generate_blackbody_lut stores."""
    import numpy as np
    g = 0.05 + (5.0 - 0.05) * (y / max(H - 1, 1))
    t = (x / max(W - 1, 1)) ** 2.5 * max_temp
    teff = t * g
    X = Y = Z = 0.0
    if teff >= 100.0:
        lam, end, step = 380.0e-9, 780.0e-9, 2.0e-9
        while lam <= end:
            inten = planck(lam, teff)
            l_nm = lam * 1e9

            def gs(mean, sd):
                q = (l_nm - mean) / sd
                return math.exp(-0.5 * q * q)
            cx = max(1.056 * gs(599.0, 37.9) + 0.362 * gs(442.0, 16.0) - 0.065 * gs(501.0, 20.4), 0.0)
            cy = max(0.821 * gs(568.0, 46.9) + 0.286 * gs(530.0, 22.1), 0.0)
            cz = max(1.217 * gs(437.0, 11.8) + 0.681 * gs(459.0, 26.0), 0.0)
            X += inten * cx * step
            Y += inten * cy * step
            Z += inten * cz * step
            lam += step
    rr = 3.2404542 * X - 1.5371385 * Y - 0.4985314 * Z
    gg = -0.9692660 * X + 1.8760108 * Y + 0.0415560 * Z
    bb = 0.0556434 * X - 0.2040259 * Y + 1.0572252 * Z
    scale = np.float32(1.0e-14 * (g * g * g * g))
    return tuple(float(np.float32(max(c, 0.0)) * scale) for c in (rr, gg, bb))


# ---- compute.wgsl.ts:159-187 camera -> state, in f64 from the f32 uniform block --------------------------------

def camera_ray(cam88, W, H, px, py):
    inv_view = [float(v) for v in cam88[32:48]]
    inv_proj = [float(v) for v in cam88[48:64]]
    uvx, uvy = px / W, py / H
    ndcx, ndcy = uvx * 2.0 - 1.0, uvy * 2.0 - 1.0
    clip = (ndcx, -ndcy, 1.0, 1.0)
    vt = [sum(inv_proj[c * 4 + r] * clip[c] for c in range(4)) for r in range(4)]
    v = [vt[0] / vt[3], vt[1] / vt[3], vt[2] / vt[3]]
    n = math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    v = [c / n for c in v]
    w = [inv_view[0 * 4 + r] * v[0] + inv_view[1 * 4 + r] * v[1] + inv_view[2 * 4 + r] * v[2] for r in range(3)]
    n = math.sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2])
    d = [c / n for c in w]
    cx, cy, cz = float(cam88[80]), float(cam88[81]), float(cam88[82])
    r0 = math.sqrt(cx * cx + cy * cy + cz * cz)
    th0 = math.acos(min(1.0, max(-1.0, cy / r0)))
    ph0 = math.atan2(cz, cx)
    st, ct, sp, cp = math.sin(th0), math.cos(th0), math.sin(ph0), math.cos(ph0)
    pr = d[0] * (st * cp) + d[1] * ct + d[2] * (st * sp)
    pth = (d[0] * (ct * cp) + d[1] * (-st) + d[2] * (ct * sp)) / r0
    pph = (d[0] * (-sp) + d[2] * cp) / (r0 * max(st, 1e-4))
    return [0.0, r0, th0, ph0, -1.0, pr, pth * r0 * r0, pph * r0 * r0 * st * st]


# ---- composite pixel (DESIGN.md "Composite RGBA definition") ----------------------------------------------------

def sample_tdisk(td, rin, rout, r):
    n = len(td)
    t = (r - rin) / (rout - rin) * (n - 1)
    t = min(max(t, 0.0), float(n - 1))
    i0 = int(math.floor(t))
    i1 = min(i0 + 1, n - 1)
    f = t - math.floor(t)
    return float(td[i0]) + (float(td[i1]) - float(td[i0])) * f


def sample_spectrum(lut, W, H, u, v):
    x = min(max(u * W - 0.5, 0.0), float(W - 1))
    y = min(max(v * H - 0.5, 0.0), float(H - 1))
    x0, y0 = int(math.floor(x)), int(math.floor(y))
    x1, y1 = min(x0 + 1, W - 1), min(y0 + 1, H - 1)
    fx, fy = x - x0, y - y0
    out = []
    for c in range(3):
        t00, t10 = float(lut[4 * (y0 * W + x0) + c]), float(lut[4 * (y0 * W + x1) + c])
        t01, t11 = float(lut[4 * (y1 * W + x0) + c]), float(lut[4 * (y1 * W + x1) + c])
        top = t00 + (t10 - t00) * fx
        bot = t01 + (t11 - t01) * fx
        out.append(top + (bot - top) * fy)
    return out


def render_pixel(cam88, W, H, px, py, mass, spin, max_steps, lut, LW, LH, td, r_out=50.0, escape=1000.0, renorm=10):
    """Fixed-step (implicit midpoint + compute.wgsl.ts:213 rule) march with thin-disk shading. -> (rgb, term, steps)"""
    a = spin * mass
    rh = horizon(mass, spin)
    r_in = isco(mass, spin)
    s = renormalize(mass, a, camera_ray(cam88, W, H, px, py))
    col, alpha, steps, term = [0.0, 0.0, 0.0], 0.0, 0, 3
    for _ in range(max_steps):
        if s[1] < rh * 1.001:
            term = 1
            break
        if s[1] > escape:
            term = 2
            break
        prev = s
        h = min(max((s[1] - rh) * 0.15, 0.05), 1.0)
        s = step_symplectic(mass, a, s, h)
        if steps % renorm == 0:
            s = renormalize(mass, a, s)
        steps += 1
        d0, d1 = prev[2] - math.pi / 2, s[2] - math.pi / 2
        if d0 * d1 <= 0.0:
            dth = s[2] - prev[2]
            f = 0.0 if dth == 0.0 else (math.pi / 2 - prev[2]) / dth
            rc = prev[1] + f * (s[1] - prev[1])
            if r_in < rc < r_out:
                lam = s[7] / (-s[4])
                g = g_factor(rc, mass, spin, lam)
                tn = sample_tdisk(td, r_in, 50.0 * mass, rc)
                rgb = sample_spectrum(lut, LW, LH, tn ** 0.4, (g - 0.05) / 4.95)
                op = 0.6 * tn * g
                wgt = (1.0 - alpha) * op
                col = [col[c] + rgb[c] * wgt for c in range(3)]
                alpha += op
        if alpha > 0.99:
            term = 4
            break
    return col, term, steps


# ---- physics/shadow.rs : Bardeen critical curve (independent restatement from the Rust, CPython libm) -----------
def photon_sphere(m, spin):  # kerr.rs photon_sphere(): prograde circular photon orbit 2M(1 + cos(2/3 acos(-|a*|)))
    return 2.0 * m * (1.0 + math.cos((2.0 / 3.0) * math.acos(-abs(spin))))


def critical_params(r, m, a):  # shadow.rs:38-58
    r2 = r * r
    r3 = r2 * r
    a2 = a * a
    denom = a * (r - m)
    if abs(denom) < 1e-30:
        return 0.0, 0.0
    xi = -(r3 - 3.0 * m * r2 + a2 * r + a2 * m) / denom
    denom2 = a2 * (r - m) * (r - m)
    if abs(denom2) < 1e-30:
        return xi, 0.0
    q = r - 3.0 * m
    eta = r3 * (4.0 * m * a2 - r * (q * q)) / denom2
    return xi, eta


def bardeen_shadow(m, spin, theta_obs, n_points):  # shadow.rs:81-183
    a = spin * m
    so, co = math.sin(theta_obs), math.cos(theta_obs)
    if abs(a) < 1e-10:
        rad = 3.0 * math.sqrt(3.0) * m
        return [(rad * math.cos(2.0 * math.pi * i / n_points), rad * math.sin(2.0 * math.pi * i / n_points))
                for i in range(n_points)]
    if abs(so) < 1e-10:
        xi, eta = critical_params(photon_sphere(m, spin), m, a)
        rad = math.sqrt(max(eta + a * a, 0.0))
        return [(rad * math.cos(2.0 * math.pi * i / (2.0 * n_points)), rad * math.sin(2.0 * math.pi * i / (2.0 * n_points)))
                for i in range(2 * n_points)]
    a_star = a / m
    r_pro = 2.0 * m * (1.0 + math.cos((2.0 / 3.0) * math.acos(-abs(a_star))))
    r_ret = 2.0 * m * (1.0 + math.cos((2.0 / 3.0) * math.acos(abs(a_star))))

    def beta_sq(r):
        xi, eta = critical_params(r, m, a)
        return eta + a * a * co * co - xi * xi * co * co / (so * so), xi

    r_min, r_max = r_pro, r_ret
    steps = 1000
    for i in range(steps + 1):
        r = r_pro + (i / steps) * (r_ret - r_pro)
        if beta_sq(r)[0] >= 0.0:
            r_min = r
            break
    for i in range(steps, -1, -1):
        r = r_pro + (i / steps) * (r_ret - r_pro)
        if beta_sq(r)[0] >= 0.0:
            r_max = r
            break
    pts = []
    for sign, order in ((-1.0, range(n_points)), (1.0, range(n_points - 1, -1, -1))):
        for i in order:
            phase = math.pi * i / max(n_points - 1, 1)
            t = 0.5 - 0.5 * math.cos(phase)
            r = r_min + t * (r_max - r_min)
            b2, xi = beta_sq(r)
            pts.append((a * so - xi / so, sign * math.sqrt(max(b2, 0.0))))
    return pts


# ---- gravitas-core/src/spacetime/*.rs : visualisation helpers (independent restatement from the Rust) ----------------
def covariant_bl(m, spin, r, theta):  # kerr.rs:241-264 -> (g_tt, g_rr, g_thth, g_phph, g_tph)
    a = spin * m
    r2, a2 = r * r, a * a
    s, c = math.sin(theta), math.cos(theta)
    sin2, cos2 = s * s, c * c
    sigma = r2 + a2 * cos2
    delta = r2 - 2.0 * m * r + a2
    return (-(1.0 - (2.0 * m * r) / sigma), sigma / delta, sigma, (r2 + a2 + (2.0 * m * r * a2 * sin2) / sigma) * sin2,
            -(2.0 * m * r * a * sin2) / sigma)


def kretschner_kerr(r, theta, mass, spin):  # curvature.rs:13-36
    a = spin * mass
    r2, a2 = r * r, a * a
    cos2 = math.cos(theta) ** 2
    cos4 = cos2 * cos2
    cos6 = cos4 * cos2
    r4 = r2 * r2
    r6 = r4 * r2
    a4 = a2 * a2
    a6 = a4 * a2
    sigma = r2 + a2 * cos2
    sigma6 = (sigma * sigma) * ((sigma * sigma) * (sigma * sigma))
    if sigma6 < 1e-30:
        return math.inf
    return 48.0 * mass * mass * (r6 - 15.0 * r4 * a2 * cos2 + 15.0 * r2 * a4 * cos4 - a6 * cos6) / sigma6


def light_cone_tilt_bl(m, spin, r, theta):  # lightcone.rs:18-34 (diagonal branch)
    g_tt, g_rr = covariant_bl(m, spin, r, theta)[:2]
    if g_tt >= 0.0:
        return math.pi / 2
    return math.atan(math.sqrt(max(-g_tt / g_rr, 0.0)))


def frame_dragging(m, spin, r, theta):  # kerr.rs:143-152
    g = covariant_bl(m, spin, r, theta)
    return 0.0 if abs(g[3]) < 1e-30 else -g[4] / g[3]


def field_lattice(r_min, r_max, n_radial, n_polar):
    for i in range(n_radial):
        r = r_min + (r_max - r_min) * i / (n_radial - 1)
        for j in range(n_polar):
            yield r, 0.1 + (math.pi - 0.2) * j / (n_polar - 1)


def flamm_height(r, mass):  # embedding.rs:14-20
    rs = 2.0 * mass
    return 0.0 if r <= rs else 2.0 * math.sqrt(rs * (r - rs))


def kerr_embedding_height(m, spin, r, r_ref, n_steps):  # embedding.rs:28-44
    dr = (r_ref - r) / n_steps
    z = 0.0
    for i in range(n_steps):
        g_rr = covariant_bl(m, spin, r + (i + 0.5) * dr, math.pi / 2)[1]
        z += math.sqrt(abs(g_rr - 1.0)) * dr
    return z


def proper_distance(m, spin, r1, r2, n_steps):  # embedding.rs:49-63
    lo, hi = (r1, r2) if r1 < r2 else (r2, r1)
    dr = (hi - lo) / n_steps
    return sum(math.sqrt(abs(covariant_bl(m, spin, lo + (i + 0.5) * dr, math.pi / 2)[1])) * dr for i in range(n_steps))


def embedding_mesh(mass, spin, r_min, r_max, n_radial, n_angular):  # embedding.rs:72-110
    cl = max(-1.0, min(1.0, spin))
    out = []
    for i in range(n_radial):
        r = r_min + (i / (n_radial - 1)) * (r_max - r_min)
        h = flamm_height(r, mass) if abs(spin) < 1e-6 else kerr_embedding_height(mass, cl, r, r_max, 100)
        for j in range(n_angular):
            phi = 2.0 * math.pi * j / n_angular
            out += [r * math.cos(phi), -h, r * math.sin(phi)]
    return out


def ergosphere_mesh(m, spin, n_polar, n_azimuthal):  # frame_drag.rs:48-68, kerr.rs:157-167
    a = spin * m
    out = []
    for i in range(n_polar):
        theta = math.pi * i / (n_polar - 1)
        disc = m * m - a * a * math.cos(theta) ** 2
        r_e = m if disc < 0.0 else m + math.sqrt(disc)
        for j in range(n_azimuthal):
            phi = 2.0 * math.pi * j / n_azimuthal
            out += [r_e * math.sin(theta) * math.cos(phi), r_e * math.cos(theta), r_e * math.sin(theta) * math.sin(phi)]
    return out
