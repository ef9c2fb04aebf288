"""Independent pure-Python (CPython float = IEEE f64, libm) restatement of the reference's WebGL2 fragment shader,
written from the GLSL text (src/shaders/blackhole/fragment.glsl.ts and chunks/*.ts), NOT from the C++ oracle. It exists
to cross-check oracle/glsl_fragment_oracle.hpp pixel by pixel (tests/test_oracle_glsl.py): two restatements agreeing
is not the reference agreeing, but a transcription slip would have to be made twice, identically.

Uniforms are a plain dict with the GLSL names (u_time, u_mass, ...); `defines` is a set of the shader manager's
macro names; textures are (256, 256) uint8 arrays holding the .r channel."""
import math

PI = 3.14159265359
MAX_DIST = 10000.0
MIN_STEP = 0.01
MAX_STEP = 1.2


# ---- tiny vector helpers (GLSL semantics) ---------------------------------------------------------------------------
def add(a, b): return [a[0] + b[0], a[1] + b[1], a[2] + b[2]]
def sub(a, b): return [a[0] - b[0], a[1] - b[1], a[2] - b[2]]
def mul(a, s): return [a[0] * s, a[1] * s, a[2] * s]
def dot(a, b): return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
def cross(a, b): return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]
def length(a): return math.sqrt(dot(a, a))
def normalize(a):
    n = length(a)
    return [a[0] / n, a[1] / n, a[2] / n]
def gmax(a, b): return b if a < b else a
def gmin(a, b): return b if b < a else a
def clamp(x, lo, hi): return gmin(gmax(x, lo), hi)
def mix(a, b, t): return a * (1.0 - t) + b * t
def mix3(a, b, t): return [mix(a[0], b[0], t), mix(a[1], b[1], t), mix(a[2], b[2], t)]
def smoothstep(e0, e1, x):
    t = clamp((x - e0) / (e1 - e0), 0.0, 1.0)
    return t * t * (3.0 - 2.0 * t)
def sign(x): return 1.0 if x > 0.0 else (-1.0 if x < 0.0 else 0.0)
def fract(x): return x - math.floor(x)
def rot_apply(ang, a, b):
    """(a, b) *= rot(ang) with rot(a) = mat2(c, -s, s, c): GLSL `v.ab *= m` is the row vector times the matrix."""
    s, c = math.sin(ang), math.cos(ang)
    return a * c - b * s, a * s + b * c


class Shader:
    def __init__(self, U, defines, noise_r, blue_r):
        self.U, self.D, self.noise, self.blue = U, set(defines), noise_r, blue_r

    # chunks/noise.ts
    def texture_noise(self, u, v):
        x, y = u * 256.0 - 0.5, v * 256.0 - 0.5
        x0, y0 = math.floor(x), math.floor(y)
        fx, fy = x - x0, y - y0
        xi, yi = int(x0), int(y0)
        t = lambda ix, iy: float(self.noise[iy % 256][ix % 256]) / 255.0
        return mix(mix(t(xi, yi), t(xi + 1, yi), fx), mix(t(xi, yi + 1), t(xi + 1, yi + 1), fx), fy)

    def hash(self, p):
        uvx, uvy = p[0] + p[2] * 37.0, p[1] + p[2] * 37.0
        return self.texture_noise((uvx + 0.5) / 256.0, (uvy + 0.5) / 256.0)

    def noise3(self, p):
        i = [math.floor(p[0]), math.floor(p[1]), math.floor(p[2])]
        f = [fract(p[0]), fract(p[1]), fract(p[2])]
        f = [c * c * (3.0 - 2.0 * c) for c in f]
        h = lambda a, b, c: self.hash([i[0] + a, i[1] + b, i[2] + c])
        return mix(mix(mix(h(0, 0, 0), h(1, 0, 0), f[0]), mix(h(0, 1, 0), h(1, 1, 0), f[0]), f[1]),
                   mix(mix(h(0, 0, 1), h(1, 0, 1), f[0]), mix(h(0, 1, 1), h(1, 1, 1), f[0]), f[1]), f[2])

    def fbm(self, p):
        f, amp = 0.0, 0.5
        for _ in range(4):
            f += amp * self.noise3(p)
            p = mul(p, 2.0)
            amp *= 0.5
        return f

    # chunks/blackbody.ts
    @staticmethod
    def blackbody(temp):
        t = gmax(temp, 1.0) / 100.0
        if t <= 66.0:
            r = 255.0
            g = 99.4708025861 * math.log(t) - 161.1195681661
            b = 0.0 if t <= 19.0 else 138.5177312231 * math.log(t - 10.0) - 305.0447927307
        else:
            r = 329.698727446 * math.pow(t - 60.0, -0.1332047592)
            g = 288.1221695283 * math.pow(t - 60.0, -0.0755148492)
            b = 255.0
        return [math.pow(gmax(c / 255.0, 0.0), 2.2) for c in (r, g, b)]

    @staticmethod
    def star_color(bv):
        t = clamp(bv, -0.4, 2.0)
        if t < 0.0: return [0.6, 0.7, 1.0]
        if t < 0.3: return [0.85, 0.88, 1.0]
        if t < 0.6: return [1.0, 0.96, 0.9]
        if t < 1.0: return [1.0, 0.85, 0.6]
        return [1.0, 0.6, 0.4]

    # chunks/background.ts
    def starfield(self, d):
        time = self.U["u_time"]
        stars = [0.0, 0.0, 0.0]
        cell = [math.floor(c * 200.0) for c in d]
        sn = self.hash(cell)
        if sn > 0.998:
            brightness = math.pow(sn, 10.0) * 2.0
            bv = self.hash([c + 127.1 for c in cell]) * 2.4 - 0.4
            twinkle = 0.85 + 0.15 * math.sin(time * (3.0 + self.hash([c + 73.7 for c in cell]) * 2.0))
            stars = mul(mul(self.star_color(bv), brightness), twinkle)
        cell = [math.floor(c * 500.0) for c in d]
        sn = self.hash(cell)
        if sn > 0.996:
            brightness = math.pow(sn, 20.0) * 1.5
            bv = self.hash([c + 217.3 for c in cell]) * 2.4 - 0.4
            stars = add(stars, mul(self.star_color(bv), brightness))
        nebula = self.fbm([c * 2.0 + time * 0.01 for c in d]) * 0.03
        glow = add([nebula * 0.2, nebula * 0.3, nebula * 0.5], mul([0.05, 0.02, 0.05], abs(nebula)))
        return add(stars, glow)

    # chunks/metric.ts
    @staticmethod
    def kerr_horizon(M, a): return M + math.sqrt(gmax(0.0, M * M - a * a))

    @staticmethod
    def kerr_isco(M, a):
        absS = abs(clamp(a / M, -0.9999, 0.9999))
        z1 = 1.0 + math.pow(1.0 - absS * absS, 1.0 / 3.0) * (math.pow(1.0 + absS, 1.0 / 3.0) + math.pow(1.0 - absS, 1.0 / 3.0))
        z2 = math.sqrt(3.0 * absS * absS + z1 * z1)
        sg = sign(a)
        if sg == 0.0: sg = 1.0
        return M * (3.0 + z2 - sg * math.sqrt((3.0 - z1) * (3.0 + z1 + 2.0 * z2)))

    @staticmethod
    def kerr_photon_sphere(M, a):
        a_star = clamp(a / M, -0.9999, 0.9999)
        return 2.0 * M * (1.0 + math.cos((2.0 / 3.0) * math.acos(clamp(-a_star, -1.0, 1.0))))

    @staticmethod
    def kerr_geodesic_accel(p, v, M, a):
        a2 = a * a
        rho2 = dot(p, p)
        diff = rho2 - a2
        disc = diff * diff + 4.0 * a2 * p[1] * p[1]
        r2 = 0.5 * (diff + math.sqrt(gmax(0.0, disc)))
        r_k = math.sqrt(gmax(1e-8, r2))
        sigma = r2 + a2 * (p[1] * p[1] / gmax(1e-8, r2))
        L = cross(p, v)
        Ly = L[1]
        Ly_eff = Ly - a
        L2_eff = Ly_eff * Ly_eff + (dot(L, L) - Ly * Ly)
        r_inv = 1.0 / r_k
        r2_inv = r_inv * r_inv
        r4_inv = r2_inv * r2_inv
        sigma_ratio = r2 / gmax(1e-8, sigma)
        nh = normalize(p)
        r_hat = [-nh[0], -nh[1], -nh[2]]
        accel = mul(r_hat, M * r2_inv * sigma_ratio + 3.0 * M * gmax(0.0, L2_eff) * r4_inv * sigma_ratio)
        r3_p_a2r = r_k * r2 + a2 * r_k
        drag = 2.0 * M * a / gmax(1e-8, r3_p_a2r)
        accel = add(accel, mul(cross([0.0, 1.0, 0.0], v), drag))
        omega = 2.0 * M * a / gmax(1e-8, r3_p_a2r)
        return accel, omega

    # chunks/disk.ts
    def sample_accretion_disk(self, p, p_prev, v, isco, M, a, dt, accC, accA):
        U = self.U
        if not (U["u_show_redshift"] < 0.5):
            return accC, accA
        crossed = p_prev[1] * p[1] < 0.0
        sampleP = list(p)
        if crossed:
            t = abs(p_prev[1]) / gmax(0.0001, abs(p_prev[1]) + abs(p[1]))
            sampleP = mix3(p_prev, p, t)
        sampleR = length(sampleP)
        esh = gmin(U["u_disk_scale_height"], 0.450)
        diskHeight = sampleR * esh
        diskInner = isco
        diskOuter = gmax(M * U["u_disk_size"], diskInner * 1.1)
        if not ((abs(sampleP[1]) < diskHeight or crossed) and sampleR > diskInner and sampleR < diskOuter):
            return accC, accA
        sqrtM = math.sqrt(M)
        sgn = sign(U["u_spin"] + 1e-8)
        OmegaPhase = (sgn * sqrtM) / (sampleR * math.sqrt(sampleR) + a * sqrtM)
        rotAngle = OmegaPhase * U["u_time"] * 0.12 * 10.0
        c, s = math.cos(rotAngle), math.sin(rotAngle)
        noiseP = [sampleP[0] * c - sampleP[2] * s, sampleP[1], sampleP[0] * s + sampleP[2] * c]   # xz *= mat2(c, -s, s, c)
        noiseP = mul(noiseP, 0.75)
        turbulence = self.noise3(noiseP) * 0.5 + self.noise3(mul(noiseP, 2.5)) * 0.25
        heightFalloff = math.exp(-abs(sampleP[1]) / gmax(0.001, (sampleR * esh) * 0.25))
        radialFalloff = smoothstep(diskOuter, diskInner, sampleR)
        baseDensity = turbulence * heightFalloff * radialFalloff
        if not (baseDensity > 0.001):
            return accC, accA
        r2 = sampleR * sampleR
        Omega = (sgn * sqrtM) / (sampleR * math.sqrt(sampleR) + a * sqrtM)
        g_tt = -(1.0 - 2.0 * M / sampleR)
        g_tphi = -2.0 * M * a / sampleR
        g_phiphi = r2 + a * a + 2.0 * M * a * a / sampleR
        u_t_sq = -(g_tt + 2.0 * Omega * g_tphi + Omega * Omega * g_phiphi)
        u_t = 1.0 / math.sqrt(gmax(1e-6, u_t_sq))
        L_photon = p[2] * v[0] - p[0] * v[2]
        delta = 1.0 / gmax(0.01, u_t * (1.0 - Omega * L_photon))
        beaming = gmax(0.01, math.pow(delta, 3.5)) if "ENABLE_DOPPLER" in self.D else 1.0
        isco_r = clamp(isco / sampleR, 0.0, 1.0)
        nt_factor = gmax(0.0, 1.0 - math.sqrt(isco_r))
        grad = math.pow(isco_r, 0.75) * math.pow(nt_factor, 0.25)
        temperature = U["u_disk_temp"] * grad * delta
        diskColor = mul(self.blackbody(temperature), beaming)
        density = baseDensity * U["u_disk_density"] * 0.12 * dt
        accC = add(accC, mul(mul(diskColor, density), 1.0 - accA))
        accA = accA + density
        return accC, accA

    def sample_relativistic_jets(self, p, v, rh, dt, accC, accA):
        jv = abs(p[1])
        if not (jv > rh * 1.8 and jv < MAX_DIST * 0.8):
            return accC, accA
        jr = math.sqrt(p[0] * p[0] + p[2] * p[2])
        jw = 1.0 + jv * 0.15
        if not (jr < jw * 2.0):
            return accC, accA
        radialFalloff = math.exp(-(jr * jr) / (jw * 0.5))
        lengthFalloff = math.exp(-jv * 0.05)
        flow = p[1] * 2.0 - self.U["u_time"] * 8.0
        uvJet = [p[0], flow, p[2]]
        noiseVal = self.noise3(mul(uvJet, 0.5)) * 0.6 + self.noise3(mul(uvJet, 1.5)) * 0.4
        jd = radialFalloff * lengthFalloff * gmax(0.0, noiseVal - 0.2)
        if not (jd > 0.001):
            return accC, accA
        jetVel = 0.92 * sign(p[1])
        cosT = dot(normalize([0.0, jetVel, 0.0]), [-v[0], -v[1], -v[2]])
        beta = abs(jetVel)
        gamma = 1.0 / math.sqrt(1.0 - beta * beta)
        dj = 1.0 / (gamma * (1.0 - beta * cosT))
        beam = math.pow(dj, 3.5)
        emission = mul(mul(mul(mul([0.4, 0.7, 1.0], jd), 0.05), beam), dt)
        accC = add(accC, mul(emission, 1.0 - accA))
        accA = accA + jd * 0.05 * dt
        return accC, accA

    # fragment.glsl.ts main()
    def main(self, px, py):
        U, D = self.U, self.D
        res = U["u_resolution"]
        fc = [px + 0.5, py + 0.5]
        minRes = min(res[0], res[1])
        uv = [(fc[0] - 0.5 * res[0]) / minRes, (fc[1] - 0.5 * res[1]) / minRes]
        if U["u_debug"] > 0.5:
            return [uv[0] + 0.5, uv[1] + 0.5, 0.0], 0, False
        if length(U["u_camPos"]) > 0.001:
            ro = list(U["u_camPos"])
            d = normalize([uv[0], uv[1], 1.2])
            q = U["u_camQuat"]
            qv = q[:3]
            rd = add(d, mul(cross(qv, add(cross(qv, d), mul(d, q[3]))), 2.0))
        else:
            ro = [0.0, 0.0, -U["u_zoom"]]
            rd = normalize([uv[0], uv[1], 1.5])
            ax = (U["u_mouse"][1] - 0.5) * PI
            ay = (U["u_mouse"][0] - 0.5) * PI * 2.0
            ro[1], ro[2] = rot_apply(ax, ro[1], ro[2]); rd[1], rd[2] = rot_apply(ax, rd[1], rd[2])
            ro[0], ro[2] = rot_apply(ay, ro[0], ro[2]); rd[0], rd[2] = rot_apply(ay, rd[0], rd[2])
        M = U["u_mass"]
        rs = M * 2.0
        a = U["u_spin"] * M
        rh = self.kerr_horizon(M, a)
        rph = self.kerr_photon_sphere(M, a)
        isco = self.kerr_isco(M, a)
        absA = abs(U["u_spin"])
        if "RAY_QUALITY_LOW" in D or "RAY_QUALITY_OFF" in D:
            bg = self.starfield(rd)
            dd = length(cross(ro, rd))
            shadow = smoothstep(rh * 1.2, rh * 0.9, dd)
            glow = math.exp(-abs(dd - rph) * 12.0) * 0.8
            mask = smoothstep(isco * 2.0, isco * 1.0, dd) * (1.0 - smoothstep(isco * 1.0, isco * 0.8, dd))
            col = add(add(mul(bg, 1.0 - shadow), mul([0.3, 0.6, 1.0], glow)), mul(mul([1.0, 0.7, 0.3], mask), 0.6))
            return [math.pow(c, 0.4545) for c in col], 0, False
        p, v = list(ro), list(rd)
        if length(ro) < rh * 1.5:
            ro = mul(mul(normalize(ro), rh), 1.5)
            p = list(ro)
        accC, accA = [0.0, 0.0, 0.0], 0.0
        hit = False
        maxRedshift = 0.0
        bNoise = float(self.blue[py % 256][px % 256]) / 255.0
        p = add(p, mul(mul(v, bNoise), MIN_STEP))
        photon = 0
        prevY = p[1]
        impact = length(cross(ro, rd))
        red_init = False
        maxSteps = int(min(float(U["u_maxRaySteps"]), 500.0))
        if impact < rh * 0.9:
            hit = True
        lens = U["u_lensing_strength"]
        steps = 0
        for _ in range(maxSteps):
            p_prev = list(p)
            r = length(p)
            if r < rh * 1.15:
                hit = True
                break
            if r > MAX_DIST:
                break
            distFactor = 1.0 + r * 0.05
            dt = clamp((r - rh) * 0.1 * distFactor, MIN_STEP, MAX_STEP * distFactor)
            if r > 30.0:
                farBoost = (r - 30.0) * 0.08
                dt = gmax(dt, MIN_STEP + farBoost)
                dt = gmin(dt, MAX_STEP * 2.5)
            dt = gmin(dt, MIN_STEP + abs(r - rph) * 0.15)
            cdt = dt * (1.0 - smoothstep(0.2, 0.0, abs(p[1])) * 0.7)
            accel, omega = [0.0, 0.0, 0.0], 0.0
            if "ENABLE_LENSING" in D:
                accel, omega = self.kerr_geodesic_accel(p, v, M, a)
                accel = mul(accel, lens)
                v[0], v[2] = rot_apply(omega * cdt, v[0], v[2])
            p = add(p, add(mul(v, cdt), mul(mul(mul(accel, 0.5), cdt), cdt)))
            r_new = length(p)
            if "ENABLE_LENSING" in D and accA < 0.95:
                accel_new, _ = self.kerr_geodesic_accel(p, v, M, a)
                accel_new = mul(accel_new, lens)
                v = add(v, mul(mul(add(accel, accel_new), 0.5), cdt))
            v = normalize(v)
            if prevY * p[1] < 0.0 and r_new < rph * 2.0 and r_new > rh:
                photon = min(photon + 1, 3)
            if U["u_show_redshift"] > 0.5:
                pot = math.sqrt(gmax(0.0, 1.0 - rs / r_new))
                maxRedshift = pot if not red_init else gmin(maxRedshift, pot)
                red_init = True
            prevY = p[1]
            steps += 1
            if "ENABLE_DISK" in D:
                accC, accA = self.sample_accretion_disk(p, p_prev, v, isco, M, a, cdt, accC, accA)
                if accA > 0.99:
                    break
            if "ENABLE_JETS" in D:
                accC, accA = self.sample_relativistic_jets(p, v, rh, dt, accC, accA)
        if "ENABLE_REDSHIFT" in D and U["u_show_redshift"] > 0.5:
            val = 0.0 if hit else maxRedshift
            heat = mix3([0.0, 0.0, 0.0], [1.0, 0.0, 0.0], smoothstep(0.0, 0.3, val))
            heat = mix3(heat, [1.0, 1.0, 0.0], smoothstep(0.3, 0.7, val))
            heat = mix3(heat, [0.0, 0.0, 1.0], smoothstep(0.7, 1.0, val))
            return heat, steps, hit
        background = self.starfield(v) if "ENABLE_STARS" in D else [0.0, 0.0, 0.0]
        photonColor = [0.0, 0.0, 0.0]
        if "ENABLE_PHOTON_GLOW" in D and not hit:
            dpr = abs(length(p) - rph)
            direct = math.exp(-dpr * 40.0) * 1.8 * lens
            higher = 0.0
            if photon > 0:
                sharp = 60.0 + float(photon) * 30.0
                bright = math.exp(-float(photon) * 1.0) * 1.2
                higher = math.exp(-dpr * sharp) * bright * lens
            photonColor = mul([1.0, 1.0, 1.0], direct + higher)
        ergoColor = [0.0, 0.0, 0.0]
        if absA > 0.1 and not hit:
            rF = length(p)
            cosT = p[1] / gmax(rF, 0.001)
            r_ergo = M + math.sqrt(gmax(0.0, M * M - a * a * cosT * cosT))
            ergoColor = mul([0.3, 0.35, 0.9], math.exp(-abs(rF - r_ergo) * 20.0) * 0.35 * absA)
        if hit:
            background = [0.0, 0.0, 0.0]
        w = 1.0 - accA
        final = add(add(add(mul(background, w), accC), mul(photonColor, w)), mul(ergoColor, w))
        if U["u_show_kerr_shadow"] > 0.5:
            cam_dir = normalize(ro)
            sky_right = normalize(cross([0.0, 1.0, 0.0], cam_dir))
            sky_up = cross(cam_dir, sky_right)
            impact_vec = mul(cross(cam_dir, rd), length(ro))
            ps = [-dot(impact_vec, sky_up), dot(impact_vec, sky_right)]
            curve = U["u_shadowCurve"]
            count = int(U["u_shadowCount"])
            minDist = 1e10

            def seg(p1, p2, md):
                pa = [ps[0] - p1[0], ps[1] - p1[1]]
                ba = [p2[0] - p1[0], p2[1] - p1[1]]
                h = clamp((pa[0] * ba[0] + pa[1] * ba[1]) / (ba[0] * ba[0] + ba[1] * ba[1]), 0.0, 1.0)
                dx, dy = pa[0] - ba[0] * h, pa[1] - ba[1] * h
                return gmin(md, math.sqrt(dx * dx + dy * dy))
            for j in range(63):
                if j >= count - 1:
                    break
                minDist = seg(curve[j], curve[j + 1], minDist)
            if count > 2:
                minDist = seg(curve[count - 1], curve[0], minDist)
            thickness = M * 0.045
            if minDist < thickness:
                edge = smoothstep(thickness, thickness * 0.5, minDist)
                final = mix3(final, [0.0, 1.0, 0.0], 1.0 * edge)
        if "ENABLE_LINEAR_OUTPUT" not in D:
            A, B, C, Dd, E = 2.51, 0.03, 2.43, 0.59, 0.14
            final = [clamp((c * (A * c + B)) / (c * (C * c + Dd) + E), 0.0, 1.0) for c in final]
            final = [math.pow(gmax(c, 0.0), 0.4545) for c in final]
        return final, steps, hit
