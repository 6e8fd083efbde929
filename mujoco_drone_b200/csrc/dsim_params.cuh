// dsim_params.cuh — the "model compiler": drone_params (6 numbers) -> the 13 rigid-body constants the step
// kernel consumes.  Replaces env_gen.make_drone + mjcf_to_mjmodel (environments/env_gen.py:7-73,128-133) and the
// part of MuJoCo's XML compiler that matters here (`inertiafromgeom`: geom masses -> body mass, COM, inertia).
// FP64, __host__ __device__: runs inside a CUDA kernel for per-env domain randomisation (no host round trip
// on regen) and on the host for the uniform-parameter fast path.
#pragma once
#include <math.h>
#include "dsim_device.cuh"

namespace dsim {

// v * 10^p for |p| <= 22 (clamped): every power of ten up to 1e22 is an exact double, and so is each partial product below.
// (No lookup table: a local `const double P10[23]` inlined into a caller that keeps its own local array was seen to share
// that array's stack slots on the device - geometry_kernel, nvcc 12.9 - and overwrite it.)
DSIM_HD double scale10(double v, int p) {
    int k = p < 0 ? -p : p;
    if (k > 22) k = 22;
    double t = 1.0;
    for (int j = 0; j < k; j++) t *= 10.0;
    return p >= 0 ? v * t : v / t;
}
// dm_control writes every float attribute with "%.5g" (to_xml_string(precision=5), env_gen.py:129) and MuJoCo
// parses the decimal back: value -> nearest 5-significant-digit decimal -> nearest double.  k / 10^j with both
// exact is correctly rounded, i.e. what strtod returns.
DSIM_HD double round_prec5(double x) {
    if (x == 0.0 || !isfinite(x)) return x;
    const double ax = fabs(x);
    int e = (int)floor(log10(ax));
    double s = scale10(ax, 4 - e);
    if (s < 10000.0) { e -= 1; s = scale10(ax, 4 - e); }
    else if (s >= 100000.0) { e += 1; s = scale10(ax, 4 - e); }
    double k = rint(s);
    if (k >= 100000.0) { k = 10000.0; e += 1; }
    const double r = scale10(k, e - 4);
    return x < 0 ? -r : r;
}

// out[C_ROWS] in the order of the C_* enum (dsim_device.cuh)
DSIM_HD void compile_consts(const double p[6], bool pendulum_enabled, bool rounding, double out[13]) {
#define RND(v) (rounding ? round_prec5(v) : (v))
    const double mass = p[0], arm = p[1], force = p[2], tau = p[3], plen = p[4], wmass = p[5];
    const double hb = 0.05, r2 = sqrt(2.0);
    // --- core_body: 1 box + 4 x (arm box rotated by theta about z, motor cylinder)          env_gen.py:45-61
    double gm[9], gx[9], gy[9], gz[9], gI[9][3], gyaw[9];
    int ng = 0;
    {
        const double sx = RND(hb), sy = RND(hb), sz = RND(hb / 3), m = RND(0.56 * mass);
        gm[ng] = m; gx[ng] = gy[ng] = gz[ng] = 0; gyaw[ng] = 0;
        gI[ng][0] = m * (sy * sy + sz * sz) / 3; gI[ng][1] = m * (sx * sx + sz * sz) / 3; gI[ng][2] = m * (sx * sx + sy * sy) / 3;
        ng++;
    }
    double site_x = 0;
    for (int i = 0; i < 4; i++) {
        const double th = i * kPi / 2 - kPi / 4, ct = cos(th), st = sin(th);
        const double ra = r2 * hb + 0.5 * arm, rr = r2 * hb + arm;
        {
            const double sx = RND(arm / 2), sy = RND(arm / 20), sz = RND(arm / 20), m = RND(0.07 * mass);
            gm[ng] = m; gx[ng] = RND(ra * ct); gy[ng] = RND(ra * st); gz[ng] = 0; gyaw[ng] = RND(th);
            gI[ng][0] = m * (sy * sy + sz * sz) / 3; gI[ng][1] = m * (sx * sx + sz * sz) / 3; gI[ng][2] = m * (sx * sx + sy * sy) / 3;
            ng++;
        }
        {
            const double r = 0.01, hh = 0.01, m = RND(0.04 * mass);
            gm[ng] = m; gx[ng] = RND(rr * ct); gy[ng] = RND(rr * st); gz[ng] = RND(0.015); gyaw[ng] = 0;
            gI[ng][0] = gI[ng][1] = m * (3 * r * r + 4 * hh * hh) / 12; gI[ng][2] = m * r * r / 2;
            ng++;
        }
        if (i == 1) site_x = RND(rr * ct);          // motorsite_1 = (+s,+s,0); the other three are its mirror images
    }
    double M = 0, cx = 0, cy = 0, cz = 0;
    for (int k = 0; k < ng; k++) { M += gm[k]; cx += gm[k] * gx[k]; cy += gm[k] * gy[k]; cz += gm[k] * gz[k]; }
    cx /= M; cy /= M; cz /= M;
    double Ixx = 0, Iyy = 0, Izz = 0;
    for (int k = 0; k < ng; k++) {
        const double c = cos(gyaw[k]), s = sin(gyaw[k]);
        const double dx = gx[k] - cx, dy = gy[k] - cy, dz = gz[k] - cz;
        Ixx += c * c * gI[k][0] + s * s * gI[k][1] + gm[k] * (dy * dy + dz * dz);
        Iyy += s * s * gI[k][0] + c * c * gI[k][1] + gm[k] * (dx * dx + dz * dz);
        Izz += gI[k][2] + gm[k] * (dx * dx + dy * dy);
    }
    out[C_MB] = M; out[C_CZ] = cz; out[C_IBX] = Ixx; out[C_IBY] = Iyy; out[C_IBZ] = Izz;
    // --- pendulum body: pole cylinder + cubic weight                                        env_gen.py:69-72
    const bool pend = pendulum_enabled && plen > 0 && wmass > 0;
    if (pend) {
        const double mp = RND(0.2 * plen), hh = RND(plen / 2), zp = RND(-plen / 2), rp = 0.005;
        const double mw = RND(wmass), sw = RND(0.1 * cbrt(wmass)), zw = RND(-plen);
        const double mD = mp + mw, zD = (mp * zp + mw * zw) / mD;
        const double Ip_t = mp * (3 * rp * rp + 4 * hh * hh) / 12, Ip_a = mp * rp * rp / 2, Iw = mw * (2 * sw * sw) / 3;
        out[C_MD] = mD; out[C_ZD] = zD;
        out[C_IDX] = Ip_t + mp * (zp - zD) * (zp - zD) + Iw + mw * (zw - zD) * (zw - zD);
        out[C_IDZ] = Ip_a + Iw;
    } else {
        out[C_MD] = 0; out[C_ZD] = 0; out[C_IDX] = 0; out[C_IDZ] = 0;
    }
    // --- actuators: gear (0,0,F,0,0,+-F/100), dyntype filter with dynprm[0] = tau            env_gen.py:62-64
    const double F = RND(force), kq = RND(force / 100);
    out[C_F] = F; out[C_FS] = F * site_x; out[C_KQ] = kq;
    const double t = RND(tau);
    out[C_INVTAU] = 1.0 / (t > kMinVal ? t : kMinVal);
#undef RND
}

}  // namespace dsim
