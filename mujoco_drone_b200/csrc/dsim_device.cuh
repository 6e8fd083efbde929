// dsim_device.cuh — device-side building blocks of the fused env-step kernel (sm_100a).
//
// Everything is templated on the scalar type T: float is the product path, double is the validation
// instantiation (same source, so a derivation error shows at 1e-12 against the CPU oracle instead of hiding
// inside FP32 round-off) and the optional FP64 mode (the reference computes in FP64).
//
// The dynamics are a HAND-SPECIALISED derivation for the fixed kinematic tree of environments/env_gen.py:7-73
// (free-floating body B + link C on an x-hinge + pendulum D on a y-hinge), written in the BODY frame:
// Newton-Euler with the composite-body 6x6 block reduced analytically to one 3x3 SPD solve and the two hinge
// DOFs eliminated last (Schur complement), so the explicit acceleration (accelerometer, mj_sensorAcc) and the
// implicit-joint-damping acceleration (mj_EulerSkip) share one factorisation.  It is NOT a transcription of
// MuJoCo's generic world-frame CRB/RNE (that is what oracle/dsim_oracle.c restates); agreement between the two
// is the derivation check.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

#define DSIM_DEV __device__ __forceinline__
#define DSIM_HD __host__ __device__ __forceinline__

namespace dsim {

// ------------------------------------------------------------------ model-wide constants (env_gen.py)
constexpr double kGravity = 9.81;        // MuJoCo default gravity (0,0,-9.81)
constexpr double kRho = 1.2;             // env_gen.py:83 density
constexpr double kEta = 0.00002;         // env_gen.py:84 viscosity
constexpr double kHingeDamping = 0.15;   // env_gen.py:23 default joint damping (hinges only; <freejoint> takes no defaults)
constexpr double kLinkDrop = 0.025;      // env_gen.py:66 link body at (0,0,-half_body_size/2)
constexpr double kSenseZ = -0.0125;      // env_gen.py:48 'sense' site at (0,0,-half_body_size/4)
constexpr double kMassC = 0.01;          // env_gen.py:68 sphere mass
constexpr double kInertiaC = 0.4 * 0.01 * 0.02 * 0.02;   // solid sphere r = 0.02
constexpr double kPi = 3.14159265358979323846;
constexpr double kMinVal = 1e-15;        // mjMINVAL

// rows of the per-env state (include/dronesim_b200.h): qpos, qvel, act = 21 rows the step both reads and writes
enum { S_POS = 0, S_QUAT = 3, S_HINGE = 7, S_VEL = 9, S_OMEGA = 12, S_HVEL = 15, S_ACT = 17, S_ROWS = 21 };
// rows of the compiled constants
enum { C_MB = 0, C_CZ, C_IBX, C_IBY, C_IBZ, C_MD, C_ZD, C_IDX, C_IDZ, C_FS, C_F, C_KQ, C_INVTAU, C_ROWS };

// ------------------------------------------------------------------ paged env memory (include/dronesim_b200.h)
// Env data lives in PAGES of kTile = 32 envs (one warp): page p of a buffer with R rows is the contiguous block
// [R][32] at base + p * R * 32, element (row r, env i) at (i / 32) * R * 32 + r * 32 + (i % 32).  One page is one
// contiguous 1D bulk-copy (TMA, cp.async.bulk) between HBM and a warp's shared-memory slot; inside the slot every
// row is a conflict-free 128-byte (FP32) line and each row is reached with an immediate offset from one base.
constexpr int kTile = 32;
// read-write page: MuJoCo state (qpos, qvel, act) + BaseDroneEnv.num_steps + running episode return + the env's reset
// count (Philox epoch of its reset stream: rides with the page so the reset path has no global round trip) = the
// RW_IN_ROWS rows a step LOADS; the sensordata rows (accelerometer) come last: a step overwrites them (mj_sensorAcc runs
// in every substep) and never reads them, so they are stored with the page but not fetched (12 of 108 bytes per env)
enum { RW_NUM_STEPS = S_ROWS, RW_EP_RETURN = S_ROWS + 1, RW_RESET_COUNT = S_ROWS + 2, RW_IN_ROWS = S_ROWS + 3, S_ACC = RW_IN_ROWS, RW_ROWS = RW_IN_ROWS + 3 };
// read-only page: compiled rigid-body constants + raw drone_params (rewritten only by regen / set_params)
enum { RO_CONSTS = 0, RO_PARAMS = C_ROWS, RO_ROWS = C_ROWS + 6 };
enum { REF_ROWS = 4 };                        // per-env setpoint page (x, y, z offsets from start_pos, yaw)
DSIM_HD size_t page_elem(int rows, int r, int i) { return (size_t)(i >> 5) * (size_t)(rows * kTile) + (size_t)(r * kTile + (i & 31)); }

// ------------------------------------------------------------------ scalar helpers
// FP32 product path: single-MUFU approximations (rcp/sqrt/rsqrt .approx.ftz, <= 1-2 ulp) instead of the IEEE sequences
// with their slow-path branches; the FP64 validation build keeps exact libm semantics.
template <typename T> DSIM_DEV T sqrt_(T x) {
    if constexpr (std::is_same<T, float>::value) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; } else return sqrt(x);
}
template <typename T> DSIM_DEV T rcp_(T x) {
    if constexpr (std::is_same<T, float>::value) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; } else return 1.0 / x;
}
template <typename T> DSIM_DEV T floor_(T x) { if constexpr (std::is_same<T, float>::value) return floorf(x); else return floor(x); }
template <typename T> DSIM_DEV T rsqrt_(T x) {
    if constexpr (std::is_same<T, float>::value) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; } else return 1.0 / sqrt(x);
}
template <typename T> DSIM_DEV T abs_(T x) { if constexpr (std::is_same<T, float>::value) return fabsf(x); else return fabs(x); }
template <typename T> DSIM_DEV T max_(T a, T b) { if constexpr (std::is_same<T, float>::value) return fmaxf(a, b); else return fmax(a, b); }
template <typename T> DSIM_DEV T min_(T a, T b) { if constexpr (std::is_same<T, float>::value) return fminf(a, b); else return fmin(a, b); }
// atan2: branch-free for FP32.  atan(t), t = min/max in [0,1], as t * P(t^2) (degree-8 least-squares Chebyshev fit,
// max abs error 1.1e-7), then the octant / quadrant / sign fix-ups as selects.  atan2(0, 0) = 0 like libm.
DSIM_DEV float atan_unit(float t, float s2) {     // atan(t) for t in [0, 1], s2 = t * t
    float r = 2.8340642988e-03f;
    r = fmaf(r, s2, -1.6005030503e-02f); r = fmaf(r, s2, 4.2587607465e-02f); r = fmaf(r, s2, -7.4954454434e-02f);
    r = fmaf(r, s2, 1.0636754098e-01f); r = fmaf(r, s2, -1.4202570512e-01f); r = fmaf(r, s2, 1.9992483579e-01f);
    r = fmaf(r, s2, -3.3333066781e-01f); r = fmaf(r, s2, 9.9999998424e-01f);
    return r * t;
}
template <typename T> DSIM_DEV T atan2_(T y, T x) {
    if constexpr (std::is_same<T, float>::value) {
        const float ax = fabsf(x), ay = fabsf(y);
        const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
        const float t = mn * rcp_(fmaxf(mx, 1e-37f));
        float r = atan_unit(t, t * t);
        r = ay > ax ? 1.57079632679f - r : r;
        r = x < 0.f ? 3.14159265359f - r : r;
        return copysignf(r, y);
    } else return atan2(y, x);
}
// atan2(sqrt(y2), sqrt(x2)) for y2, x2 >= 0 (first quadrant): one square root of the ratio, no sign fix-ups
template <typename T> DSIM_DEV T atan2_sqrt(T y2, T x2) {
    if constexpr (std::is_same<T, float>::value) {
        const float s2 = fminf(x2, y2) * rcp_(fmaxf(fmaxf(x2, y2), 1e-37f));
        const float r = atan_unit(sqrt_(s2), s2);
        return y2 > x2 ? 1.57079632679f - r : r;
    } else return atan2(sqrt(y2), sqrt(x2));
}
template <typename T> DSIM_DEV T fmod_(T a, T b) { if constexpr (std::is_same<T, float>::value) return fmodf(a, b); else return fmod(a, b); }
template <typename T> DSIM_DEV T log_(T x) { if constexpr (std::is_same<T, float>::value) return logf(x); else return log(x); }
template <typename T> DSIM_DEV T cbrt_(T x) { if constexpr (std::is_same<T, float>::value) return cbrtf(x); else return cbrt(x); }
template <typename T> DSIM_DEV void sincos_(T a, T *s, T *c) { if constexpr (std::is_same<T, float>::value) sincosf(a, s, c); else sincos(a, s, c); }
// hinge angles: FP32 sin and cos together, branch-free.  Quadrant by the round-to-nearest "magic add", three-constant
// Cody-Waite reduction to [-pi/4, pi/4], the classic degree-7 / degree-8 minimax polynomials, quadrant swap / signs as
// selects: max abs error 7e-8 on |a| <= 200 (checked against FP64; libm sincosf: 3e-8) in ~21 instructions instead of
// libm's ~35 plus its large-argument slow path.  (The two MUFU ops are cheaper still but 2^-21.4 absolute: a measured 3x
// larger FP32-vs-FP64 deviation of the pendulum-coupled rates.)  Valid for |a| < 1e5 rad; FP64 keeps libm.
template <typename T> DSIM_DEV void sincos_hinge(T a, T *s, T *c) {
    if constexpr (std::is_same<T, float>::value) {
        const float t = fmaf(a, 0.636619772f, 12582912.0f);
        const int q = __float_as_int(t);
        const float j = t - 12582912.0f;
        float r = fmaf(j, -1.57079601e+00f, a);
        r = fmaf(j, -3.13916473e-07f, r);
        r = fmaf(j, -5.39030253e-15f, r);
        const float r2 = r * r;
        float sp = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f); sp = fmaf(sp, r2, -1.6666654611e-1f);
        const float sn = fmaf(sp * r2, r, r);
        float cp = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f); cp = fmaf(cp, r2, 4.166664568298827e-2f); cp = fmaf(cp, r2, -0.5f);
        const float cs = fmaf(cp, r2, 1.0f);
        const float ss = (q & 1) ? cs : sn, cc = (q & 1) ? sn : cs;
        *s = __int_as_float(__float_as_int(ss) ^ ((q & 2) << 30));
        *c = __int_as_float(__float_as_int(cc) ^ (((q + 1) & 2) << 30));
    } else sincos(a, s, c);
}
template <typename T> DSIM_DEV bool finite_(T x) { return isfinite(x); }
template <typename T> DSIM_DEV T clamp_(T x, T lo, T hi) { return min_(max_(x, lo), hi); }
// (a + pi) % (2 pi) - pi with Python's sign convention (rewards.py / observation_wrappers.py / scipy as_euler):
// a - 2 pi floor((a + pi) / 2 pi); returns `a` itself when it already lies in [-pi, pi)
template <typename T> DSIM_DEV T wrap_pi(T a) {
    const T k = floor_((a + T(kPi)) * T(1.0 / (2 * kPi)));
    return a - k * T(2 * kPi);
}

template <typename T> struct V3 { T x, y, z; };
template <typename T> DSIM_DEV V3<T> mk(T x, T y, T z) { V3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <typename T> DSIM_DEV V3<T> operator+(V3<T> a, V3<T> b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> DSIM_DEV V3<T> operator-(V3<T> a, V3<T> b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> DSIM_DEV V3<T> operator*(T s, V3<T> a) { return mk(s * a.x, s * a.y, s * a.z); }
template <typename T> DSIM_DEV T dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> DSIM_DEV V3<T> cross(V3<T> a, V3<T> b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

// rotation matrix (row-major, body -> world) of a UNIT quaternion (w,x,y,z)
template <typename T> struct M3 { T m[9]; };
template <typename T> DSIM_DEV M3<T> quat_to_mat(T w, T x, T y, T z) {
    // unit quaternion: diagonal as 1 - 2 (.. + ..), products shared through the doubled components (24 operations)
    const T tx = x + x, ty = y + y, tz = z + z;
    const T xx = x * tx, yy = y * ty, zz = z * tz, xy = x * ty, xz = x * tz, yz = y * tz, wx = w * tx, wy = w * ty, wz = w * tz;
    M3<T> R;
    R.m[0] = T(1) - (yy + zz); R.m[1] = xy - wz;          R.m[2] = xz + wy;
    R.m[3] = xy + wz;          R.m[4] = T(1) - (xx + zz); R.m[5] = yz - wx;
    R.m[6] = xz - wy;          R.m[7] = yz + wx;          R.m[8] = T(1) - (xx + yy);
    return R;
}
template <typename T> DSIM_DEV V3<T> mul(const M3<T> &R, V3<T> v) {
    return mk(R.m[0] * v.x + R.m[1] * v.y + R.m[2] * v.z, R.m[3] * v.x + R.m[4] * v.y + R.m[5] * v.z, R.m[6] * v.x + R.m[7] * v.y + R.m[8] * v.z);
}
template <typename T> DSIM_DEV V3<T> tmul(const M3<T> &R, V3<T> v) {
    return mk(R.m[0] * v.x + R.m[3] * v.y + R.m[6] * v.z, R.m[1] * v.x + R.m[4] * v.y + R.m[7] * v.z, R.m[2] * v.x + R.m[5] * v.y + R.m[8] * v.z);
}
// R = Rz(yaw) Ry(pitch) Rx(roll): transformation.py:20-23 + :10-12
template <typename T> DSIM_DEV M3<T> rpy_to_mat(T roll, T pitch, T yaw) {
    T sr, cr, sp, cp, sy, cy;
    sincos_(roll, &sr, &cr); sincos_(pitch, &sp, &cp); sincos_(yaw, &sy, &cy);
    M3<T> R;
    R.m[0] = cy * cp; R.m[1] = cy * sp * sr - sy * cr; R.m[2] = cy * sp * cr + sy * sr;
    R.m[3] = sy * cp; R.m[4] = sy * sp * sr + cy * cr; R.m[5] = sy * sp * cr - cy * sr;
    R.m[6] = -sp;     R.m[7] = cp * sr;                R.m[8] = cp * cr;
    return R;
}

}  // namespace dsim
#include "dsim_packed.cuh"     // F2: two FP32 values per register pair on the packed FP32 pipe (needs the scalar helpers and V3 above)
namespace dsim {

// ------------------------------------------------------------------ per-env state held in registers
template <typename T> struct EnvState {
    V3<T> pos;            // OFFSET from start_pos
    T qw, qx, qy, qz;
    T hx, hy;             // hinge angles
    V3<T> vel;            // world frame
    V3<T> om;             // body frame
    T hvx, hvy;           // hinge rates
    T act[4];
    V3<T> acc;            // accelerometer
};
template <typename T> struct EnvConsts { T mB, cz, IBx, IBy, IBz, mD, zD, IDx, IDz, Fs, F, kq, inv_tau; };

// mj_inertiaBoxFluidModel for one body: local angular/linear velocity at the COM in the principal frame ->
// local force / torque.  Box from (mass, principal inertia).
template <typename T> struct FluidBox { T fq[3], tq[3], kv, kw; };
template <typename T> DSIM_DEV FluidBox<T> fluid_box(T mass, T Ix, T Iy, T Iz) {
    T s = T(6) * rcp_(mass);
    T bx = sqrt_(max_(T(kMinVal), Iy + Iz - Ix) * s), by = sqrt_(max_(T(kMinVal), Ix + Iz - Iy) * s), bz = sqrt_(max_(T(kMinVal), Ix + Iy - Iz) * s);
    T d = (bx + by + bz) * T(1.0 / 3.0);
    FluidBox<T> f;
    f.kv = T(3.0 * kPi * kEta) * d;
    f.kw = T(kPi * kEta) * d * d * d;
    f.fq[0] = T(0.5 * kRho) * by * bz; f.fq[1] = T(0.5 * kRho) * bx * bz; f.fq[2] = T(0.5 * kRho) * bx * by;
    T bx2 = bx * bx, by2 = by * by, bz2 = bz * bz;
    T bx4 = bx2 * bx2, by4 = by2 * by2, bz4 = bz2 * bz2;
    f.tq[0] = T(kRho / 64.0) * bx * (by4 + bz4); f.tq[1] = T(kRho / 64.0) * by * (bx4 + bz4); f.tq[2] = T(kRho / 64.0) * bz * (bx4 + by4);
    return f;
}
template <typename T> DSIM_DEV void fluid_apply(const FluidBox<T> &b, V3<T> w, V3<T> v, V3<T> &f, V3<T> &t) {
    // viscous + quadratic drag per axis: -k u - q |u| u = -u (k + q |u|): one FMA and one multiply per component
    f = mk(-v.x * (b.kv + b.fq[0] * abs_(v.x)), -v.y * (b.kv + b.fq[1] * abs_(v.y)), -v.z * (b.kv + b.fq[2] * abs_(v.z)));
    t = mk(-w.x * (b.kw + b.tq[0] * abs_(w.x)), -w.y * (b.kw + b.tq[1] * abs_(w.y)), -w.z * (b.kw + b.tq[2] * abs_(w.z)));
}

// symmetric 3x3 SPD solve through LDL^T (factor once, three right-hand sides per substep)
template <typename T> struct Ldl3 { T l10, l20, l21, id0, id1, id2; };
template <typename T> DSIM_DEV Ldl3<T> ldl3(T a00, T a10, T a11, T a20, T a21, T a22) {
    Ldl3<T> f;
    f.id0 = rcp_(a00);
    f.l10 = a10 * f.id0; f.l20 = a20 * f.id0;
    T d1 = a11 - f.l10 * a10;
    f.id1 = rcp_(d1);
    f.l21 = (a21 - f.l20 * a10) * f.id1;
    T d2 = a22 - f.l20 * a20 - f.l21 * f.l21 * d1;
    f.id2 = rcp_(d2);
    return f;
}
template <typename T> DSIM_DEV V3<T> ldl3_solve(const Ldl3<T> &f, V3<T> b) {
    T y0 = b.x, y1 = b.y - f.l10 * y0, y2 = b.z - f.l20 * y0 - f.l21 * y1;
    T z2 = y2 * f.id2;
    T z1 = y1 * f.id1 - f.l21 * z2;
    T z0 = y0 * f.id0 - f.l10 * z1 - f.l20 * z2;
    return mk(z0, z1, z2);
}

// mju_normalize4 of the free-joint quaternion: a (near-)zero quaternion becomes identity
template <typename T> DSIM_DEV void quat_unit(T sw, T sx, T sy, T sz, T &qw, T &qx, T &qy, T &qz) {
    const T qn = sw * sw + sx * sx + sy * sy + sz * sz;
    const bool qok = qn >= T(1e-30);
    const T qi = rsqrt_(qn);
    qw = qok ? sw * qi : T(1); qx = qok ? sx * qi : T(0); qy = qok ? sy * qi : T(0); qz = qok ? sz * qi : T(0);
}
// mju_quatIntegrate: the rotation quaternion of one step, [rw, kq * w] = [cos(a), sin(a) / |w| * w] with a = h |w| / 2 and
// w2 = |w|^2.  FP32: even Taylor series in z = a^2 (no square root, no division, exact identity for w = 0 like MuJoCo's
// |w| < mjMINVAL branch); relative error < 1e-7 for a <= 0.8, i.e. |w| <= 160 rad/s at 100 Hz; polynomial sincos beyond.
template <typename T> DSIM_DEV void quat_step(T h, T w2, T &rw, T &kq) {
    rw = T(1); kq = T(0);
    if constexpr (std::is_same<T, float>::value) {
        const float z = (0.25f * h * h) * w2;
        if (z <= 0.64f) {
            float ps = fmaf(z, 2.7557319e-6f, -1.9841270e-4f); ps = fmaf(ps, z, 8.3333333e-3f); ps = fmaf(ps, z, -1.6666667e-1f);
            kq = (0.5f * h) * fmaf(ps, z, 1.0f);
            float pc = fmaf(z, -2.7557319e-7f, 2.4801587e-5f); pc = fmaf(pc, z, -1.3888889e-3f); pc = fmaf(pc, z, 4.1666667e-2f); pc = fmaf(pc, z, -0.5f);
            rw = fmaf(pc, z, 1.0f);
        } else {
            const float iw = rsqrt_(w2), wn = w2 * iw;
            float sh;
            sincos_hinge(0.5f * h * wn, &sh, &rw);                          // branch-free polynomial (7e-8): no libm slow path in the kernel image
            kq = sh * iw;
        }
    } else if (w2 >= T(1e-30)) {
        const T wn = sqrt_(w2);
        T sh;
        sincos_(T(0.5) * h * wn, &sh, &rw);
        kq = sh / wn;
    }
}

// free-joint quaternion: normalisation and the per-step rotation, packed where it is plain arithmetic
DSIM_DEV void quat_unit(F2 sw, F2 sx, F2 sy, F2 sz, F2 &qw, F2 &qx, F2 &qy, F2 &qz) {
    const F2 qn = sw * sw + sx * sx + sy * sy + sz * sz;
    const F2 qi = rsqrt_(qn);
    qw = sw * qi; qx = sx * qi; qy = sy * qi; qz = sz * qi;
    const bool ok0 = qn.lo() >= 1e-30f, ok1 = qn.hi() >= 1e-30f;
    if (!(ok0 && ok1)) {                                            // (near-)zero quaternion -> identity, per half
        qw = F2(ok0 ? qw.lo() : 1.f, ok1 ? qw.hi() : 1.f); qx = F2(ok0 ? qx.lo() : 0.f, ok1 ? qx.hi() : 0.f);
        qy = F2(ok0 ? qy.lo() : 0.f, ok1 ? qy.hi() : 0.f); qz = F2(ok0 ? qz.lo() : 0.f, ok1 ? qz.hi() : 0.f);
    }
}
DSIM_DEV void quat_step(F2 h, F2 w2, F2 &rw, F2 &kq) {
    const F2 z = (F2(0.25f) * h * h) * w2;
    if (z.lo() <= 0.64f && z.hi() <= 0.64f) {                       // both envs in the Taylor range (|w| <= 160 rad/s at 100 Hz)
        F2 ps = fma2(z, F2(2.7557319e-6f), F2(-1.9841270e-4f)); ps = fma2(ps, z, F2(8.3333333e-3f)); ps = fma2(ps, z, F2(-1.6666667e-1f));
        kq = (F2(0.5f) * h) * fma2(ps, z, F2(1.0f));
        F2 pc = fma2(z, F2(-2.7557319e-7f), F2(2.4801587e-5f)); pc = fma2(pc, z, F2(-1.3888889e-3f)); pc = fma2(pc, z, F2(4.1666667e-2f)); pc = fma2(pc, z, F2(-0.5f));
        rw = fma2(pc, z, F2(1.0f));
    } else {
        float r0, k0, r1, k1;
        quat_step(h.lo(), w2.lo(), r0, k0); quat_step(h.hi(), w2.hi(), r1, k1);
        rw = F2(r0, r1); kq = F2(k0, k1);
    }
}


// ------------------------------------------------------------------ one mj_step (Euler, implicit hinge damping)
// ADVANCE=false evaluates only the forward part (mj_forward: accelerometer refresh after set_state).
// GROUND: floor contacts (dsim_contact.cuh) for drones whose bounding sphere reaches the floor; `g` is only read then.
#ifndef DSIM_CONTACT_CALL
#define DSIM_CONTACT_CALL __noinline__
#endif
template <typename T> struct GroundCtx;
template <typename T> struct ContactIO;
template <typename T> __device__ DSIM_CONTACT_CALL int contact_solve(ContactIO<T> &io, const EnvConsts<T> &c, const GroundCtx<T> &g);
template <typename T, bool PEND, bool ADVANCE, bool GROUND = false>
DSIM_DEV void substep(EnvState<T> &s, const EnvConsts<T> &c, const T ctrl[4], T h, const GroundCtx<T> *g = nullptr) {
    // -- kinematics (mj_kinematics normalises the free-joint quaternion)
    T qw, qx, qy, qz;
    quat_unit(s.qw, s.qx, s.qy, s.qz, qw, qx, qy, qz);
    const M3<T> R = quat_to_mat(qw, qx, qy, qz);
    const V3<T> vb = tmul(R, s.vel);                                         // origin velocity, body coords
    const V3<T> gb = mk(T(-kGravity) * R.m[6], T(-kGravity) * R.m[7], T(-kGravity) * R.m[8]);   // R^T g
    const V3<T> om = s.om;
    const T mC = PEND ? T(kMassC) : T(0), IC = PEND ? T(kInertiaC) : T(0), dl = T(kLinkDrop);
    const T mD = PEND ? c.mD : T(0);
    const T mh = mC + mD, mtot = c.mB + mh, inv_m = rcp_(mtot);

    T sx = 0, cx = 1, sy = 0, cy = 1;
    if (PEND) {
        if constexpr (std::is_same<T, float>::value) {
            // the two hinge angles through ONE packed evaluation (FP32 scalar issues every other cycle; FFMA2 carries both)
            F2 S, C;
            sincos_hinge(F2(s.hx, s.hy), &S, &C);
            sx = S.lo(); sy = S.hi(); cx = C.lo(); cy = C.hi();
        } else { sincos_hinge(s.hx, &sx, &cx); sincos_hinge(s.hy, &sy, &cy); }
    }
    const V3<T> yc = mk(T(0), cx, sx);                                       // hinge-y axis (C frame y) in body coords
    const V3<T> n = mk(sy, -sx * cy, cx * cy);                               // pendulum axis (D frame z)
    const V3<T> xd = mk(cy, sx * sy, -cx * sy);                              // D frame x
    const T mu = mD * c.zD;                                                  // first moment of D about the hinge point
    const T P = PEND ? c.IDx + mD * c.zD * c.zD : T(0);                      // transverse inertia of D about the hinge
    const T QmP = PEND ? c.IDz - P : T(0);

    // -- composite body about the origin o of the free joint: first moment H, inertia Io, then about the composite COM
    const V3<T> H = mk(mu * n.x, mu * n.y, c.mB * c.cz - mh * dl + mu * n.z);
    const T k12 = c.mB * c.cz * c.cz + IC + mh * dl * dl + P - T(2) * mu * dl * n.z;
    const T Ixx = c.IBx + k12 + QmP * n.x * n.x, Iyy = c.IBy + k12 + QmP * n.y * n.y, Izz = c.IBz + IC + P + QmP * n.z * n.z;
    const T Ixy = QmP * n.x * n.y, Ixz = QmP * n.x * n.z + mu * dl * n.x, Iyz = QmP * n.y * n.z + mu * dl * n.y;
    const Ldl3<T> L = ldl3(Ixx - (H.y * H.y + H.z * H.z) * inv_m, Ixy + H.x * H.y * inv_m, Iyy - (H.x * H.x + H.z * H.z) * inv_m,
                           Ixz + H.x * H.z * inv_m, Iyz + H.y * H.z * inv_m, Izz - (H.x * H.x + H.y * H.y) * inv_m);
    const V3<T> com = inv_m * H;

    // -- velocity products (RNE with zero acceleration) and gravity
    const T oxy2 = om.x * om.x + om.y * om.y;
    const V3<T> cen = mk(om.x * om.z, om.y * om.z, -oxy2);                   // w x (w x z^) : scale by the z offset
    // body B: COM at (0,0,cz)
    const V3<T> FB0 = c.mB * (c.cz * cen - gb);
    const V3<T> IBw = mk(c.IBx * om.x, c.IBy * om.y, c.IBz * om.z);
    V3<T> blin = FB0;
    V3<T> bang = cross(om, IBw) + mk(-c.cz * FB0.y, c.cz * FB0.x, T(0));     // N_B + r_B x F_B
    T bias_x = 0, bias_y = 0;
    V3<T> rho = mk(T(0), T(0), T(0)), omD = om, omC = om;
    if (PEND) {
        const V3<T> ah0 = (-dl) * cen;                                       // hinge point at (0,0,-dl)
        omC = mk(om.x + s.hvx, om.y, om.z);
        omD = omC + s.hvy * yc;
        const V3<T> aC0 = s.hvx * mk(T(0), om.z, -om.y);                     // hvx * (w x x^)
        const V3<T> aD0 = aC0 + s.hvy * cross(omC, yc);
        rho = c.zD * n;
        const V3<T> accD = ah0 + cross(aD0, rho) + cross(omD, cross(omD, rho));
        const V3<T> FC0 = mC * (ah0 - gb);
        const V3<T> FD0 = mD * (accD - gb);
        const T dI = c.IDz - c.IDx;
        const V3<T> IDa = c.IDx * aD0 + (dI * dot(n, aD0)) * n;
        const V3<T> IDw = c.IDx * omD + (dI * dot(n, omD)) * n;
        const V3<T> TD = IDa + cross(omD, IDw) + cross(rho, FD0);            // about the hinge point
        const V3<T> NC0 = IC * aC0;
        const V3<T> Fh = FC0 + FD0;
        blin = blin + Fh;
        bang = bang + NC0 + TD + mk(dl * Fh.y, -dl * Fh.x, T(0));            // r_h x F, r_h = (0,0,-dl)
        bias_x = NC0.x + TD.x;
        bias_y = dot(yc, TD);
    }

    // -- applied forces: actuators (site transmission, force = act), fluid, hinge damping
    const T a0 = s.act[0], a1 = s.act[1], a2 = s.act[2], a3 = s.act[3];
    V3<T> flin = mk(T(0), T(0), c.F * ((a0 + a1) + (a2 + a3)));
    V3<T> fang = mk(c.Fs * ((a1 + a2) - (a0 + a3)), c.Fs * ((a2 + a3) - (a0 + a1)), c.kq * ((a0 + a2) - (a1 + a3)));
    T f_x = 0, f_y = 0;
    // bodies B and D run the same inertia-box arithmetic on different numbers: FP32 with a pendulum evaluates both in one packed
    // pass (B | D) further down
    constexpr bool kFluidPair = PEND && std::is_same<T, float>::value;
    if constexpr (!kFluidPair) {   // body B: principal frame == body frame (4-fold symmetry)
        const FluidBox<T> fb = fluid_box(c.mB, c.IBx, c.IBy, c.IBz);
        V3<T> f, t;
        fluid_apply(fb, om, vb + mk(om.y * c.cz, -om.x * c.cz, T(0)), f, t);
        flin = flin + f;
        fang = fang + t + mk(-c.cz * f.y, c.cz * f.x, T(0));
    }
    if (PEND) {
        const V3<T> vh = vb + mk(-om.y * dl, om.x * dl, T(0));               // w x r_h
        {   // body C: sphere -> cubic box; frame = Rx(hx)
            constexpr double bC = 0.030983866769659335;                      // sqrt(6 I_C / m_C)
            FluidBox<T> fb;
            fb.kv = T(3.0 * kPi * kEta * bC); fb.kw = T(kPi * kEta * bC * bC * bC);
            fb.fq[0] = fb.fq[1] = fb.fq[2] = T(0.5 * kRho * bC * bC);
            fb.tq[0] = fb.tq[1] = fb.tq[2] = T(kRho / 64.0 * bC * 2.0 * bC * bC * bC * bC);
            const V3<T> vl = mk(vh.x, cx * vh.y + sx * vh.z, -sx * vh.y + cx * vh.z);
            const V3<T> wl = mk(omC.x, cx * omC.y + sx * omC.z, -sx * omC.y + cx * omC.z);
            V3<T> fl, tl;
            fluid_apply(fb, wl, vl, fl, tl);
            const V3<T> f = mk(fl.x, cx * fl.y - sx * fl.z, sx * fl.y + cx * fl.z);
            const V3<T> t = mk(tl.x, cx * tl.y - sx * tl.z, sx * tl.y + cx * tl.z);
            flin = flin + f;
            fang = fang + t + mk(dl * f.y, -dl * f.x, T(0));
            f_x += t.x;
        }
        {   // body D: frame axes (xd, yc, n), COM at hinge + rho
            const V3<T> vD = vh + cross(omD, rho);
            V3<T> fl, tl;
            if constexpr (kFluidPair) {
                const FluidBox<F2> fb = fluid_box(F2(c.mB, mD), F2(c.IBx, c.IDx), F2(c.IBy, c.IDx), F2(c.IBz, c.IDz));
                const V3<T> vB = vb + mk(om.y * c.cz, -om.x * c.cz, T(0));
                V3<F2> f2, t2;
                fluid_apply(fb, mk(F2(om.x, dot(xd, omD)), F2(om.y, dot(yc, omD)), F2(om.z, dot(n, omD))),
                            mk(F2(vB.x, dot(xd, vD)), F2(vB.y, dot(yc, vD)), F2(vB.z, dot(n, vD))), f2, t2);
                const V3<T> fB = mk(f2.x.lo(), f2.y.lo(), f2.z.lo()), tB = mk(t2.x.lo(), t2.y.lo(), t2.z.lo());
                flin = flin + fB;
                fang = fang + tB + mk(-c.cz * fB.y, c.cz * fB.x, T(0));
                fl = mk(f2.x.hi(), f2.y.hi(), f2.z.hi()); tl = mk(t2.x.hi(), t2.y.hi(), t2.z.hi());
            } else {
                const FluidBox<T> fb = fluid_box(mD, c.IDx, c.IDx, c.IDz);
                fluid_apply(fb, mk(dot(xd, omD), dot(yc, omD), dot(n, omD)), mk(dot(xd, vD), dot(yc, vD), dot(n, vD)), fl, tl);
            }
            const V3<T> f = fl.x * xd + fl.y * yc + fl.z * n;
            const V3<T> W = tl.x * xd + tl.y * yc + tl.z * n + cross(rho, f);
            flin = flin + f;
            fang = fang + W + mk(dl * f.y, -dl * f.x, T(0));
            f_x += W.x;
            f_y += dot(yc, W);
        }
        f_x -= T(kHingeDamping) * s.hvx;
        f_y -= T(kHingeDamping) * s.hvy;
    }

    // -- forward acceleration: base block via the COM-reduced 3x3, hinges via the Schur complement
    const V3<T> F = flin - blin, Tq = fang - bang;
    const V3<T> al0 = ldl3_solve(L, Tq - cross(com, F));
    const V3<T> ac0 = inv_m * F + cross(com, al0);
    V3<T> a_e = ac0, al_e = al0, a_i = ac0, al_i = al0;
    T hax_i = 0, hay_i = 0;
    T hax_e = 0, hay_e = 0;
    const T rx = f_x - bias_x, ry = f_y - bias_y;
    if (PEND) {
        const V3<T> px = (-mu * cy) * yc, py = mu * xd;                      // linear momentum per unit hinge rate
        const V3<T> Lx = mk(IC + P + QmP * n.x * n.x, QmP * n.x * n.y, QmP * n.x * n.z) + mk(dl * px.y, -dl * px.x, T(0));
        const V3<T> Ly = P * yc + mk(dl * py.y, -dl * py.x, T(0));
        V3<T> alx, acx, aly, acy;
        T Sxx, Syy, r1, r2;
        if constexpr (std::is_same<T, float>::value) {
            // the hinge-x and hinge-y columns are the same arithmetic on different right-hand sides: one packed pass (x | y)
            const V3<F2> pp = mk(F2(px.x, py.x), F2(px.y, py.y), F2(px.z, py.z)), LL = mk(F2(Lx.x, Ly.x), F2(Lx.y, Ly.y), F2(Lx.z, Ly.z));
            const V3<F2> com2 = mk(F2(com.x), F2(com.y), F2(com.z));
            Ldl3<F2> L2;
            L2.l10 = F2(L.l10); L2.l20 = F2(L.l20); L2.l21 = F2(L.l21); L2.id0 = F2(L.id0); L2.id1 = F2(L.id1); L2.id2 = F2(L.id2);
            const V3<F2> al = ldl3_solve(L2, LL - cross(com2, pp));
            const V3<F2> ac = F2(inv_m) * pp + cross(com2, al);
            alx = mk(al.x.lo(), al.y.lo(), al.z.lo()); aly = mk(al.x.hi(), al.y.hi(), al.z.hi());
            acx = mk(ac.x.lo(), ac.y.lo(), ac.z.lo()); acy = mk(ac.x.hi(), ac.y.hi(), ac.z.hi());
            const F2 dg = dot(pp, ac) + dot(LL, al);                         // (px.acx + Lx.alx | py.acy + Ly.aly)
            Sxx = (IC + P + QmP * n.x * n.x) - dg.lo(); Syy = P - dg.hi();
            const F2 rr = F2(rx, ry) - (dot(pp, mk(F2(ac0.x), F2(ac0.y), F2(ac0.z))) + dot(LL, mk(F2(al0.x), F2(al0.y), F2(al0.z))));
            r1 = rr.lo(); r2 = rr.hi();
        } else {
            alx = ldl3_solve(L, Lx - cross(com, px)); acx = inv_m * px + cross(com, alx);
            aly = ldl3_solve(L, Ly - cross(com, py)); acy = inv_m * py + cross(com, aly);
            Sxx = (IC + P + QmP * n.x * n.x) - (dot(px, acx) + dot(Lx, alx));
            Syy = P - (dot(py, acy) + dot(Ly, aly));
            r1 = rx - (dot(px, ac0) + dot(Lx, al0)); r2 = ry - (dot(py, ac0) + dot(Ly, al0));
        }
        const T Sxy = -(dot(px, acy) + dot(Lx, aly));
        if constexpr (ADVANCE && std::is_same<T, float>::value) {
            // the explicit solution (accelerometer) and the implicit-damping one (integration) differ only in the diagonal of the
            // 2 x 2 Schur block: one packed pass (explicit | implicit)
            const T hb = h * T(kHingeDamping);
            const F2 Sx2(Sxx, Sxx + hb), Sy2(Syy, Syy + hb), Sxy2(Sxy), r12(r1), r22(r2);
            const F2 idet = rcp_(Sx2 * Sy2 - Sxy2 * Sxy2);
            const F2 hax = F2(Sy2 * r12 - Sxy2 * r22) * idet, hay = F2(Sx2 * r22 - Sxy2 * r12) * idet;
            const V3<F2> acx2 = mk(F2(acx.x), F2(acx.y), F2(acx.z)), acy2 = mk(F2(acy.x), F2(acy.y), F2(acy.z));
            const V3<F2> alx2 = mk(F2(alx.x), F2(alx.y), F2(alx.z)), aly2 = mk(F2(aly.x), F2(aly.y), F2(aly.z));
            const V3<F2> a2 = mk(F2(ac0.x), F2(ac0.y), F2(ac0.z)) - hax * acx2 - hay * acy2;
            const V3<F2> l2 = mk(F2(al0.x), F2(al0.y), F2(al0.z)) - hax * alx2 - hay * aly2;
            a_e = mk(a2.x.lo(), a2.y.lo(), a2.z.lo()); a_i = mk(a2.x.hi(), a2.y.hi(), a2.z.hi());
            al_e = mk(l2.x.lo(), l2.y.lo(), l2.z.lo()); al_i = mk(l2.x.hi(), l2.y.hi(), l2.z.hi());
            hax_e = hax.lo(); hay_e = hay.lo(); hax_i = hax.hi(); hay_i = hay.hi();
        } else {
            {   // explicit (qacc of mj_fwdAcceleration): feeds the accelerometer
                const T idet = rcp_(Sxx * Syy - Sxy * Sxy);
                const T hax = (Syy * r1 - Sxy * r2) * idet, hay = (Sxx * r2 - Sxy * r1) * idet;
                a_e = ac0 - hax * acx - hay * acy;
                al_e = al0 - hax * alx - hay * aly;
                hax_e = hax; hay_e = hay;
            }
            if (ADVANCE) {  // implicit in joint damping (mj_EulerSkip): (M + h diag(B)) qacc = qfrc_smooth
                const T hb = h * T(kHingeDamping);
                const T Sxx2 = Sxx + hb, Syy2 = Syy + hb;
                const T idet = rcp_(Sxx2 * Syy2 - Sxy * Sxy);
                hax_i = (Syy2 * r1 - Sxy * r2) * idet; hay_i = (Sxx2 * r2 - Sxy * r1) * idet;
                a_i = ac0 - hax_i * acx - hay_i * acy;
                al_i = al0 - hax_i * alx - hay_i * aly;
            }
        }
    }

    if constexpr (GROUND) {
        // -- mj_collision + mj_fwdConstraint for the few drones that can touch the floor (slow path, not inlined)
        const T zo = g->start_z + s.pos.z;
        if (zo < g->reach) {
            ContactIO<T> io;
            io.q[0] = F.x; io.q[1] = F.y; io.q[2] = F.z; io.q[3] = Tq.x; io.q[4] = Tq.y; io.q[5] = Tq.z; io.q[6] = rx; io.q[7] = ry;
            io.v[0] = vb.x; io.v[1] = vb.y; io.v[2] = vb.z; io.v[3] = om.x; io.v[4] = om.y; io.v[5] = om.z; io.v[6] = PEND ? s.hvx : T(0); io.v[7] = PEND ? s.hvy : T(0);
            io.x[0] = a_e.x; io.x[1] = a_e.y; io.x[2] = a_e.z; io.x[3] = al_e.x; io.x[4] = al_e.y; io.x[5] = al_e.z; io.x[6] = hax_e; io.x[7] = hay_e;
            io.nb[0] = R.m[6]; io.nb[1] = R.m[7]; io.nb[2] = R.m[8];
            io.t1[0] = R.m[3]; io.t1[1] = R.m[4]; io.t1[2] = R.m[5];
            io.t2[0] = -R.m[0]; io.t2[1] = -R.m[1]; io.t2[2] = -R.m[2];
            io.zo = zo; io.sx = sx; io.cx = cx; io.sy = sy; io.cy = cy; io.h = h;
            if (contact_solve(io, c, *g) > 0) {
                a_e = mk(io.x[0], io.x[1], io.x[2]); al_e = mk(io.x[3], io.x[4], io.x[5]);
                a_i = mk(io.xi[0], io.xi[1], io.xi[2]); al_i = mk(io.xi[3], io.xi[4], io.xi[5]);
                hax_i = io.xi[6]; hay_i = io.xi[7];
            }
        }
    }

    // -- accelerometer at site 'sense' (0,0,zs), site frame == body frame; pre-integration state, explicit qacc
    const T zs = T(kSenseZ);
    s.acc = a_e - gb + mk(al_e.y * zs, -al_e.x * zs, T(0)) + zs * cen;

    if (ADVANCE) {
        // mj_advance: activation (dyntype=filter, explicit), velocity, then position with the NEW velocity
        #pragma unroll
        for (int k = 0; k < 4; k++) s.act[k] += h * ((ctrl[k] - s.act[k]) * c.inv_tau);
        s.vel = s.vel + h * mul(R, a_i);
        s.om = s.om + h * al_i;
        s.pos = s.pos + h * s.vel;
        // mju_quatIntegrate: q <- normalize(q) * axisangle(w/|w|, h|w|) = normalize(q) * [cos(a), sin(a)/|w| * w], a = h|w|/2.
        // FP32: even Taylor series in z = a^2 (no square root, no division, exact identity for w = 0 like MuJoCo's
        // |w| < mjMINVAL branch); relative error < 1e-7 for a <= 0.8, i.e. |w| <= 160 rad/s at 100 Hz; libm beyond.
        const T w2 = dot(s.om, s.om);
        T rw, kq;
        quat_step(h, w2, rw, kq);
        const T rx_ = kq * s.om.x, ry_ = kq * s.om.y, rz_ = kq * s.om.z;
        s.qw = qw * rw - qx * rx_ - qy * ry_ - qz * rz_;
        s.qx = qw * rx_ + qx * rw + qy * rz_ - qz * ry_;
        s.qy = qw * ry_ - qx * rz_ + qy * rw + qz * rx_;
        s.qz = qw * rz_ + qx * ry_ - qy * rx_ + qz * rw;
        if (PEND) {
            s.hvx += h * hax_i; s.hvy += h * hay_i;
            s.hx += h * s.hvx; s.hy += h * s.hvy;
        }
    }
}

// ------------------------------------------------------------------ transformation.py: quaternion -> (roll, pitch, yaw)
// scipy Rotation.as_euler('ZYX')[::-1] (quaternion algorithm, gimbal-lock branches, eps 1e-7)
template <typename T> DSIM_DEV void quat_to_rpy(T w, T x, T y, T z, T &roll, T &pitch, T &yaw) {
    const T a = w - y, b = x + z, c = y + w, d = z - x;
    T hs, hd;
    if constexpr (std::is_same<T, float>::value) {
        // the two half-angle arctangents through one packed evaluation: atan(t) = t P(t^2) on (hs | hd), fix-ups per half
        const float ax0 = fabsf(a), ay0 = fabsf(b), ax1 = fabsf(c), ay1 = fabsf(d);
        const F2 mx(fmaxf(ax0, ay0), fmaxf(ax1, ay1)), mn(fminf(ax0, ay0), fminf(ax1, ay1));
        const F2 t = mn * rcp_(max_(mx, F2(1e-37f))), s2 = t * t;
        F2 r(2.8340642988e-03f);
        r = fma2(r, s2, F2(-1.6005030503e-02f)); r = fma2(r, s2, F2(4.2587607465e-02f)); r = fma2(r, s2, F2(-7.4954454434e-02f));
        r = fma2(r, s2, F2(1.0636754098e-01f)); r = fma2(r, s2, F2(-1.4202570512e-01f)); r = fma2(r, s2, F2(1.9992483579e-01f));
        r = fma2(r, s2, F2(-3.3333066781e-01f)); r = fma2(r, s2, F2(9.9999998424e-01f));
        r = r * t;
        float r0 = r.lo(), r1 = r.hi();
        r0 = ay0 > ax0 ? 1.57079632679f - r0 : r0; r0 = a < 0.f ? 3.14159265359f - r0 : r0; hs = copysignf(r0, b);
        r1 = ay1 > ax1 ? 1.57079632679f - r1 : r1; r1 = c < 0.f ? 3.14159265359f - r1 : r1; hd = copysignf(r1, d);
    } else { hs = atan2_(b, a); hd = atan2_(d, c); }
    T a1 = T(2) * atan2_sqrt(c * c + d * d, a * a + b * b);
    const bool case1 = abs_(a1) <= T(1e-7), case2 = abs_(a1 - T(kPi)) <= T(1e-7);
    T a0, a2;
    if (!case1 && !case2) { a2 = hs - hd; a0 = hs + hd; }
    else { a2 = T(0); a0 = case1 ? T(2) * hs : T(2) * hd; }
    a1 -= T(kPi / 2);
    roll = wrap_pi(a2); pitch = wrap_pi(a1); yaw = wrap_pi(a0);
}

// ------------------------------------------------------------------ Philox4x32-10 and samplers
struct U4 { uint32_t x, y, z, w; };
DSIM_DEV U4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    #pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    U4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}
template <typename T> DSIM_DEV T u01(uint32_t x) { return (T(x >> 8) + T(0.5)) * T(1.0 / 16777216.0); }
// Samplers.  FP32: MUFU-based log2 / sin / cos / exp2 (arguments are already range-reduced: u in (0,1), angle in (0, 2 pi)),
// a few instructions each instead of libm's ~25 with slow-path branches; the draws are random numbers, ~1e-6 relative
// accuracy is irrelevant for them and stays inside the stated FP32 reset tolerance.  FP64 keeps libm.
template <typename T> DSIM_DEV void box_muller(uint32_t x0, uint32_t x1, T &z0, T &z1) {
    if constexpr (std::is_same<T, float>::value) {
        const float r = sqrt_(-1.3862943611f * __log2f(u01<float>(x0)));          // sqrt(-2 ln u) = sqrt(-2 ln2 log2 u)
        const float a = 6.2831853072f * u01<float>(x1);
        z0 = r * __cosf(a); z1 = r * __sinf(a);
    } else {
        const T r = sqrt_(T(-2) * log_(u01<T>(x0)));
        T s, c;
        sincos_(T(2 * kPi) * u01<T>(x1), &s, &c);
        z0 = r * c; z1 = r * s;
    }
}
// cube root of a uniform in (0,1)
template <typename T> DSIM_DEV T cbrt01(T u) {
    if constexpr (std::is_same<T, float>::value) return exp2f(__log2f(u) * 0.33333333333f); else return cbrt(u);
}
template <typename T> DSIM_DEV T clipn(T z, T sigma) { const T v = z * sigma, lim = T(2) * sigma; return clamp_(v, -lim, lim); }

// mujoco_rpy2quat (transformation.py:10-12): qz(yaw) * qy(pitch) * qx(roll).  Reset path only (angles are sampled in
// [-pi, pi], half angles in [-pi/2, pi/2]): FP32 uses the MUFU sin / cos (abs. error < 5e-7 on that range).
template <typename T> DSIM_DEV void rpy_to_quat(T roll, T pitch, T yaw, T &qw, T &qx, T &qy, T &qz) {
    T sr, cr, sp, cp, sy, cy;
    if constexpr (std::is_same<T, float>::value) {
        sr = __sinf(0.5f * roll); cr = __cosf(0.5f * roll); sp = __sinf(0.5f * pitch); cp = __cosf(0.5f * pitch);
        sy = __sinf(0.5f * yaw); cy = __cosf(0.5f * yaw);
    } else {
        sincos_(T(0.5) * roll, &sr, &cr); sincos_(T(0.5) * pitch, &sp, &cp); sincos_(T(0.5) * yaw, &sy, &cy);
    }
    qw = cy * cp * cr + sy * sp * sr;
    qx = cy * cp * sr - sy * sp * cr;
    qy = cy * sp * cr + sy * cp * sr;
    qz = sy * cp * cr - cy * sp * sr;
}

template <typename T> struct ResetCfg {
    T start_yaw, max_pos_offset;
    T angle_sigma[2], vel_sigma[3], ang_vel_sigma[3], pend_rp_sigma[2], pend_vel_sigma[2];
    int random_start_pos;
};

// BaseDroneEnv.sample_state (:218-257), same draw order, Philox stream 0 keyed by (seed, global env id, reset_count).
// Activations `act` are NOT touched (Q3: they persist across episode resets).
template <typename T, bool PEND>
DSIM_DEV void sample_state(EnvState<T> &s, const ResetCfg<T> &rc, uint32_t seed, uint32_t env, uint32_t count) {
    T roll = 0, pitch = 0, yaw = rc.start_yaw;
    s.pos = mk(T(0), T(0), T(0));
    s.vel = mk(T(0), T(0), T(0)); s.om = mk(T(0), T(0), T(0));
    s.hx = s.hy = s.hvx = s.hvy = T(0);
    if (rc.random_start_pos) {
        T n0, n1, n2, n3;
        U4 x = philox4x32(0, count, 0, 0, seed, env);
        box_muller(x.x, x.y, n0, n1); box_muller(x.z, x.w, n2, n3);
        const T inn = rsqrt_(n0 * n0 + n1 * n1 + n2 * n2);
        x = philox4x32(1, count, 0, 0, seed, env);
        const T r = rc.max_pos_offset * cbrt01(u01<T>(x.x));
        s.pos = mk(r * (n0 * inn), r * (n1 * inn), r * (n2 * inn));
        yaw = T(kPi) - T(2 * kPi) * u01<T>(x.y);
        box_muller(x.z, x.w, n0, n1);
        roll = clipn(n0, rc.angle_sigma[0]); pitch = clipn(n1, rc.angle_sigma[1]);
        x = philox4x32(2, count, 0, 0, seed, env);
        box_muller(x.x, x.y, n0, n1); box_muller(x.z, x.w, n2, n3);
        s.vel = mk(clipn(n0, rc.vel_sigma[0]), clipn(n1, rc.vel_sigma[1]), clipn(n2, rc.vel_sigma[2]));
        s.om.x = clipn(n3, rc.ang_vel_sigma[0]);
        x = philox4x32(3, count, 0, 0, seed, env);
        box_muller(x.x, x.y, n0, n1); box_muller(x.z, x.w, n2, n3);
        s.om.y = clipn(n0, rc.ang_vel_sigma[1]); s.om.z = clipn(n1, rc.ang_vel_sigma[2]);
        if (PEND) {
            s.hx = clipn(n2, rc.pend_rp_sigma[0]); s.hy = clipn(n3, rc.pend_rp_sigma[1]);
            x = philox4x32(4, count, 0, 0, seed, env);
            box_muller(x.x, x.y, n0, n1);
            s.hvx = clipn(n0, rc.pend_vel_sigma[0]); s.hvy = clipn(n1, rc.pend_vel_sigma[1]);
        }
    }
    rpy_to_quat(roll, pitch, yaw, s.qw, s.qx, s.qy, s.qz);
}

}  // namespace dsim
