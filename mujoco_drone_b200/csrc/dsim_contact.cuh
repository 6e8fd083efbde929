// dsim_contact.cuh — floor contact: the one mj_step stage the in-air fast path leaves out.
//
// Model (environments/env_gen.py:14-21,45-72,97): every drone geom has contype 1 / conaffinity 0, condim 3, friction (1, .5, .5),
// margin 0; the floor is a default plane geom at z = 0 (contype = conaffinity = 1).  Drone geoms can therefore touch the floor
// and nothing else.  Everything else is MuJoCo's default: solref (0.02, 1), solimp (0.9, 0.95, 0.001, 0.5, 2), friction combined
// by max -> mu = 1, pyramidal cone, impratio 1.  What mj_step does with that, per substep:
//   mj_collision      plane vs box (corners below the plane that point down, at most 4), plane vs cylinder (deepest rim point,
//                     the same direction on the far cap, two more rim points 120 degrees away), plane vs sphere
//   mj_makeConstraint four pyramid rows per contact, J = J_n +- mu J_t1, J_n +- mu J_t2, position = contact distance
//   mj_makeImpedance  d(r), R = 2 mu^2 (1 - d) / d * diagApprox, diagApprox = (1 + mu^2) * body_invweight0, aref = -B v - K d r
//   mj_fwdConstraint  qacc = argmin 1/2 (x - qacc_smooth)^T M (x - qacc_smooth) + sum_rows 1/2 D min(0, J x - aref)^2
//   mj_EulerSkip      (M + h B) qacc_integrated = qfrc_smooth + qfrc_constraint
// [MuJoCo's algorithm restated from its documentation; UNPINNED like the rest of the physics.]
//
// This is the slow path of the step kernel: a lane enters it only when its drone's bounding sphere reaches the floor.  It works in
// the kernel's own generalized coordinates x = (origin acceleration in body axes, angular acceleration in body axes, hinge x,
// hinge y) - a rotation of MuJoCo's, under which the convex problem is invariant - on a dense 8 x 8 mass matrix, with every
// geometric quantity expressed in the body frame (floor normal n_b = R^T z, height of a body point p: z_o + n_b . p).
// The geometry comes from a per-env table the model compile fills in FP64 with the compiler's "%.5g" rounding (geometry_kernel).
#pragma once
#include "dsim_device.cuh"
#include "dsim_params.cuh"

namespace dsim {

constexpr int kMaxContacts = 68;          // 7 boxes x 4 + 9 cylinders x 4 + 1 sphere = 65

// Per-env collision geometry, written by geometry_kernel next to the model compile (FP64, "%.5g" rounding like the compiler)
enum GeoRow : int {
    G_CORE_XY = 0, G_CORE_Z,              // core box half sizes
    G_FRONT_X, G_FRONT_SX, G_FRONT_SYZ,   // 'front' marker box: centre x, half sizes
    G_ARM_SX, G_ARM_SYZ,                  // arm boxes: half length, half thickness
    G_PROP_R,                             // propeller disc radius
    G_MOTOR_Z, G_PROP_Z,                  // heights of the motor / propeller cylinders
    G_ARM_X0,                             // [4] arm centre x      (per arm i: +i)
    G_ARM_Y0 = G_ARM_X0 + 4,              // [4] arm centre y
    G_ROT_X0 = G_ARM_Y0 + 4,              // [4] motor / propeller centre x
    G_ROT_Y0 = G_ROT_X0 + 4,              // [4] motor / propeller centre y
    G_YAW_C0 = G_ROT_Y0 + 4,              // [4] cos of the arm's rounded yaw
    G_YAW_S0 = G_YAW_C0 + 4,              // [4] sin
    G_PEND = G_YAW_S0 + 4,                // 1: this drone has a pendulum (env_gen.py:33-36)
    G_POLE_HH, G_POLE_Z, G_WEIGHT_S, G_WEIGHT_Z,
    GEO_ROWS
};
DSIM_HD void contact_geometry(const double p[6], bool pendulum_enabled, bool rounding, double out[GEO_ROWS]) {
#define RND(v) (rounding ? round_prec5(v) : (v))
    const double hb = 0.05, arm = p[1], r2 = sqrt(2.0);
    out[G_CORE_XY] = RND(hb); out[G_CORE_Z] = RND(hb / 3);
    out[G_FRONT_X] = RND(hb + hb / 3); out[G_FRONT_SX] = RND(hb / 3); out[G_FRONT_SYZ] = RND(0.15 * hb);
    out[G_ARM_SX] = RND(arm / 2); out[G_ARM_SYZ] = RND(arm / 20);
    out[G_PROP_R] = RND(arm / 1.5);
    out[G_MOTOR_Z] = RND(0.0 + 0.015); out[G_PROP_Z] = RND(0.0 + 0.025);
    for (int i = 0; i < 4; i++) {
        const double th = i * kPi / 2 - kPi / 4, ct = cos(th), st = sin(th), yaw = RND(th);
        const double ra = r2 * hb + 0.5 * arm, rr = r2 * hb + arm;
        out[G_ARM_X0 + i] = RND(ra * ct); out[G_ARM_Y0 + i] = RND(ra * st);
        out[G_ROT_X0 + i] = RND(rr * ct + 0.0); out[G_ROT_Y0 + i] = RND(rr * st + 0.0);
        out[G_YAW_C0 + i] = cos(yaw); out[G_YAW_S0 + i] = sin(yaw);
    }
    const bool pend = pendulum_enabled && p[4] > 0 && p[5] > 0;
    out[G_PEND] = pend ? 1.0 : 0.0;
    out[G_POLE_HH] = pend ? RND(p[4] / 2) : 0.0; out[G_POLE_Z] = pend ? RND(-p[4] / 2) : 0.0;
    out[G_WEIGHT_S] = pend ? RND(0.1 * cbrt(p[5])) : 0.0; out[G_WEIGHT_Z] = pend ? RND(-p[4]) : 0.0;
#undef RND
}

template <typename T> struct GroundCtx {
    T start_z;                            // world z of the origin the state's position offset refers to
    T reach;                              // radius of a sphere about the body origin that holds every geom (conservative)
    const T *geo;                         // (row 0, this env) of the geometry table [GEO_ROWS][ld]
    int ld;
};

// bounding radius from the raw parameters: propeller rim (0.0982 + 1.667 arm_len) / far corner of the weight (0.025 + l + 0.1733 cbrt(m))
template <typename T> DSIM_DEV GroundCtx<T> make_ground_ctx(T start_z, const T *prm, const T *geo, int ld, int env, int pendulum) {
    GroundCtx<T> g;
    g.start_z = start_z; g.geo = geo + env; g.ld = ld;
    const T rb = T(0.1) + T(1.7) * prm[1], rp = pendulum ? T(0.03) + prm[4] + T(0.18) * max_(T(1), prm[5]) : T(0);
    g.reach = max_(rb, rp) * T(1.001);
    return g;
}

template <typename T> struct ContactSet {
    int n;
    T p[kMaxContacts][3];                 // contact position, body coordinates (half-way between the geom point and the floor)
    T kr[kMaxContacts];                   // K d(r) r
    T D[kMaxContacts];                    // 1 / R of the contact's four rows
    unsigned char body[kMaxContacts];     // 0 core body, 1 link, 2 pendulum
};

template <typename T> struct ContactIO {
    T q[8];                               // qfrc_smooth (force, torque about the origin, hinge x, hinge y), body axes
    T v[8];                               // qvel in the same coordinates
    T x[8];                               // in: qacc_smooth; out: qacc (explicit: feeds the accelerometer)
    T xi[8];                              // out: qacc of the implicit-damping Euler step
    T nb[3], t1[3], t2[3];                // floor normal and the two tangents of the contact frame (+y, -x of the world) in body axes
    T zo;                                 // world height of the body origin
    T sx, cx, sy, cy;                     // hinge angles
    T h;                                  // timestep
};

template <typename T> DSIM_DEV void add_contact(ContactSet<T> &cs, V3<T> p, T dist, int body, V3<T> nb) {
    if (cs.n >= kMaxContacts) return;
    const V3<T> pc = p - (T(0.5) * dist) * nb;
    cs.p[cs.n][0] = pc.x; cs.p[cs.n][1] = pc.y; cs.p[cs.n][2] = pc.z;
    cs.kr[cs.n] = dist;                   // the distance for now; turned into K d(r) r by finish_contacts
    cs.body[cs.n] = (unsigned char)body;
    cs.n++;
}
// mjc_PlaneBox
template <typename T> DSIM_DEV void collide_box(ContactSet<T> &cs, V3<T> nb, T zo, V3<T> c, V3<T> ex, V3<T> ey, V3<T> ez, T hx, T hy, T hz, int body) {
    const T dist0 = zo + dot(nb, c);
    if (dist0 > abs_(hx * dot(nb, ex)) + abs_(hy * dot(nb, ey)) + abs_(hz * dot(nb, ez))) return;
    int cnt = 0;
    #pragma unroll 1
    for (int i = 0; i < 8 && cnt < 4; i++) {
        const V3<T> corner = ((i & 1) ? hx : -hx) * ex + ((i & 2) ? hy : -hy) * ey + ((i & 4) ? hz : -hz) * ez;
        const T ld = dot(nb, corner);
        if (dist0 + ld > T(0) || ld > T(0)) continue;
        add_contact(cs, c + corner, dist0 + ld, body, nb);
        cnt++;
    }
}
// mjc_PlaneSphere
template <typename T> DSIM_DEV void collide_sphere(ContactSet<T> &cs, V3<T> nb, T zo, V3<T> c, T r, int body) {
    const T dist = zo + dot(nb, c) - r;
    if (dist > T(0)) return;
    add_contact(cs, c - r * nb, dist, body, nb);
}
// mjc_PlaneCylinder (ex: the cylinder's x axis, used when the disc is parallel to the floor; ez: its axis)
template <typename T> DSIM_DEV void collide_cylinder(ContactSet<T> &cs, V3<T> nb, T zo, V3<T> c, V3<T> ex, V3<T> ez, T radius, T hh, int body) {
    const T dist0 = zo + dot(nb, c);
    if (dist0 > radius + hh) return;
    V3<T> axis = ez;
    T prjaxis = dot(nb, axis);
    if (prjaxis > T(0)) { axis = T(-1) * axis; prjaxis = -prjaxis; }
    V3<T> vec = prjaxis * axis - nb;
    const T len2 = dot(vec, vec);
    if (len2 < T(1e-24)) vec = radius * ex;
    else vec = (radius * rsqrt_(len2)) * vec;
    const T prjvec = dot(vec, nb);
    axis = hh * axis; prjaxis *= hh;
    if (dist0 + prjaxis + prjvec > T(0)) return;
    add_contact(cs, c + axis + vec, dist0 + prjaxis + prjvec, body, nb);
    if (dist0 - prjaxis + prjvec <= T(0)) add_contact(cs, c - axis + vec, dist0 - prjaxis + prjvec, body, nb);
    const T prjvec1 = T(-0.5) * prjvec;
    if (dist0 + prjaxis + prjvec1 <= T(0)) {
        V3<T> side = cross(vec, axis);
        side = (radius * T(0.8660254037844386) * rsqrt_(max_(dot(side, side), T(1e-37)))) * side;
        const V3<T> mid = c + axis - T(0.5) * vec;
        add_contact(cs, mid - side, dist0 + prjaxis + prjvec1, body, nb);
        add_contact(cs, mid + side, dist0 + prjaxis + prjvec1, body, nb);
    }
}

// dense mass matrix in the kernel's coordinates (the same composite-body quantities substep() reduces analytically)
template <typename T> DSIM_DEV void dense_mass(const EnvConsts<T> &c, bool pend, T sx, T cx, T sy, T cy, T *M /*[64]*/) {
    const T mC = pend ? T(kMassC) : T(0), IC = pend ? T(kInertiaC) : T(0), dl = T(kLinkDrop);
    const T mD = pend ? c.mD : T(0);
    const T mh = mC + mD, mtot = c.mB + mh;
    const V3<T> yc = mk(T(0), cx, sx), n = mk(sy, -sx * cy, cx * cy), xd = mk(cy, sx * sy, -cx * sy);
    const T mu = mD * c.zD, P = pend ? c.IDx + mD * c.zD * c.zD : T(0), QmP = pend ? c.IDz - P : T(0);
    const V3<T> H = mk(mu * n.x, mu * n.y, c.mB * c.cz - mh * dl + mu * n.z);
    const T k12 = c.mB * c.cz * c.cz + IC + mh * dl * dl + P - T(2) * mu * dl * n.z;
    const T Ixx = c.IBx + k12 + QmP * n.x * n.x, Iyy = c.IBy + k12 + QmP * n.y * n.y, Izz = c.IBz + IC + P + QmP * n.z * n.z;
    const T Ixy = QmP * n.x * n.y, Ixz = QmP * n.x * n.z + mu * dl * n.x, Iyz = QmP * n.y * n.z + mu * dl * n.y;
    const V3<T> px = (-mu * cy) * yc, py = mu * xd;
    const V3<T> Lx = mk(IC + P + QmP * n.x * n.x, QmP * n.x * n.y, QmP * n.x * n.z) + mk(dl * px.y, -dl * px.x, T(0));
    const V3<T> Ly = P * yc + mk(dl * py.y, -dl * py.x, T(0));
    for (int k = 0; k < 64; k++) M[k] = T(0);
    M[0] = M[9] = M[18] = mtot;
    M[0 * 8 + 4] = H.z; M[0 * 8 + 5] = -H.y; M[1 * 8 + 3] = -H.z; M[1 * 8 + 5] = H.x; M[2 * 8 + 3] = H.y; M[2 * 8 + 4] = -H.x;
    M[3 * 8 + 3] = Ixx; M[4 * 8 + 4] = Iyy; M[5 * 8 + 5] = Izz; M[3 * 8 + 4] = Ixy; M[3 * 8 + 5] = Ixz; M[4 * 8 + 5] = Iyz;
    if (pend) {
        M[0 * 8 + 6] = px.x; M[1 * 8 + 6] = px.y; M[2 * 8 + 6] = px.z; M[3 * 8 + 6] = Lx.x; M[4 * 8 + 6] = Lx.y; M[5 * 8 + 6] = Lx.z;
        M[0 * 8 + 7] = py.x; M[1 * 8 + 7] = py.y; M[2 * 8 + 7] = py.z; M[3 * 8 + 7] = Ly.x; M[4 * 8 + 7] = Ly.y; M[5 * 8 + 7] = Ly.z;
        M[6 * 8 + 6] = IC + P + QmP * n.x * n.x; M[7 * 8 + 7] = P;
    } else { M[6 * 8 + 6] = T(1); M[7 * 8 + 7] = T(1); }
    for (int i = 0; i < 8; i++) for (int j = 0; j < i; j++) M[i * 8 + j] = M[j * 8 + i];
}
// in-place Cholesky (lower triangle) and solve, 8 x 8.  Every loop is fully unrolled so that a caller's matrix with
// compile-time indices stays in registers (the slow path is bound by the latency of its local-memory arrays)
template <typename T> DSIM_DEV void chol8(T *A) {
    #pragma unroll
    for (int j = 0; j < 8; j++) {
        T d = A[j * 8 + j];
        #pragma unroll
        for (int k = 0; k < j; k++) d -= A[j * 8 + k] * A[j * 8 + k];
        d = sqrt_(max_(d, T(1e-30)));
        A[j * 8 + j] = d;
        const T id = T(1) / d;
        #pragma unroll
        for (int i = j + 1; i < 8; i++) {
            T s = A[i * 8 + j];
            #pragma unroll
            for (int k = 0; k < j; k++) s -= A[i * 8 + k] * A[j * 8 + k];
            A[i * 8 + j] = s * id;
        }
    }
}
template <typename T> DSIM_DEV void chol8_solve(const T *L, T *b) {
    #pragma unroll
    for (int i = 0; i < 8; i++) {
        T s = b[i];
        #pragma unroll
        for (int k = 0; k < i; k++) s -= L[i * 8 + k] * b[k];
        b[i] = s / L[i * 8 + i];
    }
    #pragma unroll
    for (int i = 7; i >= 0; i--) {
        T s = b[i];
        #pragma unroll
        for (int k = i + 1; k < 8; k++) s -= L[k * 8 + i] * b[k];
        b[i] = s / L[i * 8 + i];
    }
}
// Jacobian row of direction d (body axes) at body point p moving with `body`: J x = d . (a + al x p + hinge terms)
template <typename T> DSIM_DEV void contact_row(V3<T> d, V3<T> p, int body, V3<T> yc, T *J) {
    const V3<T> pd = cross(p, d);
    J[0] = d.x; J[1] = d.y; J[2] = d.z; J[3] = pd.x; J[4] = pd.y; J[5] = pd.z;
    const V3<T> r = p - mk(T(0), T(0), T(-kLinkDrop));              // lever about the hinge point
    J[6] = body >= 1 ? dot(d, cross(mk(T(1), T(0), T(0)), r)) : T(0);
    J[7] = body >= 2 ? dot(d, cross(yc, r)) : T(0);
}

// Returns the number of contacts; with none, io.x / io.xi are left alone.
template <typename T>
__device__ DSIM_CONTACT_CALL int contact_solve(ContactIO<T> &io, const EnvConsts<T> &c, const GroundCtx<T> &g) {
    ContactSet<T> cs;
    cs.n = 0;
    const V3<T> nb = mk(io.nb[0], io.nb[1], io.nb[2]);
    const T zo = io.zo;
    const V3<T> X = mk(T(1), T(0), T(0)), Y = mk(T(0), T(1), T(0)), Z = mk(T(0), T(0), T(1)), O = mk(T(0), T(0), T(0));
    auto G = [&](int row) { return g.geo[(size_t)row * g.ld]; };
    const bool pend = G(G_PEND) != T(0);
    {   // ---- core body (env_gen.py:45-61)
        collide_box(cs, nb, zo, O, X, Y, Z, G(G_CORE_XY), G(G_CORE_XY), G(G_CORE_Z), 0);
        collide_box(cs, nb, zo, mk(G(G_FRONT_X), T(0), T(0)), X, Y, Z, G(G_FRONT_SX), G(G_FRONT_SYZ), G(G_FRONT_SYZ), 0);
        const T asx = G(G_ARM_SX), asyz = G(G_ARM_SYZ), pr = G(G_PROP_R), mz = G(G_MOTOR_Z), pz = G(G_PROP_Z);
        #pragma unroll 1
        for (int i = 0; i < 4; i++) {
            const T cyw = G(G_YAW_C0 + i), syw = G(G_YAW_S0 + i);
            collide_box(cs, nb, zo, mk(G(G_ARM_X0 + i), G(G_ARM_Y0 + i), T(0)), mk(cyw, syw, T(0)), mk(-syw, cyw, T(0)), Z, asx, asyz, asyz, 0);
            const T mx = G(G_ROT_X0 + i), my = G(G_ROT_Y0 + i);
            collide_cylinder(cs, nb, zo, mk(mx, my, mz), X, Z, T(0.01), T(0.01), 0);
            collide_cylinder(cs, nb, zo, mk(mx, my, pz), X, Z, pr, T(0.0025), 0);
        }
    }
    const V3<T> yc = mk(T(0), io.cx, io.sx), n = mk(io.sy, -io.sx * io.cy, io.cx * io.cy), xd = mk(io.cy, io.sx * io.sy, -io.cx * io.sy);
    const V3<T> hp = mk(T(0), T(0), T(-kLinkDrop));
    if (pend) {   // ---- link sphere, pole, weight (env_gen.py:66-72)
        collide_sphere(cs, nb, zo, hp, T(0.02), 1);
        collide_cylinder(cs, nb, zo, hp + G(G_POLE_Z) * n, xd, n, T(0.005), G(G_POLE_HH), 2);
        const T sw = G(G_WEIGHT_S);
        collide_box(cs, nb, zo, hp + G(G_WEIGHT_Z) * n, xd, yc, n, sw, sw, sw, 2);
    }
    if (cs.n == 0) return 0;

    // ---- mj_makeImpedance: body_invweight0 at qpos0 (identity attitude, hinges at 0), then K, B, d(r), R per contact
    T M[64], L[64];
    T tran[3];
    {
        dense_mass(c, pend, T(0), T(1), T(0), T(1), M);
        for (int k = 0; k < 64; k++) L[k] = M[k];
        chol8(L);
        const V3<T> com[3] = {mk(T(0), T(0), c.cz), hp, hp + mk(T(0), T(0), c.zD)};
        const V3<T> ax[3] = {X, Y, Z};
        for (int b = 0; b < (pend ? 3 : 1); b++) {
            T acc = T(0);
            for (int a = 0; a < 3; a++) {
                T J[8], y[8];
                contact_row(ax[a], com[b], b, Y, J);
                for (int k = 0; k < 8; k++) y[k] = J[k];
                chol8_solve(L, y);
                for (int k = 0; k < 8; k++) acc += J[k] * y[k];
            }
            tran[b] = max_(T(kMinVal), acc * T(1.0 / 3.0));
        }
    }
    const T tc = max_(T(0.02), T(2) * io.h), dmax = T(0.95);
    const T K = T(1) / (dmax * dmax * tc * tc), B = T(2) / (dmax * tc);
    for (int i = 0; i < cs.n; i++) {
        const T dist = cs.kr[i];
        const T xr = min_(abs_(dist) * T(1000.0), T(1));                       // |r| / width
        const T y = xr <= T(0.5) ? T(2) * xr * xr : T(1) - T(2) * (T(1) - xr) * (T(1) - xr);
        const T imp = T(0.9) + T(0.05) * y;
        const T R = T(2) * max_(T(kMinVal), (T(1) - imp) / imp * (T(2) * tran[cs.body[i]]));   // mu = 1: diagApprox = 2 tran, Rpy = 2 R
        cs.D[i] = T(1) / R;
        cs.kr[i] = K * imp * dist;
    }

    // ---- Newton with an exact line search on the primal problem
    dense_mass(c, pend, io.sx, io.cx, io.sy, io.cy, M);
    const V3<T> t1 = mk(io.t1[0], io.t1[1], io.t1[2]), t2 = mk(io.t2[0], io.t2[1], io.t2[2]);
    auto row = [&](int i, int r, T *J) -> T {                                  // fills J, returns aref
        const V3<T> d = r == 0 ? nb + t1 : (r == 1 ? nb - t1 : (r == 2 ? nb + t2 : nb - t2));
        contact_row(d, mk(cs.p[i][0], cs.p[i][1], cs.p[i][2]), cs.body[i], yc, J);
        T vel = T(0);
        for (int k = 0; k < 8; k++) vel += J[k] * io.v[k];
        return -B * vel - cs.kr[i];
    };
    T x[8], gq[8], grad[8], dx[8];
    T jar_[4 * kMaxContacts], jd_[4 * kMaxContacts];                          // per row: J x - aref at x, J dx (the line search reads only these)
    for (int k = 0; k < 8; k++) x[k] = io.x[k];
    // Newton converges quadratically once the active set is right: the iteration whose step falls below the precision's
    // resolution of x was the last useful one.  FP32: the gradient carries ~1e-6 of rounding noise per newton of force, so the
    // tolerances sit just above that floor and the iteration cap does the rest (measured: 10x looser tolerances buy 10 %)
    T trace = T(0);
    for (int k = 0; k < 8; k++) trace += M[k * 8 + k];
    const bool f32 = sizeof(T) == 4;
    const T tol = (f32 ? T(2e-6) : T(1e-13)) * trace;
    #pragma unroll 1
    for (int it = 0; it < (f32 ? 16 : 40); it++) {
        for (int k = 0; k < 8; k++) { T s = -io.q[k]; for (int j = 0; j < 8; j++) s += M[k * 8 + j] * x[j]; gq[k] = s; grad[k] = s; }
        #pragma unroll
        for (int k = 0; k < 64; k++) L[k] = M[k];
        #pragma unroll 1
        for (int i = 0; i < cs.n; i++)
            #pragma unroll 1
            for (int r = 0; r < 4; r++) {
                T J[8];
                const T aref = row(i, r, J);
                T jar = -aref;
                #pragma unroll
                for (int k = 0; k < 8; k++) jar += J[k] * x[k];
                jar_[4 * i + r] = jar;
                if (jar >= T(0)) continue;
                const T D = cs.D[i];
                #pragma unroll
                for (int k = 0; k < 8; k++) {
                    grad[k] += D * jar * J[k];
                    const T dj = D * J[k];
                    #pragma unroll
                    for (int j = 0; j <= k; j++) L[k * 8 + j] += dj * J[j];
                }
            }
        T gn = T(0);
        for (int k = 0; k < 8; k++) gn += grad[k] * grad[k];
        if (sqrt_(gn) < tol) break;
        chol8(L);
        for (int k = 0; k < 8; k++) dx[k] = -grad[k];
        chol8_solve(L, dx);
        // phi'(a) = p0 + a p2 + sum_rows D jd min(0, jar + a jd): increasing and piecewise linear; walk its breakpoints
        T p0 = T(0), p2 = T(0);
        for (int k = 0; k < 8; k++) { p0 += gq[k] * dx[k]; T s = T(0); for (int j = 0; j < 8; j++) s += M[k * 8 + j] * dx[j]; p2 += dx[k] * s; }
        #pragma unroll 1
        for (int i = 0; i < cs.n; i++)
            #pragma unroll 1
            for (int r = 0; r < 4; r++) {
                T J[8];
                row(i, r, J);
                T jd = T(0);
                for (int k = 0; k < 8; k++) jd += J[k] * dx[k];
                jd_[4 * i + r] = jd;
            }
        T a = T(0);
        #pragma unroll 1
        for (int ls = 0; ls < 64; ls++) {
            T f1 = p0 + a * p2, f2 = p2, nxt = T(1e30);
            #pragma unroll 1
            for (int i = 0; i < 4 * cs.n; i++) {
                const T jar = jar_[i], jd = jd_[i], D = cs.D[i >> 2];
                bool act;
                if (jd != T(0)) {
                    const T bp = -jar / jd;
                    act = jd < T(0) ? (a >= bp) : (a < bp);
                    if (bp > a && bp < nxt) nxt = bp;
                } else act = jar < T(0);
                if (act) { f1 += D * jd * (jar + a * jd); f2 += D * jd * jd; }
            }
            if (f1 >= T(0)) break;
            const T root = a - f1 / f2;
            if (root <= nxt) { a = root; break; }
            a = nxt;
        }
        T big = T(0);
        for (int k = 0; k < 8; k++) { x[k] += a * dx[k]; big = max_(big, abs_(a * dx[k]) / (T(1) + abs_(x[k]))); }
        if (big < (f32 ? T(2e-6) : T(1e-15))) break;
    }
    for (int k = 0; k < 8; k++) io.x[k] = x[k];
    // ---- mj_EulerSkip with the constraint force: (M + h B) xi = qfrc_smooth + qfrc_constraint = M x  ->  xi = x - (M + h B)^-1 h B x
    {
        const T hb = pend ? io.h * T(kHingeDamping) : T(0);
        for (int k = 0; k < 64; k++) L[k] = M[k];
        L[6 * 8 + 6] += hb; L[7 * 8 + 7] += hb;
        chol8(L);
        T y[8] = {T(0), T(0), T(0), T(0), T(0), T(0), hb * x[6], hb * x[7]};
        chol8_solve(L, y);
        for (int k = 0; k < 8; k++) io.xi[k] = x[k] - y[k];
    }
    return cs.n;
}

// One mj_step of a drone that may touch the floor, as ONE out-of-line call: the step kernel's in-air path stays exactly the
// code of the kernels without floor contact (no struct in local memory, no registers saved around a call).
template <typename T, bool PEND>
__device__ __noinline__ void ground_substep(EnvState<T> &s, const EnvConsts<T> &c, const T *ctrl, T h, const GroundCtx<T> &g) {
    substep<T, PEND, true, true>(s, c, ctrl, h, &g);
}

}  // namespace dsim
