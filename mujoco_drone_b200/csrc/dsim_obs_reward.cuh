// dsim_obs_reward.cuh — state extraction, termination, the 17 reward functions and the 14 observation
// variants of the reference, evaluated in registers at the end of the fused step.
//   environments/BaseDroneEnv.py:12-16 (termination), :357-380 (get_drone_states)
//   environments/rewards.py:5-368, environments/observation_wrappers.py:7-529
// Reference quirks are replicated on purpose: PRY variants swap roll/pitch, NoPend
// variants index state[16:19], reward_pendulum_dist uses params[5], the `_en*` rewards sum a broadcast 3x3.
#pragma once
#include "dsim_device.cuh"

namespace dsim {

// everything the reward / observation code needs about one env after the physics
template <typename T> struct PostState {
    T roll, pitch, yaw;       // mujoco_quat2rpy of the post-step quaternion
    M3<T> R;                  // body -> world (== DCM(rpy2quat(rpy)) that the wrappers rebuild)
    V3<T> err;                // pos - ref[:3]  (both held as offsets from start_pos)
    T ref_yaw;
    T pos_err2;
};

template <typename T> DSIM_DEV PostState<T> post_state(const EnvState<T> &s, V3<T> ref_off, T ref_yaw) {
    PostState<T> p;
    quat_to_rpy(s.qw, s.qx, s.qy, s.qz, p.roll, p.pitch, p.yaw);
    const T qi = rsqrt_(s.qw * s.qw + s.qx * s.qx + s.qy * s.qy + s.qz * s.qz);
    p.R = quat_to_mat(s.qw * qi, s.qx * qi, s.qy * qi, s.qz * qi);
    p.err = s.pos - ref_off;
    p.ref_yaw = ref_yaw;
    p.pos_err2 = dot(p.err, p.err);
    return p;
}

// default_termination_fcn (BaseDroneEnv.py:12-16) in FP64 on the stored state, operation order fixed
// ((dx*dx + dy*dy) + dz*dz, no FMA contraction) so the truncation bit equals the CPU oracle's on identical inputs.
// `max_d2` = the largest double whose correctly rounded square root is <= max_distance (computed on the host,
// max_distance_sq_threshold): sqrt_rn is monotonic, so  sqrt_rn(d2) > max_distance  <=>  d2 > max_d2  bit for bit, and
// the kernel needs no FP64 square root.
template <typename T>
DSIM_DEV bool terminated(V3<T> pos_off, const double start[3], const double ref[3], double max_d2, int num_steps, int max_steps) {
    const double dx = __dsub_rn(__dadd_rn(start[0], (double)pos_off.x), ref[0]);
    const double dy = __dsub_rn(__dadd_rn(start[1], (double)pos_off.y), ref[1]);
    const double dz = __dsub_rn(__dadd_rn(start[2], (double)pos_off.z), ref[2]);
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    return (d2 > max_d2) || (num_steps >= max_steps);
}

template <typename T> DSIM_DEV T sq(T x) { return x * x; }

// rewards.py — `a` is the RAW action (before the 0.1 + 0.9 a remap), `ns` = num_steps after the increment
template <typename T, bool PEND>
DSIM_DEV T reward_fn(int id, const EnvState<T> &s, const PostState<T> &p, const T a[4], int ns, const T prm[6], T max_distance) {
    const T he = wrap_pi(abs_(p.yaw - p.ref_yaw));                         // ((|yaw-ref|+pi) % 2pi) - pi
    const T pe2 = p.pos_err2, pe = sqrt_(pe2);
    const T a2 = sq(a[0]) + sq(a[1]) + sq(a[2]) + sq(a[3]);
    if (id == 0) return T(3) - pe;
    if (id == 1) return T(5) - pe - T(0.1) * abs_(he);
    if (id == 2) return T(3.5) - pe2 - T(0.1) * abs_(he) - T(0.2) * a2;
    if (id == 10) {
        const T too_far = pe2 > max_distance * max_distance ? T(1) : T(0);
        return -(T(1) + T(ns / 50)) * pe2 - T(500) * too_far - abs_(he) - T(0.02) * a2;
    }
    if (!PEND) return T(0);   // remaining rewards index pendulum entries of the 33-layout: host rejects them without a pendulum
    const T prp2 = sq(s.hx) + sq(s.hy), pav2 = sq(s.hvx) + sq(s.hvy), om2 = dot(s.om, s.om);
    if (id == 3) return T(3.5) - pe2 - T(0.2) * he * he - T(0.2) * a2 - T(0.2) * prp2;
    if (id == 4) return T(3.5) - pe2 - T(0.5) * he * he - T(0.4) * a2 - T(0.2) * prp2 - T(0.1) * om2;
    if (id == 5) {
        T r = T(3.5) - pe2 - T(0.5) * he * he - T(0.4) * a2;
        r -= (T(0.1) * prp2 + T(0.2) * pav2 - T(0.3) * (sq(p.roll) + sq(p.pitch)) - T(0.4) * om2) / (T(1) + T(100) * pe2);
        return r;
    }
    if (id == 11) {
        const T close = pe2 < T(0.2) ? T(1) : T(0), too_far = pe2 > max_distance * max_distance - T(3) ? T(1) : T(0);
        return (T(7) + T(20) * close - T(3) * pe2 * (T(1) + T(ns) / T(150)) - T(10) * too_far - T(0.3) * (sq(p.roll) + sq(p.pitch))
                - T(0.7) * he * he - T(0.3) * a2 - T(0.3) * dot(s.vel, s.vel) - T(0.5) * pav2) / T(10);
    }
    T sx, cx, sy, cy;
    sincos_hinge(s.hx, &sx, &cx); sincos_hinge(s.hy, &sy, &cy);
    if (id >= 6 && id <= 9) {
        // rewards.py:81-103: tip velocity w = Rd( w_b x Rp e + (Rx Ox Ry + Rx Ry Oy) e ), e = (0,0,-L), Rp = Rx Ry
        const T Lp = prm[4];
        const V3<T> n = mk(sy, -sx * cy, cx * cy), yc = mk(T(0), cx, sx), xd = mk(cy, sx * sy, -cx * sy);
        const V3<T> wl = (-Lp) * cross(s.om, n) + (s.hvx * Lp * cy) * yc - (s.hvy * Lp) * xd;
        const V3<T> w = mul(p.R, wl);
        // `state[6:9] + (3,1) column` broadcasts to 3x3 in the reference; .sum() runs over all nine entries
        T en = 0;
        const T wv[3] = {w.x, w.y, w.z}, vv[3] = {s.vel.x, s.vel.y, s.vel.z};
        #pragma unroll
        for (int i = 0; i < 3; i++)
            #pragma unroll
            for (int j = 0; j < 3; j++) en += sq(vv[j] + wv[i]);
        if (id == 6) return T(3.5) - pe2 - T(0.5) * he * he - T(0.4) * a2 - T(0.2) * en;
        const T thr = id == 9 ? T(0.6) : T(0.5);
        T ce = 0;
        #pragma unroll
        for (int k = 0; k < 4; k++) ce += sq(max_(a[k] - thr, T(0)));
        const T angdev = sqrt_(sq(p.roll) + sq(p.pitch) + sq(p.yaw));
        if (id == 7) {
            T r = T(3.5) - T(2) * pe - T(0.6) * he * he - T(0.6) * ce;
            if (pe < T(0.15)) r = r + T(3) - T(0.2) * en - T(0.2) * angdev;
            return r;
        }
        const T ph = -Lp * (p.R.m[6] * n.x + p.R.m[7] * n.y + p.R.m[8] * n.z);
        const T tot = T(0.5) * en + T(9.81) * ph;
        if (id == 8) return T(7) - pe - T(0.4) * he * he - T(0.1) * ce - T(0.1) * tot - T(0.05) * angdev;
        return T(5) - pe - T(0.6) * he * he - T(0.1) * ce - (T(0.2) * tot + T(0.05) * angdev) / (T(0.5) + pe);
    }
    // ids 12..16: pendulum_R = DCM(rpy2quat([pr, pp, 0])) = Ry(pp) Rx(pr)  (NOT Rx Ry)
    const T Lp = (id == 12) ? prm[5] : prm[4];                              // Q12
    const V3<T> m = mk(sy * cx, -sx, cy * cx);                             // Ry Rx z^
    const V3<T> tip = p.err + mul(p.R, (-Lp) * m);                         // pendulum_pos - ref
    const T e = dot(tip, tip);
    if (id == 12) return -e;
    const T h = abs_(he);
    if (id == 13) return T(3) - e - T(0.1) * h;
    if (id == 14) return T(4) - e - T(0.001) * T(ns) * e - T(0.1) * h - T(0.05) * a2;
    const V3<T> vl = cross(mk(s.hvx, s.hvy, T(0)), (-prm[4]) * m);
    const V3<T> vg = s.vel + mul(p.R, vl);
    const T en = dot(vg, vg);
    if (id == 15) return T(4) - e - T(0.2) * h - T(0.006) * T(ns) * (e + T(0.2) * h) - T(0.05) * a2 - T(0.1) * en;
    T ce = 0;
    #pragma unroll
    for (int k = 0; k < 4; k++) ce += sq(min_(a[k] - T(0.5), T(0)));
    return T(4) - pe2 - T(0.2) * h - T(0.006) * T(ns) * (pe2 + T(0.2) * h + T(0.01) * en) - T(0.1) * ce - T(0.1) * en;
}

// get_drone_states row (BaseDroneEnv.py:357-380): 33 values with the pendulum, 29 without
template <typename T, bool PEND, typename Out>
DSIM_DEV int emit_state_row(const EnvState<T> &s, const PostState<T> &p, V3<T> start, V3<T> ref_off, const T prm[6], Out out) {
    int n = 0;
    out(n++, start.x + s.pos.x); out(n++, start.y + s.pos.y); out(n++, start.z + s.pos.z);
    out(n++, p.roll); out(n++, p.pitch); out(n++, p.yaw);
    out(n++, s.vel.x); out(n++, s.vel.y); out(n++, s.vel.z);
    out(n++, s.om.x); out(n++, s.om.y); out(n++, s.om.z);
    if (PEND) { out(n++, s.hx); out(n++, s.hy); out(n++, s.hvx); out(n++, s.hvy); }
    out(n++, s.acc.x); out(n++, s.acc.y); out(n++, s.acc.z);
    #pragma unroll
    for (int k = 0; k < 4; k++) out(n++, s.act[k]);
    out(n++, start.x + ref_off.x); out(n++, start.y + ref_off.y); out(n++, start.z + ref_off.z); out(n++, p.ref_yaw);
    #pragma unroll
    for (int k = 0; k < 6; k++) out(n++, prm[k]);
    return n;
}

// observation_wrappers.py `_get_obs` variants; `out(j, value)` stores component j
template <typename T, bool PEND, typename Out>
DSIM_DEV void emit_obs(int id, const EnvState<T> &s, const PostState<T> &p, V3<T> start, V3<T> ref_off, const T prm[6], Out out) {
    if (id == 0) { emit_state_row<T, PEND>(s, p, start, ref_off, prm, out); return; }
    const T hd = wrap_pi(p.ref_yaw - p.yaw);                               // (ref_yaw - yaw + pi) % 2pi - pi
    const V3<T> gerr = T(-1) * p.err;                                      // reference[:3] - xyz
    V3<T> lerr, lvel;
    if constexpr (std::is_same<T, float>::value) {
        // R^T applied to the position error and to the velocity: one packed pass (err | vel)
        const V3<F2> v2 = mk(F2(gerr.x, s.vel.x), F2(gerr.y, s.vel.y), F2(gerr.z, s.vel.z));
        const F2 lx = F2(p.R.m[0]) * v2.x + F2(p.R.m[3]) * v2.y + F2(p.R.m[6]) * v2.z;
        const F2 ly = F2(p.R.m[1]) * v2.x + F2(p.R.m[4]) * v2.y + F2(p.R.m[7]) * v2.z;
        const F2 lz = F2(p.R.m[2]) * v2.x + F2(p.R.m[5]) * v2.y + F2(p.R.m[8]) * v2.z;
        lerr = mk(lx.lo(), ly.lo(), lz.lo()); lvel = mk(lx.hi(), ly.hi(), lz.hi());
    } else { lerr = tmul(p.R, gerr); lvel = tmul(p.R, s.vel); }
    // the wrappers index the 33-layout: [12:14] pendulum_rp, [14:16] pendulum_ang_vel, [16:19] acc, [19:23] act.
    // Without a pendulum the 29-layout shifts: [12:15] acc, [15:19] act -> NoPend variants read act[1:4] as "acc" (Q11).
    const T s12 = PEND ? s.hx : s.acc.x, s13 = PEND ? s.hy : s.acc.y, s14 = PEND ? s.hvx : s.acc.z, s15 = PEND ? s.hvy : s.act[0];
    const T s16 = PEND ? s.acc.x : s.act[1], s17 = PEND ? s.acc.y : s.act[2], s18 = PEND ? s.acc.z : s.act[3];
    int n = 0;
    const bool global = (id == 1);
    out(n++, global ? gerr.x : lerr.x); out(n++, global ? gerr.y : lerr.y); out(n++, global ? gerr.z : lerr.z);
    const bool pry = (id == 2 || id == 3 || id == 5 || id == 6 || id == 7 || id == 11);
    if (id == 4 || id == 14) {                                             // z_vec = DCM(rpy2quat([r,p,0]))[:,2]
        T sr, cr, sp, cp;
        sincos_(p.roll, &sr, &cr); sincos_(p.pitch, &sp, &cp);
        out(n++, sp * cr); out(n++, -sr); out(n++, cp * cr);
    } else if (id == 13) {                                                 // Rm = DCM(rpy2quat([r,p,-hd])).T, flattened
        const M3<T> Rm = rpy_to_mat(p.roll, p.pitch, -hd);
        out(n++, Rm.m[0]); out(n++, Rm.m[3]); out(n++, Rm.m[6]);
        out(n++, Rm.m[1]); out(n++, Rm.m[4]); out(n++, Rm.m[7]);
        out(n++, Rm.m[2]); out(n++, Rm.m[5]); out(n++, Rm.m[8]);
    } else {
        out(n++, pry ? p.pitch : p.roll); out(n++, pry ? p.roll : p.pitch);
    }
    if (id != 13) out(n++, hd);
    if (global) { out(n++, s.vel.x); out(n++, s.vel.y); out(n++, s.vel.z); }
    else { out(n++, lvel.x); out(n++, lvel.y); out(n++, lvel.z); }
    out(n++, s.om.x); out(n++, s.om.y); out(n++, s.om.z);
    switch (id) {
    case 3: case 4:
        out(n++, s16); out(n++, s17); out(n++, s18);
        if (PEND) { for (int k = 0; k < 4; k++) out(n++, s.act[k]); }
        else { out(n++, s.act[0]); out(n++, s.act[1]); out(n++, s.act[2]); out(n++, s.act[3]); }
        out(n++, s13); out(n++, s12); out(n++, s14); out(n++, s15); break;
    case 5: out(n++, s16); out(n++, s17); out(n++, s18); out(n++, s13); out(n++, s12); out(n++, s14); out(n++, s15); break;
    case 7: out(n++, s13); out(n++, s12); out(n++, s16); out(n++, s17); out(n++, s18); out(n++, s14); out(n++, s15); break;
    case 11: out(n++, s16); out(n++, s17); out(n++, s18); break;
    case 2: case 6: out(n++, s13); out(n++, s12); out(n++, s14); out(n++, s15); break;
    default: out(n++, s12); out(n++, s13); out(n++, s14); out(n++, s15); break;   // 1, 8, 9, 10, 13, 14
    }
    if (id == 6 || id == 7 || id == 8 || id == 13) {
        #pragma unroll
        for (int k = 0; k < 6; k++) out(n++, prm[k]);
    } else if (id == 9) {                                                  // LocalFrameRPYFakeParamsEnv :328
        out(n++, T(1)); out(n++, T(0.17)); out(n++, T(7)); out(n++, T(0.01)); out(n++, T(1.2)); out(n++, T(0.3));
    }
}

}  // namespace dsim
