// dsim_step.cuh — the fused env-step kernel.  Included by dsim_kernels.cu inside its anonymous namespace.
//
// BaseDroneEnv.vector_step (BaseDroneEnv.py:259-294): ctrl remap (:269) -> mj_step x frame_skip
// (mujoco_vecenv.py:404-407) -> num_steps += 1 (:271) -> get_drone_states (:357-380) -> terminated_fcn / reward_fcn
// (:275-284) -> _get_obs (observation_wrappers.py), plus the RLlib reset_at() round trip (:334-351) folded in
// when auto_reset is on.
//
// Memory plan per CTA of kStepBlock envs (one env per thread):
//   * SoA state / constant rows: coalesced 4-byte loads, all issued up front;
//   * operands used only AFTER the physics (step counter, episode return, raw params, per-env setpoint) go
//     global -> shared with cp.async at kernel entry: no registers held, DRAM latency hidden behind the physics;
//   * the [N][D] policy-ready observation rows are written to a padded smem tile (odd pitch, conflict-free) and leave
//     the SM as coalesced 16-byte vectors; a per-thread strided store would cost 32 L1 wavefronts per value;
//   * rare paths (Philox re-sampling) are one out-of-line call so they do not inflate the hot path's registers/I-cache.
#pragma once

#ifndef DSIM_BLOCK
#define DSIM_BLOCK 128
#endif
#ifndef DSIM_MINB
#define DSIM_MINB 4
#endif
constexpr int kStepBlock = DSIM_BLOCK;

// ------------------------------------------------------------------ kernel parameter block (constant bank)
template <typename T> struct KParams {
    int n, ld;
    T *state;
    int *num_steps;
    unsigned *reset_count;
    const T *consts;          // [13][ld]
    const T *params;          // [6][ld]
    T *ref_env;               // [4][ld] (per-env setpoints) or nullptr
    T *obs, *reward, *ep_return;
    unsigned char *trunc;
    double *stats;
    const T *actions;         // [n][4]
    T uconst[C_ROWS];         // uniform-parameter fast path (random_params == False)
    T uparams[6];
    int per_env_consts, auto_reset, obs_id, reward_id, obs_layout, obs_dim, frame_skip, max_steps;
    int eval_only;            // 1: termination / reward / obs of the CURRENT state, nothing advanced or stored
    T h, max_distance_t;
    T ref_off[3], ref_yaw, start_t[3];
    double start[3], ref64[3], max_distance;
    ResetCfg<T> rc;
    unsigned seed, env_base;
};

template <typename T> DSIM_DEV EnvState<T> load_state(const KParams<T> &p, int i) {
    const T *b = p.state + i;
    const size_t ld = p.ld;
    EnvState<T> s;
    s.pos = mk(b[0 * ld], b[1 * ld], b[2 * ld]);
    s.qw = b[3 * ld]; s.qx = b[4 * ld]; s.qy = b[5 * ld]; s.qz = b[6 * ld];
    s.hx = b[7 * ld]; s.hy = b[8 * ld];
    s.vel = mk(b[9 * ld], b[10 * ld], b[11 * ld]);
    s.om = mk(b[12 * ld], b[13 * ld], b[14 * ld]);
    s.hvx = b[15 * ld]; s.hvy = b[16 * ld];
    #pragma unroll
    for (int k = 0; k < 4; k++) s.act[k] = b[(S_ACT + k) * ld];
    s.acc = mk(b[21 * ld], b[22 * ld], b[23 * ld]);
    return s;
}
template <typename T> DSIM_DEV void store_state(const KParams<T> &p, int i, const EnvState<T> &s) {
    T *b = p.state + i;
    const size_t ld = p.ld;
    b[0 * ld] = s.pos.x; b[1 * ld] = s.pos.y; b[2 * ld] = s.pos.z;
    b[3 * ld] = s.qw; b[4 * ld] = s.qx; b[5 * ld] = s.qy; b[6 * ld] = s.qz;
    b[7 * ld] = s.hx; b[8 * ld] = s.hy;
    b[9 * ld] = s.vel.x; b[10 * ld] = s.vel.y; b[11 * ld] = s.vel.z;
    b[12 * ld] = s.om.x; b[13 * ld] = s.om.y; b[14 * ld] = s.om.z;
    b[15 * ld] = s.hvx; b[16 * ld] = s.hvy;
    #pragma unroll
    for (int k = 0; k < 4; k++) b[(S_ACT + k) * ld] = s.act[k];
    b[21 * ld] = s.acc.x; b[22 * ld] = s.acc.y; b[23 * ld] = s.acc.z;
}
template <typename T> DSIM_DEV EnvConsts<T> load_consts(const KParams<T> &p, int i) {
    EnvConsts<T> c;
    T v[C_ROWS];
    if (p.per_env_consts) {
        #pragma unroll
        for (int k = 0; k < C_ROWS; k++) v[k] = __ldg(p.consts + (size_t)k * p.ld + i);
    } else {
        #pragma unroll
        for (int k = 0; k < C_ROWS; k++) v[k] = p.uconst[k];
    }
    c.mB = v[C_MB]; c.cz = v[C_CZ]; c.IBx = v[C_IBX]; c.IBy = v[C_IBY]; c.IBz = v[C_IBZ];
    c.mD = v[C_MD]; c.zD = v[C_ZD]; c.IDx = v[C_IDX]; c.IDz = v[C_IDZ];
    c.Fs = v[C_FS]; c.F = v[C_F]; c.kq = v[C_KQ]; c.inv_tau = v[C_INVTAU];
    return c;
}
template <typename T> DSIM_DEV void load_params(const KParams<T> &p, int i, T prm[6]) {
    if (p.per_env_consts) {
        #pragma unroll
        for (int k = 0; k < 6; k++) prm[k] = __ldg(p.params + (size_t)k * p.ld + i);
    } else {
        #pragma unroll
        for (int k = 0; k < 6; k++) prm[k] = p.uparams[k];
    }
}
template <typename T> DSIM_DEV void load_ref(const KParams<T> &p, int i, V3<T> &ref_off, T &ref_yaw, double ref64[3]) {
    if (p.ref_env) {
        const T *r = p.ref_env + i;
        ref_off = mk(r[0], r[(size_t)p.ld], r[2 * (size_t)p.ld]);
        ref_yaw = r[3 * (size_t)p.ld];
        ref64[0] = p.start[0] + (double)ref_off.x; ref64[1] = p.start[1] + (double)ref_off.y; ref64[2] = p.start[2] + (double)ref_off.z;
    } else {
        ref_off = mk(p.ref_off[0], p.ref_off[1], p.ref_off[2]);
        ref_yaw = p.ref_yaw;
        ref64[0] = p.ref64[0]; ref64[1] = p.ref64[1]; ref64[2] = p.ref64[2];
    }
}
template <typename T> DSIM_DEV bool state_finite(const EnvState<T> &s) {
    const T a = s.pos.x + s.pos.y + s.pos.z + s.qw + s.qx + s.qy + s.qz + s.hx + s.hy;
    const T b = s.vel.x + s.vel.y + s.vel.z + s.om.x + s.om.y + s.om.z + s.hvx + s.hvy + s.act[0] + s.act[1] + s.act[2] + s.act[3];
    // MuJoCo's mj_check* also rejects |x| > mjMAXVAL (1e10)
    return finite_(a) && finite_(b) && abs_(a) < T(1e10) && abs_(b) < T(1e10);
}
// observation component sink: `base[j * stride]`; STRIDE > 0 fixes the stride at compile time
template <typename T, int STRIDE = 0> struct ObsWriter {
    T *base; size_t stride;
    DSIM_DEV void operator()(int j, T v) const { if constexpr (STRIDE > 0) base[j * STRIDE] = v; else base[(size_t)j * stride] = v; }
};
template <typename T> DSIM_DEV ObsWriter<T> obs_writer(const KParams<T> &p, int i) {
    ObsWriter<T> w;
    if (p.obs_layout == DSIM_LAYOUT_SOA) { w.base = p.obs + i; w.stride = p.ld; }
    else { w.base = p.obs + (size_t)i * p.obs_dim; w.stride = 1; }
    return w;
}

constexpr int kObsPad = DSIM_MAX_OBS | 1;      // upper bound of the odd smem row pitch
constexpr int kLate = 11;                      // ep_return, params[6], ref[4]: needed only after the physics

// cooperative, fully coalesced copy of the CTA's observation tile (smem, row pitch Dp = D | 1) to obs[row0*D ...]
// (row pitch D).  flat element e lives at tile[e + (Dp - D) * (e / D)]; DC > 0 makes D a compile-time constant.
template <typename T, int DC>
DSIM_DEV void copy_out_obs(const T *tile, T *gout, int Drt, int nvalid, int t) {
    const int D = DC > 0 ? DC : Drt;
    const int pad = (D | 1) - D;
    const int E = nvalid * D;
    if constexpr (std::is_same<T, float>::value) {
        if ((E & 3) == 0) {                    // full CTAs: always; a ragged last CTA falls through to the scalar loop
            float4 *g4 = reinterpret_cast<float4 *>(gout);
            for (int e4 = t; e4 < (E >> 2); e4 += kStepBlock) {
                const int e = 4 * e4;
                float v[4];
                #pragma unroll
                for (int k = 0; k < 4; k++) v[k] = tile[(e + k) + pad * ((e + k) / D)];
                g4[e4] = make_float4(v[0], v[1], v[2], v[3]);
            }
            return;
        }
    }
    for (int e = t; e < E; e += kStepBlock) gout[e] = tile[e + pad * (e / D)];
}

template <typename T> constexpr int min_blocks() { return std::is_same<T, float>::value ? DSIM_MINB * 128 / DSIM_BLOCK : 1; }
__host__ __device__ constexpr int obs_dim_of(int obs_id) {
    return obs_id == 0 ? 33 : obs_id == 1 ? 16 : obs_id == 2 ? 16 : obs_id == 3 ? 23 : obs_id == 4 ? 24 : obs_id == 5 ? 19 : obs_id == 6 ? 22 :
           obs_id == 7 ? 25 : obs_id == 8 ? 22 : obs_id == 9 ? 22 : obs_id == 10 ? 16 : obs_id == 11 ? 15 : obs_id == 13 ? 28 : obs_id == 14 ? 17 : 0;
}

// OBS / REW >= 0 are compile-time specialisations of the wrapper class / reward function (smaller code, no dispatch
// branches, constant observation width); -1 reads the ids from the parameter block.
template <typename T, bool PEND, int OBS, int REW>
__global__ void __launch_bounds__(kStepBlock, min_blocks<T>()) step_kernel(const KParams<T> p) {
    __shared__ T s_obs[kStepBlock * kObsPad];
    __shared__ T s_late[kLate][kStepBlock];
    __shared__ int s_ns[kStepBlock];
    constexpr int DC = (OBS >= 0 && PEND) ? obs_dim_of(OBS) : 0;
    const int t = threadIdx.x, i = blockIdx.x * kStepBlock + t;
    const bool active = i < p.n;
    const int obs_id = OBS >= 0 ? OBS : p.obs_id, reward_id = REW >= 0 ? REW : p.reward_id;
    const int D = DC > 0 ? DC : p.obs_dim, Dp = D | 1;
    const bool staged = p.obs_layout == DSIM_LAYOUT_ENV_MAJOR;
    if (active) {
        __pipeline_memcpy_async(&s_ns[t], p.num_steps + i, sizeof(int));
        __pipeline_memcpy_async(&s_late[0][t], p.ep_return + i, sizeof(T));
        if (p.per_env_consts) {
            #pragma unroll
            for (int k = 0; k < 6; k++) __pipeline_memcpy_async(&s_late[1 + k][t], p.params + (size_t)k * p.ld + i, sizeof(T));
        }
        if (p.ref_env) {
            #pragma unroll
            for (int k = 0; k < 4; k++) __pipeline_memcpy_async(&s_late[7 + k][t], p.ref_env + (size_t)k * p.ld + i, sizeof(T));
        }
        __pipeline_commit();

        EnvState<T> s = load_state(p, i);
        const EnvConsts<T> c = load_consts(p, i);
        T a[4], ctrl[4];
        if constexpr (std::is_same<T, float>::value) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(p.actions) + i);
            a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
        } else {
            const double2 v0 = __ldg(reinterpret_cast<const double2 *>(p.actions) + 2 * (size_t)i);
            const double2 v1 = __ldg(reinterpret_cast<const double2 *>(p.actions) + 2 * (size_t)i + 1);
            a[0] = v0.x; a[1] = v0.y; a[2] = v1.x; a[3] = v1.y;
        }
        #pragma unroll
        for (int k = 0; k < 4; k++) ctrl[k] = clamp_(T(0.1) + T(0.9) * a[k], T(0), T(1));   // :269 + ctrlrange (0,1) clamp of mj_fwdActuation
        for (int f = 0; f < p.frame_skip; f++) substep<T, PEND, true>(s, c, ctrl, p.h);

        __pipeline_wait_prior(0);
        int ns = s_ns[t] + (p.eval_only ? 0 : 1);
        const unsigned env = p.env_base + (unsigned)i;
        // MuJoCo's mj_checkPos/Vel/Acc warn and reset the whole MjData; here the one env is parked on a finite state for
        // the outputs of this step, flagged truncated, counted, and re-sampled below.  Never silent.
        const bool bad = !state_finite(s);
        if (bad) {
            s.pos = mk(T(0), T(0), T(0)); s.qw = T(1); s.qx = s.qy = s.qz = T(0); s.hx = s.hy = s.hvx = s.hvy = T(0);
            s.vel = mk(T(0), T(0), T(0)); s.om = mk(T(0), T(0), T(0)); s.acc = mk(T(0), T(0), T(0));
            #pragma unroll
            for (int k = 0; k < 4; k++) s.act[k] = T(0);
        }

        V3<T> ref_off; T ref_yaw; double ref64[3];
        if (p.ref_env) {
            ref_off = mk(s_late[7][t], s_late[8][t], s_late[9][t]);
            ref_yaw = s_late[10][t];
            ref64[0] = p.start[0] + (double)ref_off.x; ref64[1] = p.start[1] + (double)ref_off.y; ref64[2] = p.start[2] + (double)ref_off.z;
        } else {
            ref_off = mk(p.ref_off[0], p.ref_off[1], p.ref_off[2]);
            ref_yaw = p.ref_yaw;
            ref64[0] = p.ref64[0]; ref64[1] = p.ref64[1]; ref64[2] = p.ref64[2];
        }
        T prm[6];
        #pragma unroll
        for (int k = 0; k < 6; k++) prm[k] = p.per_env_consts ? s_late[1 + k][t] : p.uparams[k];
        const PostState<T> ps = post_state(s, ref_off, ref_yaw);
        bool trunc = terminated(s.pos, p.start, ref64, p.max_distance, ns, p.max_steps) || bad;
        const T rew = bad ? T(0) : reward_fn<T, PEND>(reward_id, s, ps, a, ns, prm, p.max_distance_t);
        const V3<T> start_t = mk(p.start_t[0], p.start_t[1], p.start_t[2]);
        if (staged) {
            ObsWriter<T, 1> w; w.base = s_obs + t * Dp; w.stride = 1;
            emit_obs<T, PEND>(obs_id, s, ps, start_t, ref_off, prm, w);
        } else {
            ObsWriter<T> w; w.base = p.obs + i; w.stride = p.ld;
            emit_obs<T, PEND>(obs_id, s, ps, start_t, ref_off, prm, w);
        }
        p.reward[i] = rew;
        p.trunc[i] = trunc ? 1 : 0;

        if (!p.eval_only) {
            // would-be ground contact (the floor plane is out of reach in the BASELINE configs; detected, never ignored)
            if (p.start_t[2] + s.pos.z < prm[4] + T(0.5)) atomicAdd(p.stats + 4, 1.0);
            T ret = s_late[0][t] + rew;
            if (trunc) {
                atomicAdd(p.stats + 0, (double)ret); atomicAdd(p.stats + 1, (double)ns); atomicAdd(p.stats + 2, 1.0);
                if (bad) atomicAdd(p.stats + 3, 1.0);
                ret = T(0);
                if (p.auto_reset || bad) {   // native loop: the RLlib reset_at() round trip (:334-351) folded into the step
                    const unsigned rcnt = p.reset_count[i] + 1u;
                    sample_state<T, PEND>(s, p.rc, p.seed, env, rcnt);
                    p.reset_count[i] = rcnt;
                    ns = 0;
                }
            }
            p.ep_return[i] = ret;
            p.num_steps[i] = ns;
            store_state(p, i, s);
        }
    }
    if (staged) {
        __syncthreads();
        const int row0 = blockIdx.x * kStepBlock;
        copy_out_obs<T, DC>(s_obs, p.obs + (size_t)row0 * D, D, min(kStepBlock, p.n - row0), t);
    }
}
