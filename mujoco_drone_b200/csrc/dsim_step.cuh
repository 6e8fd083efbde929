// dsim_step.cuh — the fused env-step kernel.  Included by dsim_kernels.cu inside its anonymous namespace.
//
// BaseDroneEnv.vector_step (BaseDroneEnv.py:259-294): ctrl remap (:269) -> mj_step x frame_skip
// (mujoco_vecenv.py:404-407) -> num_steps += 1 (:271) -> get_drone_states (:357-380) -> terminated_fcn / reward_fcn
// (:275-284) -> _get_obs (observation_wrappers.py), plus the RLlib reset_at() round trip (:334-351) folded in
// when auto_reset is on.
//
// Execution plan (sm_100a): one WARP owns one page of 32 envs, one env per lane; warps never synchronise with
// each other.
//   * lane 0 arms the warp's mbarrier and issues 1D bulk copies (cp.async.bulk, the TMA engine: SASS UBLKCP) that
//     bring the warp's read-write page (state + step counter + episode return), read-only page (compiled constants +
//     raw parameters) and setpoint page from HBM into the warp's shared-memory slot: three instructions move ~6 KB,
//     no per-row address arithmetic, no registers held while the data is in flight;
//   * every lane reads its own column with immediate-offset LDS (row pitch = 32 lanes: conflict-free), runs the
//     physics / termination / reward / observation in registers, and writes the new column and its policy-ready
//     observation row back into the slot;
//   * after fence.proxy.async + __syncwarp lane 0 sends the read-write page and the warp's contiguous [32][obs_dim]
//     observation block back with two bulk stores and waits only for their shared-memory reads.
//   * rare paths (Philox re-sampling) are one out-of-line call so they do not inflate the hot path's registers/I-cache.
#pragma once

#ifndef DSIM_BLOCK
#define DSIM_BLOCK 128
#endif
#ifndef DSIM_MINB
#define DSIM_MINB 4
#endif
constexpr int kStepBlock = DSIM_BLOCK;
constexpr int kStepWarps = kStepBlock / 32;

// ------------------------------------------------------------------ async-proxy primitives (PTX ISA 8.x, sm_90+)
DSIM_DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
DSIM_DEV void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");      // make the init visible to the async proxy
}
DSIM_DEV void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DSIM_DEV void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// HBM -> shared, completion counted in bytes on the mbarrier.  size % 16 == 0, both addresses 16-byte aligned.
DSIM_DEV void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> HBM, tracked by the issuing thread's bulk async-group
DSIM_DEV void bulk_s2g(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
DSIM_DEV void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
DSIM_DEV void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (the bulk store that follows)
DSIM_DEV void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------ kernel parameter block (constant bank)
template <typename T> struct KParams {
    int n, npages;
    T *rw;                    // [npages][RW_ROWS][32]
    const T *ro;              // [npages][RO_ROWS][32]
    T *refp;                  // [npages][REF_ROWS][32] (per-env setpoints) or nullptr
    unsigned *reset_count;    // [npages * 32]
    T *obs, *reward;          // [n][obs_dim], [n]
    unsigned char *trunc;
    double *stats;
    const T *actions;         // [n][4]
    T uconst[C_ROWS];         // uniform-parameter fast path (random_params == False)
    T uparams[6];
    int per_env_consts, auto_reset, obs_id, reward_id, obs_dim, frame_skip, max_steps;
    int eval_only;            // 1: termination / reward / obs of the CURRENT state, nothing advanced or stored
    unsigned smem_per_warp;   // bytes
    T h, max_distance_t;
    T ref_off[3], ref_yaw, start_t[3];
    double start[3], ref64[3], max_distance;
    ResetCfg<T> rc;
    unsigned seed, env_base;
};

// integer rows of the read-write page are stored in a lane-sized slot (int32 for float pages, int64 for double)
template <typename T> struct IntOf { typedef int type; };
template <> struct IntOf<double> { typedef long long type; };
template <typename T> DSIM_DEV int slot_to_int(T v) {
    if constexpr (std::is_same<T, float>::value) return __float_as_int(v); else return (int)__double_as_longlong(v);
}
template <typename T> DSIM_DEV T int_to_slot(int v) {
    if constexpr (std::is_same<T, float>::value) return __int_as_float(v); else return __longlong_as_double((long long)v);
}

// column accessors: `col` points at (row 0, this env) of a page, in shared or global memory; row pitch = kTile
template <typename T> DSIM_DEV EnvState<T> load_state(const T *col) {
    EnvState<T> s;
    s.pos = mk(col[0 * kTile], col[1 * kTile], col[2 * kTile]);
    s.qw = col[3 * kTile]; s.qx = col[4 * kTile]; s.qy = col[5 * kTile]; s.qz = col[6 * kTile];
    s.hx = col[7 * kTile]; s.hy = col[8 * kTile];
    s.vel = mk(col[9 * kTile], col[10 * kTile], col[11 * kTile]);
    s.om = mk(col[12 * kTile], col[13 * kTile], col[14 * kTile]);
    s.hvx = col[15 * kTile]; s.hvy = col[16 * kTile];
    #pragma unroll
    for (int k = 0; k < 4; k++) s.act[k] = col[(S_ACT + k) * kTile];
    s.acc = mk(col[21 * kTile], col[22 * kTile], col[23 * kTile]);
    return s;
}
template <typename T> DSIM_DEV void store_state(T *col, const EnvState<T> &s) {
    col[0 * kTile] = s.pos.x; col[1 * kTile] = s.pos.y; col[2 * kTile] = s.pos.z;
    col[3 * kTile] = s.qw; col[4 * kTile] = s.qx; col[5 * kTile] = s.qy; col[6 * kTile] = s.qz;
    col[7 * kTile] = s.hx; col[8 * kTile] = s.hy;
    col[9 * kTile] = s.vel.x; col[10 * kTile] = s.vel.y; col[11 * kTile] = s.vel.z;
    col[12 * kTile] = s.om.x; col[13 * kTile] = s.om.y; col[14 * kTile] = s.om.z;
    col[15 * kTile] = s.hvx; col[16 * kTile] = s.hvy;
    #pragma unroll
    for (int k = 0; k < 4; k++) col[(S_ACT + k) * kTile] = s.act[k];
    col[21 * kTile] = s.acc.x; col[22 * kTile] = s.acc.y; col[23 * kTile] = s.acc.z;
}
template <typename T> DSIM_DEV EnvConsts<T> consts_from(const T v[C_ROWS]) {
    EnvConsts<T> c;
    c.mB = v[C_MB]; c.cz = v[C_CZ]; c.IBx = v[C_IBX]; c.IBy = v[C_IBY]; c.IBz = v[C_IBZ];
    c.mD = v[C_MD]; c.zD = v[C_ZD]; c.IDx = v[C_IDX]; c.IDz = v[C_IDZ];
    c.Fs = v[C_FS]; c.F = v[C_F]; c.kq = v[C_KQ]; c.inv_tau = v[C_INVTAU];
    return c;
}
// `ro_col`: (row 0, this env) of the read-only page, shared or global
template <typename T> DSIM_DEV EnvConsts<T> load_consts(const KParams<T> &p, const T *ro_col) {
    T v[C_ROWS];
    #pragma unroll
    for (int k = 0; k < C_ROWS; k++) v[k] = p.per_env_consts ? ro_col[(RO_CONSTS + k) * kTile] : p.uconst[k];
    return consts_from(v);
}
template <typename T> DSIM_DEV void load_params(const KParams<T> &p, const T *ro_col, T prm[6]) {
    #pragma unroll
    for (int k = 0; k < 6; k++) prm[k] = p.per_env_consts ? ro_col[(RO_PARAMS + k) * kTile] : p.uparams[k];
}
// `ref_col`: (row 0, this env) of the setpoint page (ignored when the reference is shared)
template <typename T> DSIM_DEV void load_ref(const KParams<T> &p, const T *ref_col, V3<T> &ref_off, T &ref_yaw, double ref64[3]) {
    if (p.refp) {
        ref_off = mk(ref_col[0], ref_col[kTile], ref_col[2 * kTile]);
        ref_yaw = ref_col[3 * kTile];
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

template <typename T> constexpr int min_blocks() { return std::is_same<T, float>::value ? DSIM_MINB * 128 / DSIM_BLOCK : 1; }
__host__ __device__ constexpr int obs_dim_of(int obs_id) {
    return obs_id == 0 ? 33 : obs_id == 1 ? 16 : obs_id == 2 ? 16 : obs_id == 3 ? 23 : obs_id == 4 ? 24 : obs_id == 5 ? 19 : obs_id == 6 ? 22 :
           obs_id == 7 ? 25 : obs_id == 8 ? 22 : obs_id == 9 ? 22 : obs_id == 10 ? 16 : obs_id == 11 ? 15 : obs_id == 13 ? 28 : obs_id == 14 ? 17 : 0;
}
// shared-memory slot of one warp: [RW page | RO page | setpoint page | observation block [32][obs_dim]]
constexpr int kSlotObsOff = (RW_ROWS + RO_ROWS + REF_ROWS) * kTile;          // elements
__host__ __device__ constexpr unsigned slot_bytes(int obs_dim, int elem) {
    return (unsigned)((((kSlotObsOff + kTile * obs_dim) * elem) + 127) / 128 * 128);
}

// rare path, out of line: RLlib's reset_at() round trip (:334-351) folded into the step.  Works on the env's column of
// the read-write page in shared memory (qpos / qvel rows and the step counter), so the hot path keeps no state
// registers alive across the call.
template <typename T, bool PEND>
__device__ __noinline__ void resample_column(T *col, const ResetCfg<T> &rc, unsigned seed, unsigned env, unsigned *reset_count_slot) {
    const unsigned rcnt = *reset_count_slot + 1u;
    EnvState<T> s;
    sample_state<T, PEND>(s, rc, seed, env, rcnt);
    *reset_count_slot = rcnt;
    col[0 * kTile] = s.pos.x; col[1 * kTile] = s.pos.y; col[2 * kTile] = s.pos.z;
    col[3 * kTile] = s.qw; col[4 * kTile] = s.qx; col[5 * kTile] = s.qy; col[6 * kTile] = s.qz;
    col[7 * kTile] = s.hx; col[8 * kTile] = s.hy;
    col[9 * kTile] = s.vel.x; col[10 * kTile] = s.vel.y; col[11 * kTile] = s.vel.z;
    col[12 * kTile] = s.om.x; col[13 * kTile] = s.om.y; col[14 * kTile] = s.om.z;
    col[15 * kTile] = s.hvx; col[16 * kTile] = s.hvy;
    col[RW_NUM_STEPS * kTile] = int_to_slot<T>(0);
}

// OBS / REW >= 0 are compile-time specialisations of the wrapper class / reward function (smaller code, no dispatch
// branches, constant observation width); -1 reads the ids from the parameter block.
template <typename T, bool PEND, int OBS, int REW>
__global__ void __launch_bounds__(kStepBlock, min_blocks<T>()) step_kernel(const __grid_constant__ KParams<T> p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar[kStepWarps];
    constexpr int DC = (OBS >= 0 && PEND) ? obs_dim_of(OBS) : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int page = blockIdx.x * kStepWarps + warp;
    if (page >= p.npages) return;                                  // warps are autonomous: no CTA-wide barrier below
    const int i = page * kTile + lane;
    const bool active = i < p.n;                                   // pad lanes of the last page compute, but publish nothing
    const int obs_id = OBS >= 0 ? OBS : p.obs_id, reward_id = REW >= 0 ? REW : p.reward_id;
    const int D = DC > 0 ? DC : p.obs_dim;
    T *slot = reinterpret_cast<T *>(smem_raw + (size_t)warp * p.smem_per_warp);
    T *s_rw = slot, *s_ro = slot + RW_ROWS * kTile, *s_ref = slot + (RW_ROWS + RO_ROWS) * kTile, *s_obs = slot + kSlotObsOff;
    uint64_t *bar = &s_bar[warp];

    // ---- page loads: HBM -> this warp's slot
    if (lane == 0) {
        mbar_init(bar, 1);
        constexpr uint32_t rwb = RW_ROWS * kTile * sizeof(T), rob = RO_ROWS * kTile * sizeof(T), rfb = REF_ROWS * kTile * sizeof(T);
        mbar_arrive_expect_tx(bar, rwb + (p.per_env_consts ? rob : 0u) + (p.refp ? rfb : 0u));
        bulk_g2s(s_rw, p.rw + (size_t)page * (RW_ROWS * kTile), rwb, bar);
        if (p.per_env_consts) bulk_g2s(s_ro, p.ro + (size_t)page * (RO_ROWS * kTile), rob, bar);
        if (p.refp) bulk_g2s(s_ref, p.refp + (size_t)page * (REF_ROWS * kTile), rfb, bar);
    }
    T a[4] = {T(0), T(0), T(0), T(0)};
    if (active) {
        if constexpr (std::is_same<T, float>::value) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(p.actions) + i);
            a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
        } else {
            const double2 v0 = __ldg(reinterpret_cast<const double2 *>(p.actions) + 2 * (size_t)i);
            const double2 v1 = __ldg(reinterpret_cast<const double2 *>(p.actions) + 2 * (size_t)i + 1);
            a[0] = v0.x; a[1] = v0.y; a[2] = v1.x; a[3] = v1.y;
        }
    }
    __syncwarp();                                                  // barrier init visible to the waiting lanes
    mbar_wait(bar, 0);

    // ---- physics
    T *col = s_rw + lane;
    const T *ro_col = s_ro + lane;
    EnvState<T> s = load_state(col);
    {
        const EnvConsts<T> c = load_consts(p, ro_col);
        T ctrl[4];
        #pragma unroll
        for (int k = 0; k < 4; k++) ctrl[k] = clamp_(T(0.1) + T(0.9) * a[k], T(0), T(1));   // :269 + ctrlrange (0,1) clamp of mj_fwdActuation
        #pragma unroll 1
        for (int f = 0; f < p.frame_skip; f++) substep<T, PEND, true>(s, c, ctrl, p.h);
    }

    // ---- counters, termination, reward, observation
    int ns = slot_to_int(col[RW_NUM_STEPS * kTile]) + (p.eval_only ? 0 : 1);
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
    load_ref(p, s_ref + lane, ref_off, ref_yaw, ref64);
    T prm[6];
    load_params(p, ro_col, prm);
    const PostState<T> ps = post_state(s, ref_off, ref_yaw);
    const bool trunc = terminated(s.pos, p.start, ref64, p.max_distance, ns, p.max_steps) || bad;
    const T rew = bad ? T(0) : reward_fn<T, PEND>(reward_id, s, ps, a, ns, prm, p.max_distance_t);
    {
        ObsWriter<T, 1> w; w.base = s_obs + lane * D; w.stride = 1;
        emit_obs<T, PEND>(obs_id, s, ps, mk(p.start_t[0], p.start_t[1], p.start_t[2]), ref_off, prm, w);
    }
    if (active) {
        p.reward[i] = rew;
        p.trunc[i] = trunc ? 1 : 0;
    }

    if (!p.eval_only) {
        // would-be ground contact (the floor plane is out of reach in the BASELINE configs; detected, never ignored)
        if (active && p.start_t[2] + s.pos.z < prm[4] + T(0.5)) atomicAdd(p.stats + 4, 1.0);
        T ret = col[RW_EP_RETURN * kTile] + rew;
        if (trunc && active) {
            atomicAdd(p.stats + 0, (double)ret); atomicAdd(p.stats + 1, (double)ns); atomicAdd(p.stats + 2, 1.0);
            if (bad) atomicAdd(p.stats + 3, 1.0);
        }
        col[RW_EP_RETURN * kTile] = trunc ? T(0) : ret;
        col[RW_NUM_STEPS * kTile] = int_to_slot<T>(ns);
        store_state(col, s);
        if (trunc && (p.auto_reset || bad)) resample_column<T, PEND>(col, p.rc, p.seed, env, p.reset_count + i);
    }

    // ---- publish: slot -> HBM
    fence_async_smem();
    __syncwarp();
    const int nvalid = min(kTile, p.n - page * kTile);
    const uint32_t obs_bytes = (uint32_t)(nvalid * D) * (uint32_t)sizeof(T);
    T *gobs = p.obs + (size_t)page * kTile * D;
    const bool obs_bulk = (obs_bytes & 15u) == 0;                  // always true for full pages
    if (lane == 0) {
        if (!p.eval_only) bulk_s2g(p.rw + (size_t)page * (RW_ROWS * kTile), s_rw, RW_ROWS * kTile * sizeof(T));
        if (obs_bulk) bulk_s2g(gobs, s_obs, obs_bytes);
        bulk_commit();
    }
    if (!obs_bulk)                                                 // ragged last page whose byte count is not a multiple of 16
        for (int e = lane; e < nvalid * D; e += kTile) gobs[e] = s_obs[e];
    if (lane == 0) bulk_wait_read();                               // the slot must outlive the bulk reads
}
