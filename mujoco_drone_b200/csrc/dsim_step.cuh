// dsim_step.cuh — the fused env-step kernel.  Included by dsim_kernels.cu inside its anonymous namespace.
//
// BaseDroneEnv.vector_step (BaseDroneEnv.py:259-294): ctrl remap (:269) -> mj_step x frame_skip
// (mujoco_vecenv.py:404-407) -> num_steps += 1 (:271) -> get_drone_states (:357-380) -> terminated_fcn / reward_fcn
// (:275-284) -> _get_obs (observation_wrappers.py), plus the RLlib reset_at() round trip (:334-351) folded in
// when auto_reset is on.
//
// Execution plan (sm_100a): one WARP owns one page of 32 envs at a time, one env per lane; warps never synchronise with
// each other (no __syncthreads).
//   * PERSISTENT + WORK STEALING: the grid holds as many CTAs as the GPU keeps resident; a warp starts on page `wid` and
//     draws further pages from a global ticket counter (asked for after the physics of the current page).
//   * lane 0 arms a slot's mbarrier and issues 1D bulk copies (cp.async.bulk, the TMA engine: SASS UBLKCP) that bring the
//     page's read-write rows (state, step counter, episode return, reset count), read-only rows (compiled constants + raw
//     parameters), setpoint rows and its 512 bytes of actions from HBM into one of the warp's TWO shared-memory slots:
//     four instructions move ~7 KB, no per-row address arithmetic, no registers held while data is in flight, and the
//     next page is in flight while the current one is computed;
//   * every lane reads its own column with immediate-offset LDS (row pitch = 32 lanes: conflict-free), runs the
//     physics / termination / reward / observation in registers, and writes the new column and its policy-ready
//     observation row back into the slot (the observation block overlays the dead read-only operands);
//   * after fence.proxy.async + __syncwarp lane 0 sends the read-write page and the warp's contiguous [32][obs_dim]
//     observation block back with two bulk stores; their shared-memory reads are waited for only when the slot is reused;
//   * truncated envs are re-sampled by the whole warp, four at a time (resample_page), out of line so the rare path does
//     not inflate the hot path's registers / I-cache;
//   * programmatic dependent launch overlaps the next step's launch latency with this step's tail.
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
// L2 eviction priorities for the bulk copies (createpolicy): the read-only pages (compiled constants + raw parameters, 76 B
// per env) are the only data a step re-reads unchanged next step, so they are fetched evict_last and stay in the 126 MB L2
// across steps even when the streaming data (state pages, actions, observations: evict_first) is far larger than L2 -
// at 524288 envs the 40 MB of constants stay resident while 170 MB stream past them each step.
#ifndef DSIM_L2HINT
#define DSIM_L2HINT 1
#endif
DSIM_DEV uint64_t l2_policy_keep() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
DSIM_DEV uint64_t l2_policy_stream() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
// HBM -> shared, completion counted in bytes on the mbarrier.  size % 16 == 0, both addresses 16-byte aligned.
DSIM_DEV void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
DSIM_DEV void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
#if DSIM_L2HINT
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
#else
    (void)policy; bulk_g2s(smem_dst, gmem_src, bytes, bar);
#endif
}
// shared -> HBM, tracked by the issuing thread's bulk async-group
DSIM_DEV void bulk_s2g(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
DSIM_DEV void bulk_s2g(void *gmem_dst, const void *smem_src, uint32_t bytes, uint64_t policy) {
#if DSIM_L2HINT
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes), "l"(policy) : "memory");
#else
    (void)policy; bulk_s2g(gmem_dst, smem_src, bytes);
#endif
}
DSIM_DEV void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
DSIM_DEV void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (the bulk store that follows)
DSIM_DEV void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------ kernel parameter block (constant bank)
// episode statistics are accumulated into kStatReplicas copies of the 8 counters (warp w adds to copy w % kStatReplicas,
// dsim_stats sums them): every truncation costs three FP64 atomics, ~3400 per C4 step in the steady state, and atomics on
// ONE address retire one after the other on the L2 slice that owns it
#ifndef DSIM_STAT_REPLICAS
#define DSIM_STAT_REPLICAS 64
#endif
constexpr int kStatReplicas = DSIM_STAT_REPLICAS;
template <typename T> struct KParams {
    int n, npages;            // npages: one past the last page of this launch
    int page0;                // first page of this launch (0 unless the host entry point steps the batch in chunks)
    T *rw;                    // [npages][RW_ROWS][32]
    const T *ro;              // [npages][RO_ROWS][32]
    T *refp;                  // [npages][REF_ROWS][32] (per-env setpoints) or nullptr
    T *obs, *reward;          // [n][obs_dim], [n]
    unsigned char *trunc;
    double *stats;            // [kStatReplicas][8]
    const T *actions;         // [n][4]
    T uconst[C_ROWS];         // uniform-parameter fast path (random_params == False)
    T uparams[6];
    int per_env_consts, auto_reset, obs_id, reward_id, obs_dim, frame_skip, max_steps;
    int eval_only;            // 1: termination / reward / obs of the CURRENT state, nothing advanced or stored
    int early_ro;             // 1: the read-only rows may be fetched before the dependency wait (no kernel that writes them is in flight)
    int early_in;             // 1: so may the state / action / setpoint rows (dsim_set_inputs_ready: the caller guarantees that the kernel
                              //    running ahead of this one in the stream does not write them - e.g. it steps another shard)
    unsigned smem_per_slot;   // bytes of one page slot (each warp owns kStages of them)
    T h, max_distance_t;
    T ref_off[3], ref_yaw, start_t[3];
    double start[3], ref64[3], max_d2;         // max_d2: see terminated()
    ResetCfg<T> rc;
    unsigned seed, env_base;
    unsigned *ticket;               // [0] work-stealing page-chunk counter, [1] finished-warp counter (both 0 between launches)
    unsigned long long *timeline;   // debug: [gridDim * warps][8] %globaltimer stamps of the last launch, or nullptr
    // host entry point with pinned buffers: the mapped host copies of the outputs, written by the kernel itself next to the
    // device-resident ones (posted PCIe writes from the same bulk stores; no copy-engine pass), or nullptr
    T *obs_host, *reward_host;
    unsigned char *trunc_host;
    // floor contact (DsimConfig.ground_contact): per-env collision geometry [GEO_ROWS][ld] (geometry_kernel)
    const T *geo;
    int ld, pendulum, ground;
};

// integer rows of the read-write page are stored in a lane-sized slot (int32 for float pages, int64 for double)
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
    s.acc = mk(col[S_ACC * kTile], col[(S_ACC + 1) * kTile], col[(S_ACC + 2) * kTile]);
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
    col[S_ACC * kTile] = s.acc.x; col[(S_ACC + 1) * kTile] = s.acc.y; col[(S_ACC + 2) * kTile] = s.acc.z;
}
template <typename T> DSIM_DEV EnvConsts<T> consts_from(const T v[C_ROWS]) {
    EnvConsts<T> c;
    c.mB = v[C_MB]; c.cz = v[C_CZ]; c.IBx = v[C_IBX]; c.IBy = v[C_IBY]; c.IBz = v[C_IBZ];
    c.mD = v[C_MD]; c.zD = v[C_ZD]; c.IDx = v[C_IDX]; c.IDz = v[C_IDZ];
    c.Fs = v[C_FS]; c.F = v[C_F]; c.kq = v[C_KQ]; c.inv_tau = v[C_INVTAU];
    return c;
}
// `ro_col`: (row 0, this env) of the read-only page, shared or global
template <typename T> DSIM_DEV EnvConsts<T> load_consts(const KParams<T> &p, const T *ro_col, bool per_env) {
    T v[C_ROWS];
    #pragma unroll
    for (int k = 0; k < C_ROWS; k++) v[k] = per_env ? ro_col[(RO_CONSTS + k) * kTile] : p.uconst[k];
    return consts_from(v);
}
template <typename T> DSIM_DEV void load_params(const KParams<T> &p, const T *ro_col, T prm[6], bool per_env) {
    #pragma unroll
    for (int k = 0; k < 6; k++) prm[k] = per_env ? ro_col[(RO_PARAMS + k) * kTile] : p.uparams[k];
}
// `ref_col`: (row 0, this env) of the setpoint page (ignored when the reference is shared)
template <typename T> DSIM_DEV void load_ref(const KParams<T> &p, const T *ref_col, V3<T> &ref_off, T &ref_yaw, double ref64[3], bool per_env) {
    if (per_env) {
        ref_off = mk(ref_col[0], ref_col[kTile], ref_col[2 * kTile]);
        ref_yaw = ref_col[3 * kTile];
        ref64[0] = p.start[0] + (double)ref_off.x; ref64[1] = p.start[1] + (double)ref_off.y; ref64[2] = p.start[2] + (double)ref_off.z;
    } else {
        ref_off = mk(p.ref_off[0], p.ref_off[1], p.ref_off[2]);
        ref_yaw = p.ref_yaw;
        ref64[0] = p.ref64[0]; ref64[1] = p.ref64[1]; ref64[2] = p.ref64[2];
    }
}
// mj_checkPos / mj_checkVel: any NaN / Inf / |x| > mjMAXVAL (1e10) in qpos, qvel, act.  One sum of magnitudes: NaN and Inf
// propagate, nothing cancels, and a NaN fails the comparison.
template <typename T> DSIM_DEV bool state_finite(const EnvState<T> &s) {
    const T a = abs_(s.pos.x) + abs_(s.pos.y) + abs_(s.pos.z) + abs_(s.qw) + abs_(s.qx) + abs_(s.qy) + abs_(s.qz) + abs_(s.hx) + abs_(s.hy)
              + abs_(s.vel.x) + abs_(s.vel.y) + abs_(s.vel.z) + abs_(s.om.x) + abs_(s.om.y) + abs_(s.om.z) + abs_(s.hvx) + abs_(s.hvy)
              + abs_(s.act[0]) + abs_(s.act[1]) + abs_(s.act[2]) + abs_(s.act[3]);
    return a < T(1e10);
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
// shared-memory slot of one page: [RW page | union(RO page + setpoint page + actions [32][4], observation block [32][obs_dim])].
// The observation block overlays the read-only operands: they are dead (held in registers) when it is written.
constexpr int kSlotObsOff = RW_ROWS * kTile;                                  // elements
constexpr int kSlotActOff = (RW_ROWS + RO_ROWS + REF_ROWS) * kTile;
__host__ __device__ constexpr unsigned slot_bytes(int obs_dim, int elem) {
    const int in_elems = (RO_ROWS + REF_ROWS + 4) * kTile, out_elems = kTile * obs_dim;
    return (unsigned)((((kSlotObsOff + (in_elems > out_elems ? in_elems : out_elems)) * elem) + 127) / 128 * 128);
}
constexpr int kStages = 2;                                                    // slots per warp: compute in one, prefetch into the other

// rare path, out of line: RLlib's reset_at() round trip (:334-351) folded into the step, WARP-COOPERATIVE.  A page
// has 0-2 truncated envs in most steps (and many right after a synchronised start); letting each of them run sample_state
// on its own lane would cost the whole warp ~1000 divergent instructions per reset.  Instead the warp handles up to FOUR
// truncated envs per pass: lane 8 e + q draws the Philox block and computes Box-Muller pair q (0..7) of env e (0..3), the
// sixteen normals + two uniforms of each env are gathered with shuffles by the env's own lane, which assembles the state
// and writes its column (qpos / qvel rows, step counter) of the read-write page in shared memory.
// Same draws, same arithmetic as sample_state (dsim_device.cuh) = BaseDroneEnv.sample_state (:218-257).
template <typename T> DSIM_DEV T shfl_(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <typename T, bool PEND>
__device__ __noinline__ void resample_page(unsigned need, T *s_rw, const ResetCfg<T> &rc, unsigned seed, unsigned env0) {
    const int lane = threadIdx.x & 31, e = lane >> 3, q = lane & 7;
    if (__popc(need) > 12) {
        // mass truncation (a synchronised start, setpoints that ran away from the drones): more than three cooperative passes
        // cost more than letting every truncated lane draw its own state at once
        if ((need >> lane) & 1u) {
            EnvState<T> s;
            const unsigned rcnt = (unsigned)slot_to_int(s_rw[RW_RESET_COUNT * kTile + lane]) + 1u;
            sample_state<T, PEND>(s, rc, seed, env0 + (unsigned)lane, rcnt);
            s_rw[RW_RESET_COUNT * kTile + lane] = int_to_slot<T>((int)rcnt);
            T *col = s_rw + lane;
            col[0 * kTile] = s.pos.x; col[1 * kTile] = s.pos.y; col[2 * kTile] = s.pos.z;
            col[3 * kTile] = s.qw; col[4 * kTile] = s.qx; col[5 * kTile] = s.qy; col[6 * kTile] = s.qz;
            col[7 * kTile] = s.hx; col[8 * kTile] = s.hy;
            col[9 * kTile] = s.vel.x; col[10 * kTile] = s.vel.y; col[11 * kTile] = s.vel.z;
            col[12 * kTile] = s.om.x; col[13 * kTile] = s.om.y; col[14 * kTile] = s.om.z;
            col[15 * kTile] = s.hvx; col[16 * kTile] = s.hvy;
            col[RW_NUM_STEPS * kTile] = int_to_slot<T>(0);
        }
        __syncwarp();
        return;
    }
    while (need) {
        // the (up to) four lowest truncated lanes of this pass: P[k]; this lane computes for env P[e], and owns env `lane` if selected
        int P[4], mine = -1, own = -1;
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            P[k] = need ? __ffs(need) - 1 : -1;
            need &= need - 1;
            if (k == e) mine = P[k];
            if (P[k] == lane) own = k;
        }
        T z0 = T(0), z1 = T(0), rad = T(0), yw = T(0);
        unsigned rcnt = 0;
        if (mine >= 0) rcnt = (unsigned)slot_to_int(s_rw[RW_RESET_COUNT * kTile + mine]) + 1u;
        if (mine >= 0 && rc.random_start_pos) {
            // Box-Muller pair q: Philox block / word half of sample_state's draw order
            const int blk = q <= 1 ? 0 : q == 2 ? 1 : q <= 4 ? 2 : q <= 6 ? 3 : 4;
            const bool hi = q == 1 || q == 2 || q == 4 || q == 6;                 // (z, w) words instead of (x, y)
            const U4 x = philox4x32(blk, rcnt, 0, 0, seed, env0 + mine);
            box_muller(hi ? x.z : x.x, hi ? x.w : x.y, z0, z1);
            if (q == 2) { rad = rc.max_pos_offset * cbrt01(u01<T>(x.x)); yw = T(kPi) - T(2 * kPi) * u01<T>(x.y); }   // block 1, low words
        }
        const int src = 8 * (own >= 0 ? own : 0);
        const T n0 = shfl_(z0, src), n1 = shfl_(z1, src), n2 = shfl_(z0, src + 1);
        const T rp0 = shfl_(z0, src + 2), rp1 = shfl_(z1, src + 2), r = shfl_(rad, src + 2), yaw_s = shfl_(yw, src + 2);
        const T v0 = shfl_(z0, src + 3), v1 = shfl_(z1, src + 3), v2 = shfl_(z0, src + 4), w0 = shfl_(z1, src + 4);
        const T w1 = shfl_(z0, src + 5), w2 = shfl_(z1, src + 5), h0 = shfl_(z0, src + 6), h1 = shfl_(z1, src + 6);
        const T g0 = shfl_(z0, src + 7), g1 = shfl_(z1, src + 7);
        const unsigned rc_own = __shfl_sync(0xffffffffu, rcnt, src);
        if (own >= 0) {
            EnvState<T> s;
            s.pos = mk(T(0), T(0), T(0)); s.vel = mk(T(0), T(0), T(0)); s.om = mk(T(0), T(0), T(0));
            s.hx = s.hy = s.hvx = s.hvy = T(0);
            T roll = 0, pitch = 0, yaw = rc.start_yaw;
            if (rc.random_start_pos) {
                const T inn = rsqrt_(n0 * n0 + n1 * n1 + n2 * n2);
                s.pos = mk(r * (n0 * inn), r * (n1 * inn), r * (n2 * inn));
                yaw = yaw_s;
                roll = clipn(rp0, rc.angle_sigma[0]); pitch = clipn(rp1, rc.angle_sigma[1]);
                s.vel = mk(clipn(v0, rc.vel_sigma[0]), clipn(v1, rc.vel_sigma[1]), clipn(v2, rc.vel_sigma[2]));
                s.om = mk(clipn(w0, rc.ang_vel_sigma[0]), clipn(w1, rc.ang_vel_sigma[1]), clipn(w2, rc.ang_vel_sigma[2]));
                if (PEND) {
                    s.hx = clipn(h0, rc.pend_rp_sigma[0]); s.hy = clipn(h1, rc.pend_rp_sigma[1]);
                    s.hvx = clipn(g0, rc.pend_vel_sigma[0]); s.hvy = clipn(g1, rc.pend_vel_sigma[1]);
                }
            }
            rpy_to_quat(roll, pitch, yaw, s.qw, s.qx, s.qy, s.qz);
            s_rw[RW_RESET_COUNT * kTile + lane] = int_to_slot<T>((int)rc_own);
            T *col = s_rw + lane;
            col[0 * kTile] = s.pos.x; col[1 * kTile] = s.pos.y; col[2 * kTile] = s.pos.z;
            col[3 * kTile] = s.qw; col[4 * kTile] = s.qx; col[5 * kTile] = s.qy; col[6 * kTile] = s.qz;
            col[7 * kTile] = s.hx; col[8 * kTile] = s.hy;
            col[9 * kTile] = s.vel.x; col[10 * kTile] = s.vel.y; col[11 * kTile] = s.vel.z;
            col[12 * kTile] = s.om.x; col[13 * kTile] = s.om.y; col[14 * kTile] = s.om.z;
            col[15 * kTile] = s.hvx; col[16 * kTile] = s.hvy;
            col[RW_NUM_STEPS * kTile] = int_to_slot<T>(0);
        }
        __syncwarp();
    }
}

// this env's raw action row from the slot.  asm volatile: a fresh shared-memory read at each call site (never CSE'd into
// four registers that would live across the physics)
template <typename T> DSIM_DEV void load_action(const T *row, T a[4]) {
    if constexpr (std::is_same<T, float>::value) {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]) : "r"(smem_u32(row)));
    } else {
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a[0]), "=d"(a[1]) : "r"(smem_u32(row)));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a[2]), "=d"(a[3]) : "r"(smem_u32(row) + 16u));
    }
}

// lane 0: arm the slot's mbarrier and start the page loads HBM -> slot
// `parts`: bit 2 = arm the barrier with the page's total byte count, bit 0 = load the READ-ONLY rows (compiled constants +
// raw parameters: never written by a step kernel), bit 1 = load everything an earlier kernel of the stream may have
// written (state rows, the policy's actions, the setpoint rows).  7 = the whole page.
// `all_rows`: also fetch the sensordata rows (evaluate-only launches emit the STORED accelerometer; a step recomputes it)
template <typename T> DSIM_DEV void issue_page_loads(const KParams<T> &p, int page, T *slot, uint64_t *bar, int parts, bool pec, bool pref, bool all_rows,
                                                     uint64_t pol_keep, uint64_t pol_stream) {
    constexpr uint32_t rob = RO_ROWS * kTile * sizeof(T), rfb = REF_ROWS * kTile * sizeof(T);
    const uint32_t rwb = (all_rows ? RW_ROWS : RW_IN_ROWS) * kTile * (uint32_t)sizeof(T);
    const uint32_t acb = (uint32_t)min(kTile, p.n - page * kTile) * 4u * (uint32_t)sizeof(T);   // the policy's [n][4] action rows of this page
    if (parts & 4) mbar_arrive_expect_tx(bar, rwb + acb + (pec ? rob : 0u) + (pref ? rfb : 0u));
    if ((parts & 1) && pec) bulk_g2s(slot + RW_ROWS * kTile, p.ro + (size_t)page * (RO_ROWS * kTile), rob, bar, pol_keep);
    if (parts & 2) {
        bulk_g2s(slot, p.rw + (size_t)page * (RW_ROWS * kTile), rwb, bar, pol_stream);
        bulk_g2s(slot + kSlotActOff, p.actions + (size_t)page * (kTile * 4), acb, bar, pol_stream);
        if (pref) bulk_g2s(slot + (RW_ROWS + RO_ROWS) * kTile, p.refp + (size_t)page * (REF_ROWS * kTile), rfb, bar, pol_stream);
    }
}

// OBS / REW >= 0 are compile-time specialisations of the wrapper class / reward function (smaller code, no dispatch
// branches, constant observation width); -1 reads the ids from the parameter block.
// PERSISTENT + WORK-STEALING: the grid is sized to the resident-CTA capacity of the GPU (W warps).  Warp w starts on page
// w; every further page comes from a global ticket counter (page = W + ticket), grabbed one iteration ahead so that the
// page can be prefetched into the warp's second slot while the current one is being computed.  Pages are not equally
// expensive (in-kernel resets) and W rarely divides the page count, so static striding leaves a long tail.  Every active
// warp's last grab fails; the warp that draws the last ticket of the launch re-zeroes the counter (nobody grabs after it,
// and the next launch only starts grabbing once this grid has completed), so launches and CUDA-graph replays need no host reset.
// CFG >= 0 (specialised instantiations): bit 0 per-env constants, bit 1 per-env setpoints, bit 2 frame_skip == 1 - the host
// launches such an instantiation only when the handle's configuration matches; -1: everything is a run-time option
// GROUND: the generic instantiation with the floor-contact slow path compiled in (DsimConfig.ground_contact)
template <typename T, bool PEND, int OBS, int REW, int CFG = -1, bool GROUND = false>
__global__ void __launch_bounds__(kStepBlock, min_blocks<T>()) step_kernel(const __grid_constant__ KParams<T> p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar[kStepWarps][kStages];
    constexpr int DC = (OBS >= 0 && PEND) ? obs_dim_of(OBS) : 0;
    // the specialised instantiations (BASELINE configs) are only launched for plain steps: no debug timeline, not the
    // evaluate-only mode (dsim_evaluate / DSIM_TIMELINE take the generic instantiation), so those branches fold away
    constexpr bool kPlain = OBS >= 0;
    const bool pec = CFG >= 0 ? (CFG & 1) != 0 : (p.per_env_consts != 0);
    const bool pref = CFG >= 0 ? (CFG & 2) != 0 : (p.refp != nullptr);
    const int frame_skip = (CFG >= 0 && (CFG & 4)) ? 1 : p.frame_skip;
#ifdef DSIM_TL_ALL                        // instrumented experiment builds: the debug timeline stays in the specialised kernels too
    unsigned long long *const timeline = p.timeline;
#else
    unsigned long long *const timeline = kPlain ? nullptr : p.timeline;
#endif
    const bool eval_only = kPlain ? false : (p.eval_only != 0);
    // warp-uniform values (warp index, page numbers) go through redux.sync: the compiler then knows they are uniform, keeps
    // the page / slot address arithmetic in the uniform datapath and hands the bulk-copy instructions uniform registers
    // directly instead of wrapping each one in a register-broadcast loop (~15 instructions per copy, 7 copies per page)
    const int lane = threadIdx.x & 31, warp = __reduce_max_sync(0xffffffffu, (int)(threadIdx.x >> 5));
    const int wid = blockIdx.x * kStepWarps + warp, nwarps = gridDim.x * kStepWarps;
    // Programmatic dependent launch: the NEXT kernel of the stream may be scheduled onto SMs as soon as this grid's CTAs
    // retire (its launch latency and this grid's tail overlap).  Before the dependency wait a warp only sets up its barriers
    // and starts the bulk load of its first page's READ-ONLY rows (no step kernel writes them; the host clears `early_ro`
    // for the first step after a kernel that does - regen / set_params); everything an earlier kernel may have written is
    // touched after griddepcontrol.wait.
    unsigned long long t_entry = 0;
    if (timeline) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_entry));
    // (Scheduling variants that were measured and dropped - both static pages fetched up front, rotating page-less warp slot,
    // dynamically claimed second pages, L2 prefetch ahead of the wait, deferred first-page publish, late release in strict
    // mode - are listed with their numbers in DESIGN.md 4.1 item 6; the code is in the history, not here.)
    // Strict mode releases the dependent kernel at entry (its CTAs are scheduled as this grid's CTAs retire; it touches
    // nothing before its own wait).  Inputs-ready mode releases it only AFTER this kernel's own wait: a kernel that starts
    // early then knows that everything up to its predecessor's predecessor has completed, which is what makes its early
    // loads of state written R launches ago safe by construction.
    const bool trigger_late = p.early_in != 0;
    if (!trigger_late) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int my_pages = p.npages - p.page0;                       // pages of THIS launch
    const bool has_work = wid < my_pages;
    unsigned char *wslots = smem_raw + (size_t)warp * kStages * p.smem_per_slot;
    const uint64_t pol_keep = l2_policy_keep(), pol_stream = l2_policy_stream();
    // Page assignment.  The first two pages of a warp are static: page `wid`, then one of the next `nwarps` pages, dealt
    // round-robin over the CTAs so that every CTA (hence every SM) gets the same number of second pages (+-1).
    int page = p.page0 + wid;
    int next = p.page0 + nwarps + warp * (int)gridDim.x + (int)blockIdx.x;
    // Loads of the first page that may start BEFORE the dependency wait: the read-only rows (`early_ro`), and with
    // dsim_set_inputs_ready also the state / action / setpoint rows (`early_in`).
    const int pre = (p.early_ro ? 1 : 0) | (p.early_in ? 2 : 0);
    if (has_work && lane == 0) {
        #pragma unroll
        for (int k = 0; k < kStages; k++) mbar_init(&s_bar[warp][k], 1);
        issue_page_loads(p, page, reinterpret_cast<T *>(wslots), &s_bar[warp][0], 4 | pre, pec, pref, eval_only, pol_keep, pol_stream);
    }
    // With inputs_ready the wait moves to just before this warp's first store (the physics of the first page overlaps the
    // tail of the kernel ahead); otherwise nothing an earlier kernel may have written is touched before it.  The dependents
    // are released only AFTER the wait: a kernel that starts early therefore knows that everything before its predecessor
    // has completed (the early loads of the inputs-ready mode rely on exactly that).
    bool waited = p.early_in == 0;
    if (waited) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (trigger_late) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
    if (!has_work) {
        if (!waited) {
            asm volatile("griddepcontrol.wait;" ::: "memory");
            asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        }
        return;                                                    // warps are autonomous: no CTA-wide barrier below
    }
    int tl_k = 1;
    auto stamp = [&]() {
        if (timeline && lane == 0 && tl_k < 8) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            timeline[(size_t)wid * 8 + tl_k++] = t;
        }
    };
    if (timeline && lane == 0) timeline[(size_t)wid * 8] = t_entry;
    stamp();                                                       // [1] dependency wait passed (inputs-ready mode: prologue done, the wait comes later)
    const int obs_id = OBS >= 0 ? OBS : p.obs_id, reward_id = REW >= 0 ? REW : p.reward_id;
    const int D = DC > 0 ? DC : p.obs_dim;
    if (lane == 0 && pre != 3)
        issue_page_loads(p, page, reinterpret_cast<T *>(wslots), &s_bar[warp][0], 3 & ~pre, pec, pref, eval_only, pol_keep, pol_stream);
    __syncwarp();                                                  // barrier init visible to the waiting lanes
    unsigned parity = 0;                                           // bit b: phase of this warp's barrier b
    int buf = 0;
    // No atomic sits in front of the first two (static) pages: 2368 warps drawing from one counter at kernel entry cost each of them ~2 us (same-address atomics
    // serialise in L2), more than the page load they were meant to overlap.  From the third page on, work stealing: a warp
    // draws one ticket per processed page, half an iteration before it claims it (raw atom: the compiler's warp-aggregated
    // atomicAdd consumes its result at once), so the round trip to L2 hides behind the post-processing.  Launches with at
    // most two pages per warp (C4: 1.73) draw nothing.
    // Very large batches draw CHUNKS of consecutive pages per ticket: same-address atomics retire at ~3 ns each on the one
    // L2 slice that owns the counter (measured: 60 800 draws = 182 of the 189 us a 2 M-env launch took, ncu stall reason
    // long_sb on the ticket's consumer), so a launch that needs more draws than it has time for becomes atomic-bound.  The
    // chunk keeps >= 8 draws per warp for load balance.  A second counter of finished warps re-zeroes both.
    const bool stealing = my_pages > 2 * nwarps;
    const int chunk = stealing ? max(1, min(16, (my_pages - 2 * nwarps) / (8 * nwarps))) : 1;
    int chunk_next = 0, chunk_end = 0;                             // pages of the current chunk still to be handed out
    auto draw = [&]() {
        unsigned tk = 0;
        if (stealing && chunk_next >= chunk_end && lane == 0) asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(tk) : "l"(p.ticket));
        return tk;
    };
    auto claim = [&](unsigned tk) {                                // tk: lane 0's draw
        if (!stealing) return p.npages;
        if (chunk_next >= chunk_end) {
            chunk_next = __reduce_max_sync(0xffffffffu, lane == 0 ? p.page0 + 2 * nwarps + (int)tk * chunk : 0);
            chunk_end = min(chunk_next + chunk, p.npages);
        }
        return chunk_next++;
    };
    #pragma unroll 1
    while (page < p.npages) {
        const int i = page * kTile + lane;
        const bool active = i < p.n;                               // pad lanes of the last page compute, but publish nothing
        T *slot = reinterpret_cast<T *>(wslots + (size_t)buf * p.smem_per_slot);
        T *s_rw = slot, *s_ro = slot + RW_ROWS * kTile, *s_ref = slot + (RW_ROWS + RO_ROWS) * kTile, *s_obs = slot + kSlotObsOff;
        mbar_wait(&s_bar[warp][buf], (parity >> buf) & 1u);
        parity ^= 1u << buf;
        stamp();                                                   // [2], [4]: page landed

        // ---- physics
        T *col = s_rw + lane;
        const T *ro_col = s_ro + lane;
        const T *s_act = slot + kSlotActOff + 4 * lane;             // this env's raw action row (pad lanes: stale, finite or not — never published)
        EnvState<T> s = load_state(col);
        // ---- prefetch the next page into the other slot.  Its previous contents left with the bulk stores issued at the
        // end of the previous iteration; their shared-memory reads complete within a few hundred cycles.
        if (lane == 0 && next < p.npages) {
            bulk_wait_read();
            issue_page_loads(p, next, reinterpret_cast<T *>(wslots + (size_t)(buf ^ 1) * p.smem_per_slot), &s_bar[warp][buf ^ 1], 7, pec, pref, eval_only, pol_keep, pol_stream);
        }

        {
            const EnvConsts<T> c = load_consts(p, ro_col, pec);
            T a[4], ctrl[4];
            load_action(s_act, a);
            #pragma unroll
            for (int k = 0; k < 4; k++) ctrl[k] = clamp_(T(0.1) + T(0.9) * a[k], T(0), T(1));   // :269 + ctrlrange (0,1) clamp of mj_fwdActuation
            #pragma unroll 1
            if constexpr (GROUND) {
                T gprm[6];
                load_params(p, ro_col, gprm, pec);
                const GroundCtx<T> g = make_ground_ctx<T>(p.start_t[2], gprm, p.geo, p.ld, active ? i : 0, p.pendulum);
                #pragma unroll 1
                for (int f = 0; f < frame_skip; f++) {
                    if (g.start_z + s.pos.z < g.reach) {
                        // rare, out of line.  Through COPIES: a variable whose address reaches a call lives in local memory for
                        // its whole lifetime, and the in-air path must keep its state in registers
                        EnvState<T> s2 = s;
                        const EnvConsts<T> c2 = c;
                        const GroundCtx<T> g2 = g;
                        const T ctrl2[4] = {ctrl[0], ctrl[1], ctrl[2], ctrl[3]};
                        ground_substep<T, PEND>(s2, c2, ctrl2, p.h, g2);
                        s = s2;
                    } else substep<T, PEND, true>(s, c, ctrl, p.h);
                }
            } else {
                #pragma unroll 1
                for (int f = 0; f < frame_skip; f++) substep<T, PEND, true>(s, c, ctrl, p.h);
            }
        }
        const unsigned drawn = draw();                             // for the page after `next`; claimed at the end of the iteration
        // ---- counters, termination, reward, observation
        int ns = slot_to_int(col[RW_NUM_STEPS * kTile]) + (eval_only ? 0 : 1);
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
        load_ref(p, s_ref + lane, ref_off, ref_yaw, ref64, pref);
        T prm[6];
        load_params(p, ro_col, prm, pec);
        const PostState<T> ps = post_state(s, ref_off, ref_yaw);
        const bool trunc = terminated(s.pos, p.start, ref64, p.max_d2, ns, p.max_steps) || bad;
        T a[4];
        load_action(s_act, a);                                     // re-read instead of holding four registers across the physics
        const T rew = bad ? T(0) : reward_fn<T, PEND>(reward_id, s, ps, a, ns, prm, p.max_distance_t);
        __syncwarp();                                              // every lane has its read-only operands in registers:
        if constexpr (DC > 0 && DC % 2 == 0 && std::is_same<T, float>::value) {
            // the observation block may now overlay them.  Compile-time layout: components are gathered in registers and
            // leave as 8- / 16-byte shared-memory stores (row pitch DC floats: conflict-free for DC = 22, 4-way for DC = 16
            // instead of the 2- / 16-way conflicts of scalar stores)
            float o[DC];
            emit_obs<T, PEND>(obs_id, s, ps, mk(p.start_t[0], p.start_t[1], p.start_t[2]), ref_off, prm, [&](int j, float v) { o[j] = v; });
            float *row = s_obs + lane * DC;
            if constexpr (DC % 4 == 0) {
                #pragma unroll
                for (int k = 0; k < DC / 4; k++) reinterpret_cast<float4 *>(row)[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
            } else {
                #pragma unroll
                for (int k = 0; k < DC / 2; k++) reinterpret_cast<float2 *>(row)[k] = make_float2(o[2 * k], o[2 * k + 1]);
            }
        } else {
            ObsWriter<T, 1> w; w.base = s_obs + lane * D; w.stride = 1;
            emit_obs<T, PEND>(obs_id, s, ps, mk(p.start_t[0], p.start_t[1], p.start_t[2]), ref_off, prm, w);
        }
        if (!eval_only) {
            // would-be ground contact (the floor plane is out of reach in the BASELINE configs; detected, never ignored)
            double *const st = p.stats + (wid & (kStatReplicas - 1)) * 8;
            if (active && p.start_t[2] + s.pos.z < prm[4] + T(0.5)) atomicAdd(st + 4, 1.0);
            T ret = col[RW_EP_RETURN * kTile] + rew;
            if (trunc && active) {
                atomicAdd(st + 0, (double)ret); atomicAdd(st + 1, (double)ns); atomicAdd(st + 2, 1.0);
                if (bad) atomicAdd(st + 3, 1.0);
            }
            col[RW_EP_RETURN * kTile] = trunc ? T(0) : ret;
            col[RW_NUM_STEPS * kTile] = int_to_slot<T>(ns);
            store_state(col, s);
        }
        {
            const unsigned need = __ballot_sync(0xffffffffu, !eval_only && active && trunc && (p.auto_reset || bad));
            if (need) resample_page<T, PEND>(need, s_rw, p.rc, p.seed, p.env_base + (unsigned)(page * kTile));
        }
        fence_async_smem();                                        // generic-proxy writes of the slot -> visible to the bulk stores
        __syncwarp();
        // ---- publish: slot -> HBM (plus the two per-env scalars that go straight to global memory)
        auto publish = [&](int pg, T *pslot, T prew, bool ptrunc) {
            const int pi = pg * kTile + lane;
            if (pi < p.n) {
                p.reward[pi] = prew;
                p.trunc[pi] = ptrunc ? 1 : 0;
                if (p.reward_host) p.reward_host[pi] = prew;
                if (p.trunc_host) p.trunc_host[pi] = ptrunc ? 1 : 0;
            }
            const int nvalid = min(kTile, p.n - pg * kTile);
            const uint32_t obs_bytes = (uint32_t)(nvalid * D) * (uint32_t)sizeof(T);
            T *gobs = p.obs + (size_t)pg * kTile * D, *pobs = pslot + kSlotObsOff;
            const bool obs_bulk = (obs_bytes & 15u) == 0;          // always true for full pages
            if (lane == 0) {
                if (!eval_only) bulk_s2g(p.rw + (size_t)pg * (RW_ROWS * kTile), pslot, RW_ROWS * kTile * sizeof(T), pol_stream);
                if (obs_bulk) bulk_s2g(gobs, pobs, obs_bytes, pol_stream);
                if (obs_bulk && p.obs_host) bulk_s2g(p.obs_host + (size_t)pg * kTile * D, pobs, obs_bytes);
                bulk_commit();
            }
            if (!obs_bulk)                                         // ragged last page whose byte count is not a multiple of 16
                #pragma unroll 1
                for (int e = lane; e < nvalid * D; e += kTile) {
                    gobs[e] = pobs[e];
                    if (p.obs_host) p.obs_host[(size_t)pg * kTile * D + e] = pobs[e];
                }
        };
        if (!waited) {
            // inputs-ready mode: the dependency wait is the very LAST thing before this warp's first global store (warp-uniform),
            // after the statistics, the write-back of the state into the slot and the reset path: 9.3 us per C4 step against
            // 10.1 us with the wait in front of that tail work
            waited = true;
            asm volatile("griddepcontrol.wait;" ::: "memory");
            asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        }
        publish(page, slot, rew, trunc);
        stamp();                                                   // [3], [5]: page published
        page = next;
        next = claim(drawn);
        buf ^= 1;
    }
    if (lane == 0) bulk_wait_read();                               // the slots must outlive the bulk reads
    if (stealing && lane == 0) {                                   // the last warp out re-zeroes the counters: launches and graph replays need no host reset
        unsigned d;
        asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(d) : "l"(p.ticket + 1));
        if (d == (unsigned)(nwarps - 1)) { p.ticket[0] = 0u; p.ticket[1] = 0u; }
    }
    if (timeline && lane == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        timeline[(size_t)wid * 8 + 7] = t;                        // [7] exit
    }
}
