// dsim_step_x2.cuh — the fused env-step kernel with TWO envs per lane (included by dsim_kernels.cu after dsim_step.cuh).
//
// Same step, same page layout, same bulk-copy / mbarrier / programmatic-launch protocol as step_kernel; what changes is the unit
// of work.  A warp owns a PAIR of consecutive pages (64 envs) at a time: lane l holds env l of page A in the low halves and env l
// of page B in the high halves of 64-bit registers, and the physics (substep<F2>, dsim_packed.cuh) runs on the packed FP32
// forms - one FFMA2 / FMUL2 / FADD2 where the one-env kernel issues two scalar instructions.  Scalar FP32 issues every other
// cycle per scheduler on sm_100, so a wave of pages in the one-env kernel is FMA-pipe-bound (4 warps x 886 FP32 instructions
// x 2 cycles = 3.7 us per page wave); the packed physics halves the 707 of those 886 that belong to the substep.
// Termination, reward, observation and the reset path stay scalar and run once per page of the pair.
// Registers double (two states live), so two CTAs of four warps per SM instead of four, each warp with FOUR page slots (the pair
// being computed + the pair in flight): the same shared memory, the same number of envs resident per SM.
// Pairs are dealt statically (pair = warp + k * warps), for plain FP32 steps of the specialised BASELINE configurations while a
// warp has at most two pairs (<= ~150 000 envs).
//
// MEASURED, AND NOT THE DEFAULT (DSIM_X2=1 selects it; tools/gpu_x2.sh): correct - the whole parity suite passes through it -
// and the FMA-pipe work per env does drop (substep: 707 scalar FP32 instructions -> 342 packed per env + 74 LOP3 for the sign
// flips of fused subtractions), but C4 takes 10.9 us per step against 9.2 us (hot L2: 12.6 against 9.4).  With 252 registers
// per thread only 8 warps are resident per SM instead of 16, each with the same dependency depth per page and now both pages'
// scalar tails in series: per-warp IPC is ~0.26, two warps per scheduler cannot cover the latency, and the kernel turns from
// FMA-pipe-bound into latency-bound.  Three CTAs per SM (168 registers, 96 B of spills, two slots per warp, no prefetch):
// 12.7 us.  The packed FP32 pipe pays where the work per thread has spare ILP at constant occupancy (the FP32 policy kernel:
// 2.77 -> 2.31 ms), not where it has to be bought with occupancy.
#pragma once

#ifndef DSIM_X2_SLOTS
#define DSIM_X2_SLOTS 4
#endif
#ifndef DSIM_X2_MINB
#define DSIM_X2_MINB 2
#endif
constexpr int kX2Slots = DSIM_X2_SLOTS;                                       // page slots per warp (4: the pair in flight is prefetched; 2: no prefetch)

template <int OBS, int REW, int CFG>
__global__ void __launch_bounds__(kStepBlock, DSIM_X2_MINB) step_kernel_x2(const __grid_constant__ KParams<float> p) {
    using T = float;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar[kStepWarps][kX2Slots];
    constexpr int DC = obs_dim_of(OBS);
    static_assert(OBS >= 0 && REW >= 0 && CFG >= 0 && DC > 0 && DC % 2 == 0, "specialised instantiations only");
    constexpr bool pec = (CFG & 1) != 0, pref = (CFG & 2) != 0;
    const int frame_skip = (CFG & 4) ? 1 : p.frame_skip;
    const int lane = threadIdx.x & 31, warp = __reduce_max_sync(0xffffffffu, (int)(threadIdx.x >> 5));
    const int wid = blockIdx.x * kStepWarps + warp, nwarps = gridDim.x * kStepWarps;
    const bool trigger_late = p.early_in != 0;
    if (!trigger_late) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int npairs = (p.npages - p.page0 + 1) / 2;
    const bool has_work = wid < npairs;
    unsigned char *wslots = smem_raw + (size_t)warp * kX2Slots * p.smem_per_slot;
    auto slot_ptr = [&](int k) { return reinterpret_cast<T *>(wslots + (size_t)k * p.smem_per_slot); };
    const uint64_t pol_keep = l2_policy_keep(), pol_stream = l2_policy_stream();
    int pair = wid, next = wid + nwarps;
    // pages of a pair; the second one may not exist (odd page count): its half of the lanes computes on the first page's data
    auto pages_of = [&](int pr, int &a, int &b) { a = p.page0 + 2 * pr; b = a + 1 < p.npages ? a + 1 : -1; };
    auto issue_pair = [&](int pr, int buf, int parts) {
        int a, b;
        pages_of(pr, a, b);
        issue_page_loads(p, a, slot_ptr(2 * buf), &s_bar[warp][2 * buf], parts, pec, pref, false, pol_keep, pol_stream);
        if (b >= 0) issue_page_loads(p, b, slot_ptr(2 * buf + 1), &s_bar[warp][2 * buf + 1], parts, pec, pref, false, pol_keep, pol_stream);
    };
    const int pre = (p.early_ro ? 1 : 0) | (p.early_in ? 2 : 0);
    if (has_work && lane == 0) {
        #pragma unroll
        for (int k = 0; k < kX2Slots; k++) mbar_init(&s_bar[warp][k], 1);
        issue_pair(pair, 0, 4 | pre);
    }
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
        return;
    }
    if (lane == 0 && pre != 3) issue_pair(pair, 0, 3 & ~pre);
    __syncwarp();
    unsigned parity = 0;                                                       // bit k: phase of barrier k
    int buf = 0;
    #pragma unroll 1
    while (pair < npairs) {
        int pgA, pgB;
        pages_of(pair, pgA, pgB);
        const bool two = pgB >= 0;
        T *slotA = slot_ptr(2 * buf), *slotB = two ? slot_ptr(2 * buf + 1) : slotA;
        mbar_wait(&s_bar[warp][2 * buf], (parity >> (2 * buf)) & 1u);
        parity ^= 1u << (2 * buf);
        if (two) { mbar_wait(&s_bar[warp][2 * buf + 1], (parity >> (2 * buf + 1)) & 1u); parity ^= 1u << (2 * buf + 1); }

        // ---- physics of both pages, packed
        EnvState<T> sA, sB;
        {
            const T *colA = slotA + lane, *colB = slotB + lane;
            const EnvState<T> a = load_state(colA), b = load_state(colB);
            if (kX2Slots == 4 && lane == 0 && next < npairs) {                 // prefetch the next pair into the other two slots
                bulk_wait_read();
                issue_pair(next, buf ^ 1, 7);
            }
            EnvState<F2> s;
            s.pos = mk(F2(a.pos.x, b.pos.x), F2(a.pos.y, b.pos.y), F2(a.pos.z, b.pos.z));
            s.qw = F2(a.qw, b.qw); s.qx = F2(a.qx, b.qx); s.qy = F2(a.qy, b.qy); s.qz = F2(a.qz, b.qz);
            s.hx = F2(a.hx, b.hx); s.hy = F2(a.hy, b.hy);
            s.vel = mk(F2(a.vel.x, b.vel.x), F2(a.vel.y, b.vel.y), F2(a.vel.z, b.vel.z));
            s.om = mk(F2(a.om.x, b.om.x), F2(a.om.y, b.om.y), F2(a.om.z, b.om.z));
            s.hvx = F2(a.hvx, b.hvx); s.hvy = F2(a.hvy, b.hvy);
            #pragma unroll
            for (int k = 0; k < 4; k++) s.act[k] = F2(a.act[k], b.act[k]);
            s.acc = mk(F2(0.f), F2(0.f), F2(0.f));
            const T *roA = slotA + RW_ROWS * kTile + lane, *roB = slotB + RW_ROWS * kTile + lane;
            F2 cv[C_ROWS];
            #pragma unroll
            for (int k = 0; k < C_ROWS; k++) cv[k] = pec ? F2(roA[(RO_CONSTS + k) * kTile], roB[(RO_CONSTS + k) * kTile]) : F2(p.uconst[k]);
            const EnvConsts<F2> c = consts_from(cv);
            T aA[4], aB[4];
            load_action(slotA + kSlotActOff + 4 * lane, aA);
            load_action(slotB + kSlotActOff + 4 * lane, aB);
            F2 ctrl[4];
            #pragma unroll
            for (int k = 0; k < 4; k++) ctrl[k] = F2(clamp_(0.1f + 0.9f * aA[k], 0.f, 1.f), clamp_(0.1f + 0.9f * aB[k], 0.f, 1.f));   // :269 + ctrlrange clamp
            const F2 h2(p.h);
            #pragma unroll 1
            for (int f = 0; f < frame_skip; f++) substep<F2, true, true>(s, c, ctrl, h2);
            auto half = [&](EnvState<T> &o, bool hi) {
                auto g = [&](F2 v) { return hi ? v.hi() : v.lo(); };
                o.pos = mk(g(s.pos.x), g(s.pos.y), g(s.pos.z));
                o.qw = g(s.qw); o.qx = g(s.qx); o.qy = g(s.qy); o.qz = g(s.qz); o.hx = g(s.hx); o.hy = g(s.hy);
                o.vel = mk(g(s.vel.x), g(s.vel.y), g(s.vel.z)); o.om = mk(g(s.om.x), g(s.om.y), g(s.om.z));
                o.hvx = g(s.hvx); o.hvy = g(s.hvy);
                #pragma unroll
                for (int k = 0; k < 4; k++) o.act[k] = g(s.act[k]);
                o.acc = mk(g(s.acc.x), g(s.acc.y), g(s.acc.z));
            };
            half(sA, false); half(sB, true);
        }

        // ---- per page: counters, termination, reward, observation, statistics, state write-back, resets (scalar, as in step_kernel)
        T rewA = T(0), rewB = T(0);
        bool truncA = false, truncB = false;
        auto finish = [&](int pg, T *slot, EnvState<T> &s, T &rew_out, bool &trunc_out) {
            const int i = pg * kTile + lane;
            const bool active = i < p.n;
            T *s_rw = slot, *s_ro = slot + RW_ROWS * kTile, *s_ref = slot + (RW_ROWS + RO_ROWS) * kTile, *s_obs = slot + kSlotObsOff;
            T *col = s_rw + lane;
            const T *ro_col = s_ro + lane;
            const int ns = slot_to_int(col[RW_NUM_STEPS * kTile]) + 1;
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
            load_action(slot + kSlotActOff + 4 * lane, a);
            const T rew = bad ? T(0) : reward_fn<T, true>(REW, s, ps, a, ns, prm, p.max_distance_t);
            __syncwarp();                                                      // every lane holds its read-only operands: the observation block may overlay them
            {
                float o[DC];
                emit_obs<T, true>(OBS, s, ps, mk(p.start_t[0], p.start_t[1], p.start_t[2]), ref_off, prm, [&](int j, float v) { o[j] = v; });
                float *row = s_obs + lane * DC;
                if constexpr (DC % 4 == 0) {
                    #pragma unroll
                    for (int k = 0; k < DC / 4; k++) reinterpret_cast<float4 *>(row)[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
                } else {
                    #pragma unroll
                    for (int k = 0; k < DC / 2; k++) reinterpret_cast<float2 *>(row)[k] = make_float2(o[2 * k], o[2 * k + 1]);
                }
            }
            double *const st = p.stats + (wid & (kStatReplicas - 1)) * 8;
            if (active && p.start_t[2] + s.pos.z < prm[4] + T(0.5)) atomicAdd(st + 4, 1.0);
            const T ret = col[RW_EP_RETURN * kTile] + rew;
            if (trunc && active) {
                atomicAdd(st + 0, (double)ret); atomicAdd(st + 1, (double)ns); atomicAdd(st + 2, 1.0);
                if (bad) atomicAdd(st + 3, 1.0);
            }
            col[RW_EP_RETURN * kTile] = trunc ? T(0) : ret;
            col[RW_NUM_STEPS * kTile] = int_to_slot<T>(ns);
            store_state(col, s);
            const unsigned need = __ballot_sync(0xffffffffu, active && trunc && (p.auto_reset || bad));
            if (need) resample_page<T, true>(need, s_rw, p.rc, p.seed, p.env_base + (unsigned)(pg * kTile));
            rew_out = rew; trunc_out = trunc;
        };
        finish(pgA, slotA, sA, rewA, truncA);
        if (two) finish(pgB, slotB, sB, rewB, truncB);
        fence_async_smem();
        __syncwarp();
        // ---- publish: slots -> HBM
        auto publish = [&](int pg, T *pslot, T prew, bool ptrunc) {
            const int pi = pg * kTile + lane;
            if (pi < p.n) {
                p.reward[pi] = prew;
                p.trunc[pi] = ptrunc ? 1 : 0;
                if (p.reward_host) p.reward_host[pi] = prew;
                if (p.trunc_host) p.trunc_host[pi] = ptrunc ? 1 : 0;
            }
            const int nvalid = min(kTile, p.n - pg * kTile);
            const uint32_t obs_bytes = (uint32_t)(nvalid * DC) * (uint32_t)sizeof(T);
            T *gobs = p.obs + (size_t)pg * kTile * DC, *pobs = pslot + kSlotObsOff;
            const bool obs_bulk = (obs_bytes & 15u) == 0;
            if (lane == 0) {
                bulk_s2g(p.rw + (size_t)pg * (RW_ROWS * kTile), pslot, RW_ROWS * kTile * sizeof(T), pol_stream);
                if (obs_bulk) bulk_s2g(gobs, pobs, obs_bytes, pol_stream);
                if (obs_bulk && p.obs_host) bulk_s2g(p.obs_host + (size_t)pg * kTile * DC, pobs, obs_bytes);
            }
            if (!obs_bulk)
                #pragma unroll 1
                for (int e = lane; e < nvalid * DC; e += kTile) {
                    gobs[e] = pobs[e];
                    if (p.obs_host) p.obs_host[(size_t)pg * kTile * DC + e] = pobs[e];
                }
        };
        if (!waited) {
            waited = true;
            asm volatile("griddepcontrol.wait;" ::: "memory");
            asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        }
        publish(pgA, slotA, rewA, truncA);
        if (two) publish(pgB, slotB, rewB, truncB);
        if (lane == 0) bulk_commit();
        pair = next;
        next += nwarps;
        if (kX2Slots == 4) buf ^= 1;
        else if (lane == 0 && pair < npairs) {                                 // two slots: the next pair is fetched once this one has left them
            bulk_wait_read();
            issue_pair(pair, 0, 7);
        }
    }
    if (lane == 0) bulk_wait_read();
}
