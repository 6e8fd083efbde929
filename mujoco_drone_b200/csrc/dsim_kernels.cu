// dsim_kernels.cu — kernels and the C ABI of libdronesim_b200.so (see include/dronesim_b200.h).
// Target: sm_100a (B200).  One thread per env; SoA state rows are read and written fully coalesced; the
// physics, state extraction, termination, reward, observation, episode statistics and (optionally) the
// Philox reset of truncated envs all happen in ONE kernel launch per env-step.  There is no CPU fallback: every
// entry point either launches on the GPU or returns an error code.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/dronesim_b200.h"
#include "dsim_device.cuh"
#include "dsim_obs_reward.cuh"
#include "dsim_params.cuh"
#include "dsim_contact.cuh"
#include "dsim_packed.cuh"
#include "dsim_policy.cuh"

using namespace dsim;

namespace {

constexpr int kBlock = 128;

#include "dsim_step.cuh"
#include "dsim_step_x2.cuh"

// mj_forward after set_state (mujoco_vecenv.py:396-402): refresh sensordata (+ obs) from the current state
template <typename T, bool PEND>
__global__ void __launch_bounds__(kBlock) forward_kernel(const KParams<T> p, int refresh_obs) {
    const int i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= p.n) return;
    T *col = p.rw + page_elem(RW_ROWS, 0, i);
    const T *ro_col = p.ro + page_elem(RO_ROWS, 0, i);
    EnvState<T> s = load_state(col);
    const EnvConsts<T> c = load_consts(p, ro_col, p.per_env_consts != 0);
    const T ctrl[4] = {T(0), T(0), T(0), T(0)};
    if (p.ground) {
        T gprm[6];
        load_params(p, ro_col, gprm, p.per_env_consts != 0);
        const GroundCtx<T> g = make_ground_ctx<T>(p.start_t[2], gprm, p.geo, p.ld, i, p.pendulum);
        substep<T, PEND, false, true>(s, c, ctrl, p.h, &g);
    } else substep<T, PEND, false>(s, c, ctrl, p.h);
    col[S_ACC * kTile] = s.acc.x; col[(S_ACC + 1) * kTile] = s.acc.y; col[(S_ACC + 2) * kTile] = s.acc.z;
    if (refresh_obs) {
        V3<T> ref_off; T ref_yaw; double ref64[3];
        load_ref(p, p.refp ? p.refp + page_elem(REF_ROWS, 0, i) : nullptr, ref_off, ref_yaw, ref64, p.refp != nullptr);
        T prm[6];
        load_params(p, ro_col, prm, p.per_env_consts != 0);
        const PostState<T> ps = post_state(s, ref_off, ref_yaw);
        ObsWriter<T, 1> w; w.base = p.obs + (size_t)i * p.obs_dim; w.stride = 1;
        emit_obs<T, PEND>(p.obs_id, s, ps, mk(p.start_t[0], p.start_t[1], p.start_t[2]), ref_off, prm, w);
    }
}

// reset_model / reset_at: sample_state into qpos/qvel, num_steps = 0; act, ctrl persist (Q3)
template <typename T, bool PEND>
__global__ void __launch_bounds__(kBlock) reset_kernel(const KParams<T> p, const unsigned char *mask, int single, int first) {
    int i = blockIdx.x * kBlock + threadIdx.x;
    if (single >= 0) { if (i != 0) return; i = single; }
    if (i >= p.n) return;
    if (mask && !mask[i]) return;
    T *col = p.rw + page_elem(RW_ROWS, 0, i);
    EnvState<T> s = load_state(col);
    const unsigned rcnt = first ? 0u : (unsigned)slot_to_int(col[RW_RESET_COUNT * kTile]) + 1u;
    sample_state<T, PEND>(s, p.rc, p.seed, p.env_base + (unsigned)i, rcnt);
    col[RW_RESET_COUNT * kTile] = int_to_slot<T>((int)rcnt);
    col[RW_NUM_STEPS * kTile] = int_to_slot<T>(0);
    col[RW_EP_RETURN * kTile] = T(0);
    store_state(col, s);
}

// get_drone_states (:357-380) for every env -> [n][33|29]
template <typename T, bool PEND>
__global__ void __launch_bounds__(kBlock) states_kernel(const KParams<T> p, T *out, int width) {
    const int i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= p.n) return;
    const EnvState<T> s = load_state(p.rw + page_elem(RW_ROWS, 0, i));
    V3<T> ref_off; T ref_yaw; double ref64[3];
    load_ref(p, p.refp ? p.refp + page_elem(REF_ROWS, 0, i) : nullptr, ref_off, ref_yaw, ref64, p.refp != nullptr);
    T prm[6];
    load_params(p, p.ro + page_elem(RO_ROWS, 0, i), prm, p.per_env_consts != 0);
    const PostState<T> ps = post_state(s, ref_off, ref_yaw);
    ObsWriter<T> w; w.base = out + (size_t)i * width; w.stride = 1;
    emit_state_row<T, PEND>(s, ps, mk(p.start_t[0], p.start_t[1], p.start_t[2]), ref_off, prm, w);
}

// generate_drone_params (:180-216) on device: Philox stream 1 keyed by (seed, global env, epoch)
__global__ void draw_params_kernel(int n, int ld, double *params64, unsigned seed, unsigned env_base, unsigned epoch,
                                   int random_params, int pendulum, double difficulty, const double *center_hw /*12*/) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v[6];
    for (int k = 0; k < 6; k++) v[k] = center_hw[k];
    if (random_params) {
        const U4 x0 = philox4x32(0, epoch, 1, 0, seed, env_base + i), x1 = philox4x32(1, epoch, 1, 0, seed, env_base + i);
        const uint32_t u[6] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y};
        for (int k = 0; k < 6; k++) { const double w = center_hw[6 + k]; v[k] = center_hw[k] + (-w + 2 * w * u01<double>(u[k])) * difficulty; }
    }
    if (!pendulum) { v[4] = 0; v[5] = 0; }                      // self.pendulum * pendulum_lens[i] (:212-213)
    for (int k = 0; k < 6; k++) params64[(size_t)k * ld + i] = v[k];
}
// env_gen + MuJoCo compile on device: FP64 drone_params [6][ld] -> read-only page (constants + raw params)
template <typename T>
__global__ void compile_kernel(int n, int ld, const double *params64, T *ro, int pendulum, int rounding) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p[6], c[C_ROWS];
    for (int k = 0; k < 6; k++) p[k] = params64[(size_t)k * ld + i];
    compile_consts(p, pendulum != 0, rounding != 0, c);
    T *col = ro + page_elem(RO_ROWS, 0, i);
    for (int k = 0; k < C_ROWS; k++) col[(RO_CONSTS + k) * kTile] = (T)c[k];
    for (int k = 0; k < 6; k++) col[(RO_PARAMS + k) * kTile] = (T)p[k];
}
// collision geometry of every drone for the floor-contact path (DsimConfig.ground_contact), same FP64 + "%.5g" stage as the compile
template <typename T>
__global__ void geometry_kernel(int n, int ld, const double *params64, T *geo, int pendulum, int rounding) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p[6], g[GEO_ROWS];
    for (int k = 0; k < 6; k++) p[k] = params64[(size_t)k * ld + i];
    contact_geometry(p, pendulum != 0, rounding != 0, g);
    for (int k = 0; k < GEO_ROWS; k++) geo[(size_t)k * ld + i] = (T)g[k];
}
// control_reference (:151-172) per env: axes [4][ld] are the already sign-flipped joystick values (x, -y, -z, -yaw).
// The dead-zone DECISIONS (:160-161) are index logic and are taken in FP64 whatever the page precision: an axis value that is
// a 2-decimal number which went through FP32 (joystick.py:36 rounds to 2 decimals; 0.12f != 0.12) is restored to the double
// the reference sees, products and sums are not contracted, so the active / inactive bits equal the reference's; the
// setpoint itself is accumulated in FP64 and stored in the page's precision.
template <typename T>
__global__ void control_reference_kernel(int n, int ld, const T *axes, T *refp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    T *col = refp + page_elem(REF_ROWS, 0, i);
    double pert[4], r[4];
    for (int k = 0; k < 4; k++) {
        const double a = (double)axes[(size_t)k * ld + i], c = rint(a * 100.0);
        pert[k] = fabs(a * 100.0 - c) < 1e-3 ? c / 100.0 : a;
        r[k] = (double)col[k * kTile];
    }
    const bool xy = sqrt(__dadd_rn(__dmul_rn(pert[0], pert[0]), __dmul_rn(pert[1], pert[1]))) > 0.2;
    const bool zy = sqrt(__dadd_rn(__dmul_rn(pert[2], pert[2]), __dmul_rn(pert[3], pert[3]))) > 0.2;
    const double lim[3] = {5.0, 5.0, 6.0};
    for (int k = 0; k < 4; k++) {
        const double mag = fmax(fabs(pert[k]) - 0.1, 0.0);
        const double sg = pert[k] > 0.0 ? 1.0 : (pert[k] < 0.0 ? -1.0 : 0.0);
        const bool active = k < 2 ? xy : zy;
        r[k] = __dadd_rn(r[k], __dmul_rn(__dmul_rn(__dmul_rn(0.1, mag), sg), active ? 1.0 : 0.0));
    }
    double y = fmod(__dadd_rn(r[3], kPi), 2 * kPi);                         // Python's float %: sign of the divisor
    if (y < 0.0) y += 2 * kPi;
    r[3] = __dsub_rn(y, kPi);                                               // (yaw + pi) % (2 pi) - pi  (:168)
    for (int k = 0; k < 3; k++) r[k] = fmin(fmax(r[k], -lim[k]), lim[k]);   // clip to start_pos +- (5,5,6); ref rows are offsets
    for (int k = 0; k < 4; k++) col[k * kTile] = (T)r[k];
}
// setpoint streams on the device (evaluation.py:135-152 gen_circle / gen_step / gen_ramp_trajectory): env i follows the
// trajectory at time t + i * phase_step, so a batch covers every phase of it; the setpoint page holds offsets from start_pos
template <typename T>
__global__ void trajectory_kernel(int n, T *refp, int kind, double t, double phase_step, double p0, double p1, double p2,
                                  double s0, double s1, double s2, double s3, double e0, double e1, double e2, double e3,
                                  double st0, double st1, double st2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double ti = t + phase_step * i;
    double r[4];
    if (kind == DSIM_TRAJ_CIRCLE) {               // p0 = f, p1 = r, p2 = h  (:135-138)
        r[0] = p1 * cos(2 * kPi * p0 * ti); r[1] = p1 * sin(2 * kPi * p0 * ti); r[2] = p2; r[3] = 0;
    } else if (kind == DSIM_TRAJ_STEP) {          // p0 = step_time  (:141-144)
        const bool a = ti < p0;
        r[0] = a ? s0 : e0; r[1] = a ? s1 : e1; r[2] = a ? s2 : e2; r[3] = a ? s3 : e3;
    } else {                                      // ramp: p0 = start_time, p1 = duration  (:147-152)
        const double w = ti < p0 ? 0.0 : (ti - p0) / (p1 - p0);
        r[0] = s0 + w * (e0 - s0); r[1] = s1 + w * (e1 - s1); r[2] = s2 + w * (e2 - s2); r[3] = s3 + w * (e3 - s3);
    }
    T *col = refp + page_elem(REF_ROWS, 0, i);
    col[0] = (T)(r[0] - st0); col[kTile] = (T)(r[1] - st1); col[2 * kTile] = (T)(r[2] - st2); col[3 * kTile] = (T)r[3];
}
// qpos0 of the freshly built model: every drone sits on env_gen.make_sim's spawn grid (env_gen.py:114-122: sz = ceil(sqrt(N)),
// pitch 0.5 m, centred, z = 0.15 m; drone i at column i % sz, row i / sz of numpy's default 'xy' meshgrid), quaternion
// identity, hinges 0 - what MjData holds until the first vector_reset
template <typename T> __global__ void spawn_grid_kernel(int n, T *rw, double sx, double sy, double sz_) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int side = (int)ceil(sqrt((double)n));
    const double x = ((i % side) - 0.5 * (side - 1)) * 0.5, y = ((i / side) - 0.5 * (side - 1)) * 0.5;
    T *col = rw + page_elem(RW_ROWS, 0, i);
    col[0 * kTile] = (T)(x - sx); col[1 * kTile] = (T)(y - sy); col[2 * kTile] = (T)(0.15 - sz_);
    col[S_QUAT * kTile] = T(1);
}
// one row of a paged buffer := value, for the first n envs
template <typename T> __global__ void fill_row_kernel(int n, T *base, int page_rows, int row, T value) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    base[page_elem(page_rows, row, i)] = value;
}

}  // namespace

// ====================================================================== host side
constexpr int kHostChunks = 8;
struct DsimHandle {
    DsimConfig cfg;
    int device, n, ld, npages, obs_dim, state_width;
    size_t rs;                        // sizeof(real)
    void *rw, *ro, *refp, *obs, *reward, *states33, *actions_stage;
    double *params64, *stats, *center_hw;
    void *geo;                       // floor contact: [GEO_ROWS][ld] per-env collision geometry, or nullptr
    unsigned long long *timeline;
    unsigned *ticket;
    unsigned char *trunc;
    int per_env_consts;
    double uconst[C_ROWS], uparams[6];
    double h;
    int first_reset_done;
    int ro_dirty;                      // a kernel that rewrote the read-only pages was queued since the last step
    int inputs_ready;                  // dsim_set_inputs_ready
    int guard;                         // DSIM_GUARD=1 at create: every device buffer sits between two 4 KB canary regions (dsim_debug_guard_check)
    void *gbase[16]; size_t gsize[16]; int ng;
    cudaStream_t hs[3];                // host entry point: copy-in, compute, copy-out streams (created on first use)
    cudaEvent_t ev_in[kHostChunks], ev_k[kHostChunks], ev_a, ev_b, ev_c;
    int host_pipeline_ready;
    int64_t launches;
    char err[512];
};

static char g_create_err[512] = "";
constexpr size_t kGuardBytes = 4096;

static int fail(DsimHandle *h, int code, const char *fmt, const char *detail) {
    char *dst = h ? h->err : g_create_err;
    snprintf(dst, 512, fmt, detail ? detail : "");
    return code;
}
#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) return fail(h, DSIM_ECUDA, "CUDA error: %s", cudaGetErrorString(e_)); \
    } while (0)

static const int kObsDimPend[DSIM_NUM_OBS] = {33, 16, 16, 23, 24, 19, 22, 25, 22, 22, 16, 15, 21, 28, 17};

extern "C" int dsim_abi_version(void) { return DSIM_ABI_VERSION; }

extern "C" int dsim_obs_dim(int obs_id, int pendulum) {
    if (obs_id < 0 || obs_id >= DSIM_NUM_OBS) return DSIM_EINVAL;
    if (obs_id == 0) return pendulum ? 33 : 29;
    return kObsDimPend[obs_id];
}

extern "C" const char *dsim_last_error(const DsimHandle *h) { return h ? h->err : g_create_err; }

// largest double t with sqrt_rn(t) <= d: turns `sqrt(d2) > max_distance` into the bit-identical `d2 > t` (see terminated())
static double max_distance_sq_threshold(double d) {
    if (!(d >= 0)) return -1.0;                       // negative / NaN radius: everything is "too far" (sqrt(d2) > d always holds for d < 0)
    if (isinf(d)) return d;
    double t = d * d;
    while (sqrt(t) > d) t = nextafter(t, 0.0);
    while (sqrt(nextafter(t, INFINITY)) <= d) t = nextafter(t, INFINITY);
    return t;
}

template <typename T> static KParams<T> make_params(const DsimHandle *h, const void *actions) {
    KParams<T> p;
    memset(&p, 0, sizeof p);
    const DsimConfig &c = h->cfg;
    p.n = h->n; p.npages = h->npages;
    p.rw = (T *)h->rw; p.ro = (const T *)h->ro;
    p.refp = c.per_env_reference ? (T *)h->refp : nullptr;
    p.obs = (T *)h->obs; p.reward = (T *)h->reward;
    p.trunc = h->trunc; p.stats = h->stats; p.actions = (const T *)actions;
    for (int k = 0; k < C_ROWS; k++) p.uconst[k] = (T)h->uconst[k];
    for (int k = 0; k < 6; k++) p.uparams[k] = (T)h->uparams[k];
    p.per_env_consts = h->per_env_consts; p.auto_reset = c.auto_reset; p.obs_id = c.obs_id; p.reward_id = c.reward_id;
    p.obs_dim = h->obs_dim; p.frame_skip = c.frame_skip;
    p.max_steps = (int)(c.max_steps > 2147483647LL ? 2147483647LL : c.max_steps);
    p.smem_per_slot = slot_bytes(h->obs_dim, (int)sizeof(T));
    p.h = (T)h->h; p.max_distance_t = (T)c.max_distance; p.max_d2 = max_distance_sq_threshold(c.max_distance);
    for (int k = 0; k < 3; k++) {
        p.ref_off[k] = (T)(c.reference[k] - c.start_pos[k]);
        p.start_t[k] = (T)c.start_pos[k]; p.start[k] = c.start_pos[k]; p.ref64[k] = c.reference[k];
    }
    p.ref_yaw = (T)c.reference[3];
    p.rc.start_yaw = (T)c.start_pos[3]; p.rc.max_pos_offset = (T)c.max_pos_offset;
    for (int k = 0; k < 2; k++) { p.rc.angle_sigma[k] = (T)c.angle_sigma[k]; p.rc.pend_rp_sigma[k] = (T)c.pend_rp_sigma[k]; p.rc.pend_vel_sigma[k] = (T)c.pend_vel_sigma[k]; }
    for (int k = 0; k < 3; k++) { p.rc.vel_sigma[k] = (T)c.vel_sigma[k]; p.rc.ang_vel_sigma[k] = (T)c.ang_vel_sigma[k]; }
    p.rc.random_start_pos = c.random_start_pos;
    p.seed = c.seed; p.env_base = (unsigned)c.env_id_offset;
    p.timeline = h->timeline; p.ticket = h->ticket;
    p.early_ro = h->ro_dirty ? 0 : 1;
    p.early_in = h->inputs_ready ? 1 : 0;
    p.geo = (const T *)h->geo; p.ld = h->ld; p.pendulum = c.pendulum; p.ground = c.ground_contact;
    return p;
}

#define DISPATCH(h, KERNEL, stream, ...)                                                                  \
    do {                                                                                                  \
        const int grid_ = ((h)->n + kBlock - 1) / kBlock;                                                 \
        cudaStream_t st_ = (cudaStream_t)(stream);                                                        \
        if ((h)->cfg.precision == DSIM_FP32) {                                                            \
            auto kp = make_params<float>((h), actions_);                                                  \
            if ((h)->cfg.pendulum) KERNEL<float, true><<<grid_, kBlock, 0, st_>>>(kp, ##__VA_ARGS__);     \
            else KERNEL<float, false><<<grid_, kBlock, 0, st_>>>(kp, ##__VA_ARGS__);                      \
        } else {                                                                                          \
            auto kp = make_params<double>((h), actions_);                                                 \
            if ((h)->cfg.pendulum) KERNEL<double, true><<<grid_, kBlock, 0, st_>>>(kp, ##__VA_ARGS__);    \
            else KERNEL<double, false><<<grid_, kBlock, 0, st_>>>(kp, ##__VA_ARGS__);                     \
        }                                                                                                 \
        (h)->launches++;                                                                                  \
        CK(cudaGetLastError());                                                                           \
    } while (0)

static int validate(const DsimConfig *c) {
    if (!c || c->struct_size != (int)sizeof(DsimConfig)) return fail(nullptr, DSIM_EINVAL, "DsimConfig.struct_size mismatch%s", "");
    if (c->abi_version != DSIM_ABI_VERSION) return fail(nullptr, DSIM_EINVAL, "ABI version mismatch%s", "");
    if (c->num_envs <= 0) return fail(nullptr, DSIM_EINVAL, "num_envs must be positive%s", "");
    if (c->precision != DSIM_FP32 && c->precision != DSIM_FP64) return fail(nullptr, DSIM_EINVAL, "precision must be DSIM_FP32 or DSIM_FP64%s", "");
    if (c->obs_id < 0 || c->obs_id >= DSIM_NUM_OBS) return fail(nullptr, DSIM_EINVAL, "unknown obs_id%s", "");
    if (c->obs_id == DSIM_OBS_LOCAL_PRY_ACC_PARAMS_NOPEND)
        return fail(nullptr, DSIM_EUNSUPPORTED, "LocalFramePRYaccParamsNoPendEnv raises NameError in the reference (observation_wrappers.py:448)%s", "");
    if (c->reward_id < 0 || c->reward_id >= DSIM_NUM_REWARDS) return fail(nullptr, DSIM_EINVAL, "unknown reward_id%s", "");
    if (!c->pendulum) {
        if (!(c->obs_id == 0 || c->obs_id == DSIM_OBS_LOCAL_PRY_ACC_NOPEND))
            return fail(nullptr, DSIM_EUNSUPPORTED, "pendulum=False supports only BaseDroneEnv and LocalFramePRYaccNoPendEnv observations%s", "");
        if (!(c->reward_id <= 2 || c->reward_id == 10))
            return fail(nullptr, DSIM_EUNSUPPORTED, "pendulum=False supports only rewards that do not index pendulum state%s", "");
    }
    if (c->ground_contact != 0 && c->ground_contact != 1) return fail(nullptr, DSIM_EINVAL, "ground_contact must be 0 or 1%s", "");
    if (c->frame_skip < 1 || c->frequency <= 0) return fail(nullptr, DSIM_EINVAL, "frame_skip >= 1 and frequency > 0 required%s", "");
    return DSIM_OK;
}

template <typename T> static void fill_row(DsimHandle *h, void *base, int page_rows, int row, double value) {
    fill_row_kernel<T><<<(h->n + 255) / 256, 256>>>(h->n, (T *)base, page_rows, row, (T)value);
    h->launches++;
}

extern "C" int dsim_create(const DsimConfig *cfg, int device, DsimHandle **out) {
    if (!out) return fail(nullptr, DSIM_EINVAL, "out is NULL%s", "");
    *out = nullptr;
    int rc = validate(cfg);
    if (rc) return rc;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, DSIM_ECUDA, "no CUDA device available (%s): libdronesim_b200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, DSIM_EINVAL, "bad device ordinal%s", "");
    DsimHandle *h = new (std::nothrow) DsimHandle();
    if (!h) return fail(nullptr, DSIM_ENOMEM, "out of host memory%s", "");
    memset(h, 0, sizeof *h);
    h->cfg = *cfg; h->device = device; h->n = cfg->num_envs;
    h->npages = (cfg->num_envs + kTile - 1) / kTile; h->ld = h->npages * kTile;
    h->rs = cfg->precision == DSIM_FP32 ? 4 : 8;
    h->obs_dim = dsim_obs_dim(cfg->obs_id, cfg->pendulum);
    h->state_width = cfg->pendulum ? 33 : 29;
    h->h = cfg->round_precision ? round_prec5(1.0 / cfg->frequency) : 1.0 / cfg->frequency;
#define ALLOC(ptr, bytes)                                                                                  \
    do {                                                                                                   \
        const size_t g_ = h->guard ? kGuardBytes : 0, b_ = ((size_t)(bytes) + 15) / 16 * 16;               \
        char *base_ = nullptr;                                                                             \
        cudaError_t e2 = cudaMalloc((void **)&base_, b_ + 2 * g_);                                         \
        if (e2 == cudaSuccess && g_) e2 = cudaMemset(base_, 0xA5, b_ + 2 * g_);                            \
        if (e2 == cudaSuccess) e2 = cudaMemset(base_ + g_, 0, b_);                                         \
        if (e2 == cudaSuccess) { *(void **)&(ptr) = base_ + g_; if (g_ && h->ng < 16) { h->gbase[h->ng] = base_; h->gsize[h->ng++] = b_; } } \
        if (e2 != cudaSuccess) { fail(nullptr, e2 == cudaErrorMemoryAllocation ? DSIM_ENOMEM : DSIM_ECUDA, "allocation failed: %s", cudaGetErrorString(e2)); dsim_destroy(h); return e2 == cudaErrorMemoryAllocation ? DSIM_ENOMEM : DSIM_ECUDA; } \
    } while (0)
    if ((e = cudaSetDevice(device)) != cudaSuccess) { delete h; return fail(nullptr, DSIM_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e)); }
    if (const char *mb = getenv("DSIM_L2_PERSIST_MB")) {      // experiment knob: L2 set-aside for the evict_last (read-only page) lines
        int maxb = 0;
        cudaDeviceGetAttribute(&maxb, cudaDevAttrMaxPersistingL2CacheSize, device);
        size_t want = (size_t)atoi(mb) << 20;
        if (want > (size_t)maxb) want = (size_t)maxb;
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
        if (getenv("DSIM_VERBOSE")) fprintf(stderr, "dsim: persisting L2 set-aside %zu MB (max %d MB)\n", want >> 20, maxb >> 20);
    }
    const size_t ld = h->ld, n = h->n, rs = h->rs;
    h->guard = getenv("DSIM_GUARD") ? 1 : 0;                  // debug: exact-size buffers between canaries (compute-sanitizer is closed on the pool)
    ALLOC(h->rw, (size_t)RW_ROWS * ld * rs);
    ALLOC(h->ro, (size_t)RO_ROWS * ld * rs);
    ALLOC(h->refp, (size_t)REF_ROWS * ld * rs);
    ALLOC(h->obs, h->guard ? (size_t)h->obs_dim * n * rs : (size_t)DSIM_MAX_OBS * ld * rs);
    ALLOC(h->reward, (h->guard ? n : ld) * rs);
    ALLOC(h->states33, (size_t)h->state_width * n * rs);
    ALLOC(h->actions_stage, 4 * (h->guard ? n : ld) * rs);
    ALLOC(h->params64, 6 * ld * sizeof(double));
    if (cfg->ground_contact) ALLOC(h->geo, (size_t)GEO_ROWS * ld * rs);
    ALLOC(h->stats, (size_t)kStatReplicas * 8 * sizeof(double));
    ALLOC(h->center_hw, 12 * sizeof(double));
    ALLOC(h->trunc, h->guard ? n : ld);
    ALLOC(h->ticket, 256);
    if (getenv("DSIM_TIMELINE")) ALLOC(h->timeline, (size_t)h->npages * 8 * sizeof(unsigned long long));   // debug instrumentation
#undef ALLOC
    double chw[12];
    for (int k = 0; k < 6; k++) { chw[k] = cfg->param_center[k]; chw[6 + k] = cfg->param_halfwidth[k]; }
    CK(cudaMemcpy(h->center_hw, chw, sizeof chw, cudaMemcpyHostToDevice));
    // MjData qpos0: spawn grid, identity quaternions; per-env reference = shared reference
    if (h->rs == 4) spawn_grid_kernel<float><<<(h->n + 255) / 256, 256>>>(h->n, (float *)h->rw, cfg->start_pos[0], cfg->start_pos[1], cfg->start_pos[2]);
    else spawn_grid_kernel<double><<<(h->n + 255) / 256, 256>>>(h->n, (double *)h->rw, cfg->start_pos[0], cfg->start_pos[1], cfg->start_pos[2]);
    h->launches++;
    for (int k = 0; k < 4; k++) {
        const double v = k < 3 ? cfg->reference[k] - cfg->start_pos[k] : cfg->reference[3];
        if (h->rs == 4) fill_row<float>(h, h->refp, REF_ROWS, k, v); else fill_row<double>(h, h->refp, REF_ROWS, k, v);
    }
    CK(cudaGetLastError());
    rc = dsim_regen_params(h, 0, nullptr);
    if (rc) { snprintf(g_create_err, 512, "%s", h->err); dsim_destroy(h); return rc; }
    CK(cudaDeviceSynchronize());
    *out = h;
    return DSIM_OK;
}

extern "C" void dsim_destroy(DsimHandle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->host_pipeline_ready) {
        for (int i = 0; i < 3; i++) cudaStreamDestroy(h->hs[i]);
        for (int i = 0; i < kHostChunks; i++) { cudaEventDestroy(h->ev_in[i]); cudaEventDestroy(h->ev_k[i]); }
        cudaEventDestroy(h->ev_a); cudaEventDestroy(h->ev_b); cudaEventDestroy(h->ev_c);
    }
    void *ptrs[] = {h->rw, h->ro, h->refp, h->obs, h->reward, h->states33, h->actions_stage,
                    h->params64, h->stats, h->center_hw, h->trunc, h->timeline, h->ticket, h->geo};
    for (void *p : ptrs) if (p) cudaFree(h->guard ? (char *)p - kGuardBytes : (char *)p);
    delete h;
}

static int compile_on_device(DsimHandle *h, cudaStream_t st) {
    h->ro_dirty = 1;
    const int grid = (h->n + 127) / 128;
    if (h->rs == 4) compile_kernel<float><<<grid, 128, 0, st>>>(h->n, h->ld, h->params64, (float *)h->ro, h->cfg.pendulum, h->cfg.round_precision);
    else compile_kernel<double><<<grid, 128, 0, st>>>(h->n, h->ld, h->params64, (double *)h->ro, h->cfg.pendulum, h->cfg.round_precision);
    h->launches++;
    if (h->geo) {
        if (h->rs == 4) geometry_kernel<float><<<grid, 128, 0, st>>>(h->n, h->ld, h->params64, (float *)h->geo, h->cfg.pendulum, h->cfg.round_precision);
        else geometry_kernel<double><<<grid, 128, 0, st>>>(h->n, h->ld, h->params64, (double *)h->geo, h->cfg.pendulum, h->cfg.round_precision);
        h->launches++;
    }
    CK(cudaGetLastError());
    return DSIM_OK;
}

extern "C" int dsim_regen_params(DsimHandle *h, uint32_t epoch, void *stream) {
    if (!h) return DSIM_EINVAL;
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const DsimConfig &c = h->cfg;
    draw_params_kernel<<<(h->n + 127) / 128, 128, 0, st>>>(h->n, h->ld, h->params64, c.seed, (unsigned)c.env_id_offset, epoch,
                                                          c.random_params, c.pendulum, c.param_difficulty, h->center_hw);
    h->launches++;
    CK(cudaGetLastError());
    h->per_env_consts = c.random_params ? 1 : 0;
    if (!c.random_params) {
        for (int k = 0; k < 6; k++) h->uparams[k] = c.param_center[k];
        if (!c.pendulum) { h->uparams[4] = 0; h->uparams[5] = 0; }
        compile_consts(h->uparams, c.pendulum != 0, c.round_precision != 0, h->uconst);
    }
    return compile_on_device(h, st);
}

extern "C" int dsim_set_params(DsimHandle *h, const double *params_host, void *stream) {
    if (!h || !params_host) return DSIM_EINVAL;
    CK(cudaSetDevice(h->device));
    std::vector<double> t((size_t)6 * h->ld, 0.0);
    bool uniform = true;
    for (int i = 0; i < h->n; i++)
        for (int k = 0; k < 6; k++) {
            t[(size_t)k * h->ld + i] = params_host[(size_t)i * 6 + k];
            if (params_host[(size_t)i * 6 + k] != params_host[k]) uniform = false;
        }
    CK(cudaMemcpyAsync(h->params64, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    h->per_env_consts = uniform ? 0 : 1;
    if (uniform) {
        for (int k = 0; k < 6; k++) h->uparams[k] = params_host[k];
        compile_consts(h->uparams, h->cfg.pendulum != 0, h->cfg.round_precision != 0, h->uconst);
    }
    return compile_on_device(h, (cudaStream_t)stream);
}

extern "C" int dsim_get_params(DsimHandle *h, double *out) {
    if (!h || !out) return DSIM_EINVAL;
    CK(cudaSetDevice(h->device));
    std::vector<double> t((size_t)6 * h->ld);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(t.data(), h->params64, t.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int i = 0; i < h->n; i++) for (int k = 0; k < 6; k++) out[(size_t)i * 6 + k] = t[(size_t)k * h->ld + i];
    return DSIM_OK;
}

// whole paged buffer -> host vector
template <typename T> static cudaError_t fetch_pages(const DsimHandle *h, const void *dev, int rows, std::vector<T> &t) {
    t.resize((size_t)rows * h->ld);
    return cudaMemcpy(t.data(), dev, t.size() * sizeof(T), cudaMemcpyDeviceToHost);
}

extern "C" int dsim_get_consts(DsimHandle *h, double *out) {
    if (!h || !out) return DSIM_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    if (h->rs == 4) {
        std::vector<float> t;
        CK(fetch_pages(h, h->ro, RO_ROWS, t));
        for (int i = 0; i < h->n; i++) for (int k = 0; k < C_ROWS; k++) out[(size_t)i * C_ROWS + k] = t[page_elem(RO_ROWS, RO_CONSTS + k, i)];
    } else {
        std::vector<double> t;
        CK(fetch_pages(h, h->ro, RO_ROWS, t));
        for (int i = 0; i < h->n; i++) for (int k = 0; k < C_ROWS; k++) out[(size_t)i * C_ROWS + k] = t[page_elem(RO_ROWS, RO_CONSTS + k, i)];
    }
    return DSIM_OK;
}

extern "C" int dsim_forward(DsimHandle *h, int refresh_obs, void *stream) {
    if (!h) return DSIM_EINVAL;
    CK(cudaSetDevice(h->device));
    const void *actions_ = nullptr;
    DISPATCH(h, forward_kernel, stream, refresh_obs);
    return DSIM_OK;
}

static int reset_impl(DsimHandle *h, const uint8_t *mask, int single, void *stream) {
    CK(cudaSetDevice(h->device));
    const void *actions_ = nullptr;
    const int first = (!h->first_reset_done && !mask && single < 0) ? 1 : 0;
    DISPATCH(h, reset_kernel, stream, mask, single, first);
    return DSIM_OK;
}
extern "C" int dsim_reset_all(DsimHandle *h, void *stream) {
    if (!h) return DSIM_EINVAL;
    int rc = reset_impl(h, nullptr, -1, stream);
    if (rc) return rc;
    h->first_reset_done = 1;
    return dsim_forward(h, 1, stream);                 // reset_model: set_state -> mj_forward, then states/_get_obs (:324-326)
}
extern "C" int dsim_reset_masked(DsimHandle *h, const uint8_t *mask_dev, void *stream) {
    if (!h || !mask_dev) return DSIM_EINVAL;
    return reset_impl(h, mask_dev, -1, stream);
}
extern "C" int dsim_reset_at(DsimHandle *h, int index, void *stream) {
    if (!h) return DSIM_EINVAL;
    if (index < 0 || index >= h->n) return fail(h, DSIM_EINVAL, "reset_at: index out of range%s", "");   // assert index < num_drones (:338)
    return reset_impl(h, nullptr, index, stream);
}
extern "C" int dsim_zero_act(DsimHandle *h, void *stream) {
    if (!h) return DSIM_EINVAL;
    CK(cudaSetDevice(h->device));
    for (int r = S_ACT; r < RW_ROWS; r++) {                                  // act + sensordata of a fresh MjData
        if (r >= S_ROWS && r < S_ACC) continue;
        if (h->rs == 4) fill_row_kernel<float><<<(h->n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->n, (float *)h->rw, RW_ROWS, r, 0.0f);
        else fill_row_kernel<double><<<(h->n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->n, (double *)h->rw, RW_ROWS, r, 0.0);
        h->launches++;
    }
    CK(cudaGetLastError());
    return DSIM_OK;
}

// step-kernel dispatch: compile-time specialisations for the BASELINE configs, generic kernel otherwise.  The kernel is
// persistent: grid = min(CTAs needed, CTAs the GPU can hold at once).  The > 48 KB dynamic shared-memory opt-in and the
// occupancy are properties of ONE kernel instantiation on ONE device (and of the slot size), and one handle launches
// several instantiations (specialised plain steps, the generic kernel for dsim_evaluate / after a set_params that flips
// per_env_consts): both are looked up per (function, device, smem) in a small process-wide table.
struct StepFnInfo { const void *fn; int device; unsigned smem; int cap; };
static StepFnInfo g_step_fn[128];
static int g_step_fn_count = 0;
static std::mutex g_step_fn_mutex;
static cudaError_t step_fn_capacity(const void *fn, int device, unsigned smem, int *cap) {
    std::lock_guard<std::mutex> lock(g_step_fn_mutex);
    for (int k = 0; k < g_step_fn_count; k++)
        if (g_step_fn[k].fn == fn && g_step_fn[k].device == device && g_step_fn[k].smem == smem) { *cap = g_step_fn[k].cap; return cudaSuccess; }
    // opt in to the largest slot any handle can ask for
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(slot_bytes(DSIM_MAX_OBS, 8) * kStages * kStepWarps));
    if (e != cudaSuccess) return e;
    int sms = 0, per_sm = 0;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kStepBlock, smem)) != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    *cap = sms * per_sm;
    if (g_step_fn_count < 128) g_step_fn[g_step_fn_count++] = StepFnInfo{fn, device, smem, *cap};
    return cudaSuccess;
}
template <typename K> static cudaError_t launch_one(DsimHandle *h, K kernel, unsigned smem, cudaStream_t st, const void *kp_ptr, int pages) {
    int cap = 0;
    cudaError_t e = step_fn_capacity((const void *)kernel, h->device, smem, &cap);
    if (e != cudaSuccess) return e;
    const int need = (pages + kStepWarps - 1) / kStepWarps;
    const int grid = need < cap ? need : cap;
    void *args[] = {const_cast<void *>(kp_ptr)};
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof lc);
    lc.gridDim = dim3(grid); lc.blockDim = dim3(kStepBlock); lc.dynamicSmemBytes = smem; lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // pairs with griddepcontrol.* in step_kernel
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = 1;
    return cudaLaunchKernelExC(&lc, (const void *)kernel, args);
}
template <typename T> static cudaError_t launch_step(DsimHandle *h, const KParams<T> &kp, cudaStream_t st) {
    const unsigned smem = kp.smem_per_slot * kStages * kStepWarps;
    const int pages = kp.npages - kp.page0;
    if (h->cfg.ground_contact) {                                           // generic instantiation + floor-contact slow path
        if (!h->cfg.pendulum) return launch_one(h, step_kernel<T, false, -1, -1, -1, true>, smem, st, &kp, pages);
        return launch_one(h, step_kernel<T, true, -1, -1, -1, true>, smem, st, &kp, pages);
    }
    if (!h->cfg.pendulum) return launch_one(h, step_kernel<T, false, -1, -1>, smem, st, &kp, pages);
#ifdef DSIM_TL_ALL
    if (std::is_same<T, float>::value && !kp.eval_only) {
#else
    if (std::is_same<T, float>::value && !kp.eval_only && !kp.timeline) {   // specialised instantiations: plain steps only
#endif
        const int o = kp.obs_id, r = kp.reward_id;
        const int cfg = (kp.per_env_consts ? 1 : 0) | (kp.refp ? 2 : 0) | (kp.frame_skip == 1 ? 4 : 0);
        if constexpr (std::is_same<T, float>::value) {
            // two envs per lane on the packed FP32 pipe (dsim_step_x2.cuh) while a warp has at most two page pairs
            static const int x2 = getenv("DSIM_X2") ? atoi(getenv("DSIM_X2")) : 0;
            if (x2 && !kp.timeline) {
                const int npairs = (pages + 1) / 2;
                const unsigned smem2 = kp.smem_per_slot * kX2Slots * kStepWarps;
                cudaError_t e2 = cudaSuccess;
                auto try_x2 = [&](auto kernel) {
                    int cap = 0;
                    if (step_fn_capacity((const void *)kernel, h->device, smem2, &cap) != cudaSuccess) { cudaGetLastError(); return false; }
                    if (npairs > 2 * cap * kStepWarps) return false;
                    e2 = launch_one(h, kernel, smem2, st, &kp, npairs);
                    return true;
                };
                if (o == DSIM_OBS_LOCAL_RPY_PARAMS && r == 2 && cfg == 5 && try_x2(step_kernel_x2<DSIM_OBS_LOCAL_RPY_PARAMS, 2, 5>)) return e2;
                if (o == DSIM_OBS_LOCAL_RPY && r == 1 && cfg == 6 && try_x2(step_kernel_x2<DSIM_OBS_LOCAL_RPY, 1, 6>)) return e2;
            }
        }
        // BASELINE configs 4 / 5 (per-env randomised parameters), 3 (moving per-env setpoints, one parameter set), 2 (base_config: per-env parameters, raw 33-float rows)
        if (o == DSIM_OBS_LOCAL_RPY_PARAMS && r == 2 && cfg == 5) return launch_one(h, step_kernel<float, true, DSIM_OBS_LOCAL_RPY_PARAMS, 2, 5>, smem, st, &kp, pages);
        if (o == DSIM_OBS_LOCAL_RPY && r == 1 && cfg == 6) return launch_one(h, step_kernel<float, true, DSIM_OBS_LOCAL_RPY, 1, 6>, smem, st, &kp, pages);
        if (o == DSIM_OBS_BASE && r == 0 && cfg == 5) return launch_one(h, step_kernel<float, true, DSIM_OBS_BASE, 0, 5>, smem, st, &kp, pages);
    }
    return launch_one(h, step_kernel<T, true, -1, -1>, smem, st, &kp, pages);
}
static int step_impl(DsimHandle *h, const void *actions_dev, void *stream, int eval_only) {
    CK(cudaSetDevice(h->device));
    if (h->cfg.precision == DSIM_FP32) {
        auto kp = make_params<float>(h, actions_dev);
        if (eval_only) { kp.frame_skip = 0; kp.eval_only = 1; }
        CK(launch_step<float>(h, kp, (cudaStream_t)stream));
    } else {
        auto kp = make_params<double>(h, actions_dev);
        if (eval_only) { kp.frame_skip = 0; kp.eval_only = 1; }
        CK(launch_step<double>(h, kp, (cudaStream_t)stream));
    }
    h->launches++;
    h->ro_dirty = 0;
    CK(cudaGetLastError());
    return DSIM_OK;
}

extern "C" int dsim_step(DsimHandle *h, const void *actions_dev, void *stream) {
    if (!h || !actions_dev) return DSIM_EINVAL;
    if ((uintptr_t)actions_dev & 15u) return fail(h, DSIM_EINVAL, "actions must be 16-byte aligned (they are moved with bulk copies)%s", "");
    return step_impl(h, actions_dev, stream, 0);
}

extern "C" int dsim_evaluate(DsimHandle *h, const void *actions_dev, void *stream) {
    if (!h || !actions_dev) return DSIM_EINVAL;
    if ((uintptr_t)actions_dev & 15u) return fail(h, DSIM_EINVAL, "actions must be 16-byte aligned (they are moved with bulk copies)%s", "");
    return step_impl(h, actions_dev, stream, 1);
}

// vector_step with HOST buffers.  Pinned buffers take the zero-copy path in dsim_step_host.  Otherwise large batches are
// stepped in page ranges on three internal streams so that the PCIe legs overlap the kernel: H2D actions(k+1) | step
// kernel(k) | D2H observations(k-1).  Either way the observation read-back (88 B per env for the 22-float wrappers) is what
// bounds this path.
static int step_host_pipeline_init(DsimHandle *h) {
    if (h->host_pipeline_ready) return DSIM_OK;
    for (int i = 0; i < 3; i++) CK(cudaStreamCreateWithFlags(&h->hs[i], cudaStreamNonBlocking));
    for (int i = 0; i < kHostChunks; i++) {
        CK(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_k[i], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&h->ev_a, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_b, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_c, cudaEventDisableTiming));
    h->host_pipeline_ready = 1;
    return DSIM_OK;
}

// page ranges of the host pipeline: two small leading chunks so the read-back starts early, then doubling sizes so that
// few copy boundaries remain (1/16, 1/16, 1/8, 1/4, 1/2 of the pages).  Returns the chunk count; bounds[k] .. bounds[k+1].
static int host_chunk_bounds(int npages, int *bounds) {
    static const int sixteenths[] = {0, 1, 2, 4, 8, 16};
    int nb = 0, last = -1;
    for (int k = 0; k < 6; k++) {
        const int b = (int)((long long)npages * sixteenths[k] / 16);
        if (b == last) continue;
        if (bounds) bounds[nb] = b;
        nb++; last = b;
    }
    return nb - 1;
}

// fork from `root` onto the copy-in / copy-out streams, queue every chunk, join back into `root` (also valid under capture)
static int enqueue_host_pipeline(DsimHandle *h, const KParams<float> &base, const float *actions_host, float *obs_host, float *reward_host,
                                 uint8_t *trunc_host, cudaStream_t root) {
    int bounds[kHostChunks + 1];
    const int chunks = host_chunk_bounds(h->npages, bounds);
    cudaStream_t s_in = h->hs[0], s_out = h->hs[2];
    CK(cudaEventRecord(h->ev_c, root));
    CK(cudaStreamWaitEvent(s_in, h->ev_c, 0));
    CK(cudaStreamWaitEvent(s_out, h->ev_c, 0));
    const size_t D = (size_t)h->obs_dim;
    float *act_dev = (float *)h->actions_stage, *obs_dev = (float *)h->obs;
    for (int k = 0; k < chunks; k++) {
        const int p0 = bounds[k], p1 = bounds[k + 1];
        const size_t e0 = (size_t)p0 * kTile, e1 = (size_t)p1 * kTile < (size_t)h->n ? (size_t)p1 * kTile : (size_t)h->n, cnt = e1 - e0;
        CK(cudaMemcpyAsync(act_dev + e0 * 4, actions_host + e0 * 4, cnt * 4 * sizeof(float), cudaMemcpyHostToDevice, s_in));
        CK(cudaEventRecord(h->ev_in[k], s_in));
        CK(cudaStreamWaitEvent(root, h->ev_in[k], 0));
        KParams<float> kp = base;
        kp.page0 = p0; kp.npages = p1; kp.ticket = h->ticket + 4 * k;       // its own work-stealing counters
        kp.early_in = 0;                                                    // chunks of one step: strict ordering (the action rows arrive through event-ordered copies)
        CK(launch_step<float>(h, kp, root));
        CK(cudaEventRecord(h->ev_k[k], root));
        CK(cudaStreamWaitEvent(s_out, h->ev_k[k], 0));
        if (obs_host) CK(cudaMemcpyAsync(obs_host + e0 * D, obs_dev + e0 * D, cnt * D * sizeof(float), cudaMemcpyDeviceToHost, s_out));
    }
    if (reward_host) CK(cudaMemcpyAsync(reward_host, h->reward, (size_t)h->n * sizeof(float), cudaMemcpyDeviceToHost, s_out));
    if (trunc_host) CK(cudaMemcpyAsync(trunc_host, h->trunc, (size_t)h->n, cudaMemcpyDeviceToHost, s_out));
    CK(cudaEventRecord(h->ev_b, s_out));
    CK(cudaStreamWaitEvent(root, h->ev_b, 0));                              // s_in joined through the per-chunk events already
    return DSIM_OK;
}

extern "C" int dsim_step_host(DsimHandle *h, const float *actions_host, float *obs_host, float *reward_host, uint8_t *trunc_host, void *stream) {
    if (!h || !actions_host) return DSIM_EINVAL;
    if (h->cfg.precision != DSIM_FP32) return fail(h, DSIM_EUNSUPPORTED, "dsim_step_host moves float32 buffers; use dsim_step with precision=FP64%s", "");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    // Zero-copy path: every buffer the caller passed is pinned (cudaHostAlloc / cudaHostRegister / torch pin_memory), i.e.
    // mapped into the device address space.  ONE launch on the caller's stream: the kernel's bulk loads fetch the action rows
    // over PCIe and its bulk stores write observations / reward / truncated to the host buffers as posted PCIe writes, page
    // by page while later pages are still being computed - no copy-engine pass, no chunk boundaries, no staging.  The
    // device-resident outputs are written as well (dsim_buffer views stay current).
    {
        auto mapped = [](const void *q, size_t align, void **dp) {
            *dp = nullptr;
            if (!q) return true;
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, q) != cudaSuccess) { cudaGetLastError(); return false; }
            if (at.type != cudaMemoryTypeHost || !at.devicePointer || ((uintptr_t)at.devicePointer & (align - 1))) return false;
            *dp = at.devicePointer;
            return true;
        };
        void *d_act, *d_obs, *d_rew, *d_tr;
        if (mapped(actions_host, 16, &d_act) && mapped(obs_host, 16, &d_obs) && mapped(reward_host, 4, &d_rew) && mapped(trunc_host, 1, &d_tr)) {
            KParams<float> kp = make_params<float>(h, d_act);
            kp.obs_host = (float *)d_obs; kp.reward_host = (float *)d_rew; kp.trunc_host = (unsigned char *)d_tr;
            CK(launch_step<float>(h, kp, st));
            h->launches++;
            h->ro_dirty = 0;
            CK(cudaStreamSynchronize(st));
            return DSIM_OK;
        }
    }
    // Pageable host memory: staged copies.
    const int chunks = h->npages >= 2048 ? kHostChunks : 1;                 // >= 65536 envs: worth pipelining
    if (chunks == 1) {
        CK(cudaMemcpyAsync(h->actions_stage, actions_host, (size_t)h->n * 4 * sizeof(float), cudaMemcpyHostToDevice, st));
        int rc = dsim_step(h, h->actions_stage, stream);
        if (rc) return rc;
        if (obs_host) CK(cudaMemcpyAsync(obs_host, h->obs, (size_t)h->n * h->obs_dim * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (reward_host) CK(cudaMemcpyAsync(reward_host, h->reward, (size_t)h->n * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (trunc_host) CK(cudaMemcpyAsync(trunc_host, h->trunc, (size_t)h->n, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return DSIM_OK;
    }
    int rc = step_host_pipeline_init(h);
    if (rc) return rc;
    KParams<float> kp = make_params<float>(h, h->actions_stage);
    CK(cudaEventRecord(h->ev_a, st));
    CK(cudaStreamWaitEvent(h->hs[1], h->ev_a, 0));
    rc = enqueue_host_pipeline(h, kp, actions_host, obs_host, reward_host, trunc_host, h->hs[1]);
    if (rc) return rc;
    h->launches += host_chunk_bounds(h->npages, nullptr);
    CK(cudaEventRecord(h->ev_b, h->hs[1]));
    CK(cudaStreamWaitEvent(st, h->ev_b, 0));
    CK(cudaStreamSynchronize(h->hs[1]));
    return DSIM_OK;
}

extern "C" int dsim_set_inputs_ready(DsimHandle *h, int ready) {
    if (!h) return DSIM_EINVAL;
    h->inputs_ready = ready ? 1 : 0;
    return DSIM_OK;
}

extern "C" int dsim_set_reference(DsimHandle *h, const double ref[4]) {
    if (!h || !ref) return DSIM_EINVAL;
    for (int k = 0; k < 4; k++) h->cfg.reference[k] = ref[k];
    return DSIM_OK;
}

extern "C" int dsim_control_reference(DsimHandle *h, const void *axes_dev, void *stream) {
    if (!h || !axes_dev) return DSIM_EINVAL;
    if (!h->cfg.per_env_reference) return fail(h, DSIM_EUNSUPPORTED, "dsim_control_reference needs per_env_reference=1%s", "");
    CK(cudaSetDevice(h->device));
    const int grid = (h->n + 127) / 128;
    if (h->rs == 4) control_reference_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>(h->n, h->ld, (const float *)axes_dev, (float *)h->refp);
    else control_reference_kernel<double><<<grid, 128, 0, (cudaStream_t)stream>>>(h->n, h->ld, (const double *)axes_dev, (double *)h->refp);
    h->launches++;
    CK(cudaGetLastError());
    return DSIM_OK;
}

extern "C" int dsim_trajectory_reference(DsimHandle *h, int kind, double t, double phase_step, const double params[3],
                                         const double start_pos[4], const double end_pos[4], void *stream) {
    if (!h || !params) return DSIM_EINVAL;
    if (!h->cfg.per_env_reference) return fail(h, DSIM_EUNSUPPORTED, "dsim_trajectory_reference needs per_env_reference=1%s", "");
    if (kind < DSIM_TRAJ_CIRCLE || kind > DSIM_TRAJ_RAMP) return fail(h, DSIM_EINVAL, "unknown trajectory kind%s", "");
    if (kind != DSIM_TRAJ_CIRCLE && (!start_pos || !end_pos)) return DSIM_EINVAL;
    CK(cudaSetDevice(h->device));
    const double z4[4] = {0, 0, 0, 0};
    const double *s = start_pos ? start_pos : z4, *e = end_pos ? end_pos : z4, *c = h->cfg.start_pos;
    const int grid = (h->n + 127) / 128;
    if (h->rs == 4) trajectory_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>(h->n, (float *)h->refp, kind, t, phase_step, params[0], params[1], params[2],
                                                                                     s[0], s[1], s[2], s[3], e[0], e[1], e[2], e[3], c[0], c[1], c[2]);
    else trajectory_kernel<double><<<grid, 128, 0, (cudaStream_t)stream>>>(h->n, (double *)h->refp, kind, t, phase_step, params[0], params[1], params[2],
                                                                           s[0], s[1], s[2], s[3], e[0], e[1], e[2], e[3], c[0], c[1], c[2]);
    h->launches++;
    CK(cudaGetLastError());
    return DSIM_OK;
}

template <typename T> static T int_slot(int v) {
    T out;
    if (sizeof(T) == 4) { int32_t x = v; memcpy(&out, &x, 4); } else { int64_t x = v; memcpy(&out, &x, 8); }
    return out;
}
template <typename T> static int slot_int(T v) {
    if (sizeof(T) == 4) { int32_t x; memcpy(&x, &v, 4); return x; }
    int64_t x; memcpy(&x, &v, 8); return (int)x;
}
template <typename T>
static void pack_state(const DsimHandle *h, std::vector<T> &t, const double *qpos, const double *qvel, const double *act, const int32_t *num_steps) {
    const int nq = h->cfg.pendulum ? 9 : 7, nv = h->cfg.pendulum ? 8 : 6;
    for (int i = 0; i < h->n; i++) {
        T *c = t.data() + page_elem(RW_ROWS, 0, i);
        if (qpos) {
            const double *q = qpos + (size_t)i * nq;
            for (int k = 0; k < 3; k++) c[k * kTile] = (T)(q[k] - h->cfg.start_pos[k]);
            for (int k = 0; k < 4; k++) c[(S_QUAT + k) * kTile] = (T)q[3 + k];
            if (nq == 9) { c[S_HINGE * kTile] = (T)q[7]; c[(S_HINGE + 1) * kTile] = (T)q[8]; }
        }
        if (qvel) {
            const double *v = qvel + (size_t)i * nv;
            for (int k = 0; k < 6; k++) c[(S_VEL + k) * kTile] = (T)v[k];
            if (nv == 8) { c[S_HVEL * kTile] = (T)v[6]; c[(S_HVEL + 1) * kTile] = (T)v[7]; }
        }
        if (act) for (int k = 0; k < 4; k++) c[(S_ACT + k) * kTile] = (T)act[(size_t)i * 4 + k];
        if (num_steps) c[RW_NUM_STEPS * kTile] = int_slot<T>(num_steps[i]);
    }
}
template <typename T>
static void unpack_state(const DsimHandle *h, const std::vector<T> &t, double *qpos, double *qvel, double *act, double *sens, int32_t *num_steps) {
    const int nq = h->cfg.pendulum ? 9 : 7, nv = h->cfg.pendulum ? 8 : 6;
    for (int i = 0; i < h->n; i++) {
        const T *c = t.data() + page_elem(RW_ROWS, 0, i);
        if (qpos) {
            double *q = qpos + (size_t)i * nq;
            for (int k = 0; k < 3; k++) q[k] = h->cfg.start_pos[k] + (double)c[k * kTile];
            for (int k = 0; k < 4; k++) q[3 + k] = (double)c[(S_QUAT + k) * kTile];
            if (nq == 9) { q[7] = (double)c[S_HINGE * kTile]; q[8] = (double)c[(S_HINGE + 1) * kTile]; }
        }
        if (qvel) {
            double *v = qvel + (size_t)i * nv;
            for (int k = 0; k < 6; k++) v[k] = (double)c[(S_VEL + k) * kTile];
            if (nv == 8) { v[6] = (double)c[S_HVEL * kTile]; v[7] = (double)c[(S_HVEL + 1) * kTile]; }
        }
        if (act) for (int k = 0; k < 4; k++) act[(size_t)i * 4 + k] = (double)c[(S_ACT + k) * kTile];
        if (sens) for (int k = 0; k < 3; k++) sens[(size_t)i * 3 + k] = (double)c[(S_ACC + k) * kTile];
        if (num_steps) num_steps[i] = slot_int<T>(c[RW_NUM_STEPS * kTile]);
    }
}

extern "C" int dsim_set_state(DsimHandle *h, const double *qpos, const double *qvel, const double *act, const int32_t *num_steps, void *stream) {
    if (!h) return DSIM_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    if (h->rs == 4) {
        std::vector<float> t;
        CK(fetch_pages(h, h->rw, RW_ROWS, t));
        pack_state(h, t, qpos, qvel, act, num_steps);
        CK(cudaMemcpy(h->rw, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
    } else {
        std::vector<double> t;
        CK(fetch_pages(h, h->rw, RW_ROWS, t));
        pack_state(h, t, qpos, qvel, act, num_steps);
        CK(cudaMemcpy(h->rw, t.data(), t.size() * 8, cudaMemcpyHostToDevice));
    }
    return DSIM_OK;
}

extern "C" int dsim_get_state(DsimHandle *h, double *qpos, double *qvel, double *act, double *sens, int32_t *num_steps) {
    if (!h) return DSIM_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    if (h->rs == 4) {
        std::vector<float> t;
        CK(fetch_pages(h, h->rw, RW_ROWS, t));
        unpack_state(h, t, qpos, qvel, act, sens, num_steps);
    } else {
        std::vector<double> t;
        CK(fetch_pages(h, h->rw, RW_ROWS, t));
        unpack_state(h, t, qpos, qvel, act, sens, num_steps);
    }
    return DSIM_OK;
}

extern "C" int dsim_compute_states(DsimHandle *h, void *stream) {
    if (!h) return DSIM_EINVAL;
    CK(cudaSetDevice(h->device));
    const int grid = (h->n + kBlock - 1) / kBlock;
    cudaStream_t st = (cudaStream_t)stream;
    if (h->rs == 4) {
        auto kp = make_params<float>(h, nullptr);
        if (h->cfg.pendulum) states_kernel<float, true><<<grid, kBlock, 0, st>>>(kp, (float *)h->states33, h->state_width);
        else states_kernel<float, false><<<grid, kBlock, 0, st>>>(kp, (float *)h->states33, h->state_width);
    } else {
        auto kp = make_params<double>(h, nullptr);
        if (h->cfg.pendulum) states_kernel<double, true><<<grid, kBlock, 0, st>>>(kp, (double *)h->states33, h->state_width);
        else states_kernel<double, false><<<grid, kBlock, 0, st>>>(kp, (double *)h->states33, h->state_width);
    }
    h->launches++;
    CK(cudaGetLastError());
    return DSIM_OK;
}

extern "C" int dsim_buffer(DsimHandle *h, int id, void **ptr, int64_t *rows, int64_t *cols, int64_t *ld, int32_t *dtype, int64_t *page_rows) {
    if (!h || !ptr || !rows || !cols || !ld || !dtype || !page_rows) return DSIM_EINVAL;
    const int32_t rdt = h->rs == 4 ? DSIM_DT_F32 : DSIM_DT_F64, idt = h->rs == 4 ? DSIM_DT_I32 : DSIM_DT_I64;
    const int64_t n = h->n, L = h->ld;
    *page_rows = 0;
    char *rw = (char *)h->rw, *ro = (char *)h->ro;
    const size_t row_bytes = (size_t)kTile * h->rs;
    switch (id) {
    case DSIM_BUF_STATE: *ptr = rw; *rows = S_ROWS; *cols = n; *ld = L; *dtype = rdt; *page_rows = RW_ROWS; break;
    case DSIM_BUF_NUM_STEPS: *ptr = rw + RW_NUM_STEPS * row_bytes; *rows = 1; *cols = n; *ld = L; *dtype = idt; *page_rows = RW_ROWS; break;
    case DSIM_BUF_EP_RETURN: *ptr = rw + RW_EP_RETURN * row_bytes; *rows = 1; *cols = n; *ld = L; *dtype = rdt; *page_rows = RW_ROWS; break;
    case DSIM_BUF_CONSTS: *ptr = ro + RO_CONSTS * row_bytes; *rows = C_ROWS; *cols = n; *ld = L; *dtype = rdt; *page_rows = RO_ROWS; break;
    case DSIM_BUF_PARAMS: *ptr = ro + RO_PARAMS * row_bytes; *rows = 6; *cols = n; *ld = L; *dtype = rdt; *page_rows = RO_ROWS; break;
    case DSIM_BUF_REFERENCE: *ptr = h->refp; *rows = REF_ROWS; *cols = n; *ld = L; *dtype = rdt; *page_rows = REF_ROWS; break;
    case DSIM_BUF_OBS: *ptr = h->obs; *rows = n; *cols = h->obs_dim; *ld = h->obs_dim; *dtype = rdt; break;
    case DSIM_BUF_REWARD: *ptr = h->reward; *rows = 1; *cols = n; *ld = L; *dtype = rdt; break;
    case DSIM_BUF_TRUNCATED: *ptr = h->trunc; *rows = 1; *cols = n; *ld = L; *dtype = DSIM_DT_U8; break;
    case DSIM_BUF_RESET_COUNT: *ptr = rw + RW_RESET_COUNT * row_bytes; *rows = 1; *cols = n; *ld = L; *dtype = idt; *page_rows = RW_ROWS; break;
    case DSIM_BUF_STATES33: *ptr = h->states33; *rows = n; *cols = h->state_width; *ld = h->state_width; *dtype = rdt; break;
    case DSIM_BUF_SENSORDATA: *ptr = rw + S_ACC * row_bytes; *rows = 3; *cols = n; *ld = L; *dtype = rdt; *page_rows = RW_ROWS; break;
    case DSIM_BUF_GEOMETRY:
        if (!h->geo) return fail(h, DSIM_EINVAL, "no collision geometry: the handle was created without ground_contact%s", "");
        *ptr = h->geo; *rows = GEO_ROWS; *cols = n; *ld = L; *dtype = rdt; break;
    case DSIM_BUF_STATS: *ptr = h->stats; *rows = kStatReplicas; *cols = 8; *ld = 8; *dtype = DSIM_DT_F64; break;
    default: return fail(h, DSIM_EINVAL, "unknown buffer id%s", "");
    }
    return DSIM_OK;
}

extern "C" int dsim_stats(DsimHandle *h, double out[8], int reset) {
    if (!h || !out) return DSIM_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    double t[kStatReplicas * 8];
    CK(cudaMemcpy(t, h->stats, sizeof t, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 8; k++) { out[k] = 0; for (int r = 0; r < kStatReplicas; r++) out[k] += t[r * 8 + k]; }
    if (reset) CK(cudaMemset(h->stats, 0, sizeof t));
    return DSIM_OK;
}

extern "C" int dsim_sync(DsimHandle *h, void *stream) {
    if (!h) return DSIM_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return DSIM_OK;
}

// debug: per-warp %globaltimer stamps of the last step launch (only when the handle was created with DSIM_TIMELINE set)
extern "C" int dsim_debug_timeline(DsimHandle *h, uint64_t *out, int64_t capacity) {
    if (!h || !out) return DSIM_EINVAL;
    if (!h->timeline) return fail(h, DSIM_EUNSUPPORTED, "create the handle with DSIM_TIMELINE set in the environment%s", "");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    const int64_t cnt = (int64_t)h->npages * 8;
    CK(cudaMemcpy(out, h->timeline, (size_t)(cnt < capacity ? cnt : capacity) * 8, cudaMemcpyDeviceToHost));
    return (int)(cnt < capacity ? cnt : capacity) > 0 ? DSIM_OK : DSIM_EINVAL;
}

// MyBetaDist sampling / mean + log-probability for the policy head's logits (distributions.py:6-38); no handle needed
extern "C" int dsim_beta_policy(const void *logits_dev, int n, int precision, uint32_t seed, int64_t env_id_offset, uint32_t step,
                                const uint32_t *step_dev, int deterministic, void *actions_dev, void *logp_dev, void *stream) {
    if (!logits_dev || !actions_dev || n <= 0) return DSIM_EINVAL;
    if (precision != DSIM_FP32 && precision != DSIM_FP64) return DSIM_EINVAL;
    const int grid = (n + 127) / 128;
    cudaStream_t st = (cudaStream_t)stream;
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof lc);
    lc.gridDim = dim3(grid); lc.blockDim = dim3(128); lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // pairs with griddepcontrol.* at the top of the kernel
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = 1;
    uint32_t env0 = (uint32_t)env_id_offset;
    void *args[] = {&n, &logits_dev, &seed, &env0, &step, &step_dev, &deterministic, &actions_dev, &logp_dev};
    const void *fn = precision == DSIM_FP32 ? (const void *)beta_policy_kernel<float, 4> : (const void *)beta_policy_kernel<double, 4>;
    if (cudaLaunchKernelExC(&lc, fn, args) != cudaSuccess) return DSIM_ECUDA;
    return cudaGetLastError() == cudaSuccess ? DSIM_OK : DSIM_ECUDA;
}

// debug (handles created with DSIM_GUARD=1): number of canary bytes around the handle's device buffers that no longer hold
// their pattern = out-of-bounds device writes since dsim_create; -1 when the handle has no canaries
extern "C" int64_t dsim_debug_guard_check(DsimHandle *h) {
    if (!h || !h->guard) return -1;
    if (cudaSetDevice(h->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return -2;
    std::vector<unsigned char> t(kGuardBytes);
    int64_t bad = 0;
    for (int k = 0; k < h->ng; k++)
        for (int side = 0; side < 2; side++) {
            const char *src = (const char *)h->gbase[k] + (side ? kGuardBytes + h->gsize[k] : 0);
            if (cudaMemcpy(t.data(), src, kGuardBytes, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
            for (unsigned char c : t) bad += c != 0xA5;
        }
    return bad;
}

extern "C" int64_t dsim_launch_count(const DsimHandle *h) { return h ? h->launches : 0; }

extern "C" int dsim_kernel_info(int which, int32_t *regs, int32_t *local_bytes, int32_t *max_threads) {
    cudaFuncAttributes a;
    cudaError_t e = which == 0 ? cudaFuncGetAttributes(&a, step_kernel<float, true, DSIM_OBS_LOCAL_RPY_PARAMS, 2, 5>) : cudaFuncGetAttributes(&a, step_kernel<double, true, -1, -1>);
    if (e != cudaSuccess) return DSIM_ECUDA;
    if (regs) *regs = a.numRegs;
    if (local_bytes) *local_bytes = (int32_t)a.localSizeBytes;
    if (max_threads) *max_threads = a.maxThreadsPerBlock;
    return DSIM_OK;
}
