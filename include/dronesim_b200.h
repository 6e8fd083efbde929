/*
 * dronesim_b200.h — C ABI of libdronesim_b200.so: the B200-native replacement for the vectorised
 * env-step path of TichyTech/mujoco-drone.
 *
 * The reference has NO native FFI (it is pure Python on top of the `mujoco` wheel); the boundary this
 * library is bound behind is the RLlib VectorEnv surface of environments/BaseDroneEnv.py.  Each entry
 * point cites the reference interface it replaces.  Plain pointers and sizes only; no torch types, no
 * exceptions.  Every int-returning function returns DSIM_OK (0) or a negative DSIM_E* code and leaves a
 * message retrievable with dsim_last_error().  One handle = one shard of envs on one GPU, one caller
 * thread, one stream at a time (reference threading model: one env object per Ray worker process,
 * BaseDroneEnv.py:53).
 *
 * Device memory layout (all buffers owned by the handle, valid until dsim_destroy).  `real` = float, or double when
 * precision == DSIM_FP64.  Per-env data lives in PAGES of 32 envs (one warp of the step kernel); a page of a buffer
 * with R rows is the contiguous block real[R][32], so ONE 1D bulk copy (TMA) moves a warp's whole working set:
 *     element (row r, env i)  =  base[(i / 32) * R * 32 + r * 32 + (i % 32)]          ("PAGED", R = page_rows)
 *   read-write page (R = 27):  rows 0-20 state = MuJoCo's qpos[9], qvel[8], act[4] per drone
 *           (BaseDroneEnv.py:367-375): 0-2 position OFFSET from start_pos[0:3] | 3-6 quat (w,x,y,z) | 7-8 hinge angles
 *           | 9-11 world linear velocity | 12-14 body angular velocity | 15-16 hinge rates | 17-20 `act`;
 *           row 21 BaseDroneEnv.num_steps (integer of the width of `real`), row 22 running episode return,
 *           row 23 reset count (integer; Philox epoch of the env's reset stream);
 *           rows 24-26 sensordata (accelerometer): written by every step, never read by one - the step kernel fetches
 *           rows 0-23 only and stores all 27
 *   read-only page  (R = 19):  rows 0-12 compiled rigid-body constants, rows 13-18 raw drone_params
 *   setpoint page   (R = 4):   x, y, z offsets from start_pos and yaw (only read when per_env_reference)
 *   ld = num_envs rounded up to a multiple of 32.
 */
#ifndef DRONESIM_B200_H
#define DRONESIM_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSIM_ABI_VERSION 3

enum { DSIM_OK = 0, DSIM_EINVAL = -1, DSIM_ECUDA = -2, DSIM_ENOMEM = -3, DSIM_EUNSUPPORTED = -4, DSIM_ESHAPE = -5 };
enum { DSIM_FP32 = 0, DSIM_FP64 = 1 };
enum { DSIM_NSTATE_ROWS = 21, DSIM_NCONST = 13, DSIM_NPARAM = 6, DSIM_MAX_OBS = 33 };

/* buffer ids for dsim_buffer() */
enum {
    DSIM_BUF_STATE = 0,      /* real  PAGED 21 rows of the read-write page       data.qpos/qvel/act */
    DSIM_BUF_NUM_STEPS = 1,  /* int   PAGED 1 row (int32 | int64 as `real`)      BaseDroneEnv.num_steps (:110) */
    DSIM_BUF_OBS = 2,        /* real  [N][obs_dim] policy-ready rows             vector_step()[0] */
    DSIM_BUF_REWARD = 3,     /* real  [N]                                        vector_step()[1] */
    DSIM_BUF_TRUNCATED = 4,  /* uint8 [N]                                        vector_step()[3] */
    DSIM_BUF_PARAMS = 5,     /* real  PAGED 6 rows of the read-only page         BaseDroneEnv.drone_params (:117) */
    DSIM_BUF_CONSTS = 6,     /* real  PAGED 13 rows of the read-only page        (MjModel of env_gen.py) */
    DSIM_BUF_REFERENCE = 7,  /* real  PAGED 4 rows: per-env setpoint (xyz offset from start_pos, yaw); only if per_env_reference */
    DSIM_BUF_RESET_COUNT = 8,/* int   PAGED 1 row: Philox epoch of each env's reset stream */
    DSIM_BUF_STATES33 = 9,   /* real  [N][33|29] get_drone_states() rows, filled by dsim_compute_states */
    DSIM_BUF_EP_RETURN = 10, /* real  PAGED 1 row: running return of the current episode */
    DSIM_BUF_STATS = 11,     /* double[64][8] partial sums (dsim_stats adds the 64 rows): sum_return, sum_length, n_episodes, n_nonfinite, n_near_ground, 0,0,0 */
    DSIM_BUF_SENSORDATA = 12, /* real PAGED 3 rows of the read-write page        data.sensordata (accelerometer) */
    DSIM_BUF_GEOMETRY = 13   /* real  [39][ld]  collision geometry of every drone (handles created with ground_contact; csrc/dsim_contact.cuh GeoRow) */
};
enum { DSIM_DT_F32 = 0, DSIM_DT_F64 = 1, DSIM_DT_I32 = 2, DSIM_DT_U8 = 3, DSIM_DT_U32 = 4, DSIM_DT_I64 = 5 };

/* observation variants: class names of environments/observation_wrappers.py (ids == oracle OBS_IDS) */
enum {
    DSIM_OBS_BASE = 0, DSIM_OBS_GLOBAL_RPY = 1, DSIM_OBS_LOCAL_PRY = 2, DSIM_OBS_LOCAL_FULLSTATE = 3,
    DSIM_OBS_LOCAL_FULLSTATE_ZVEC = 4, DSIM_OBS_LOCAL_PRY_ACC = 5, DSIM_OBS_LOCAL_PRY_PARAMS = 6,
    DSIM_OBS_LOCAL_PRY_ACC_PARAMS = 7, DSIM_OBS_LOCAL_RPY_PARAMS = 8, DSIM_OBS_LOCAL_RPY_FAKEPARAMS = 9,
    DSIM_OBS_LOCAL_RPY = 10, DSIM_OBS_LOCAL_PRY_ACC_NOPEND = 11, DSIM_OBS_LOCAL_PRY_ACC_PARAMS_NOPEND = 12 /* raises in the reference */,
    DSIM_OBS_LOCAL_RM_PARAMS = 13, DSIM_OBS_LOCAL_ZVEC = 14, DSIM_NUM_OBS = 15
};
/* reward functions of environments/rewards.py in file order (ids == oracle REWARD_IDS) */
enum { DSIM_NUM_REWARDS = 17 };

/* Everything BaseDroneEnv.__init__ reads from `config` (BaseDroneEnv.py:60-106), already resolved to numbers. */
typedef struct DsimConfig {
    int32_t struct_size;           /* sizeof(DsimConfig): ABI guard */
    int32_t abi_version;           /* DSIM_ABI_VERSION */
    int32_t num_envs;              /* drones on THIS device (config['num_drones'] of this shard) */
    int32_t precision;             /* DSIM_FP32 | DSIM_FP64 (reference arithmetic is FP64) */
    int64_t env_id_offset;         /* global id of local env 0: RNG streams are keyed by the global id */
    uint32_t seed;                 /* config['seed'] */
    int32_t pendulum;              /* config['pendulum'] */
    int32_t frame_skip;            /* config['skip_steps'] */
    int32_t round_precision;       /* 1: apply mjcf precision=5 ("%.5g") to every model attribute (env_gen.py:129) */
    double frequency;              /* config['frequency'] -> timestep = 1/frequency (env_gen.py:82) */
    int32_t obs_id, reward_id;     /* wrapper class / config['reward_fcn'] resolved by name */
    int32_t ground_contact;        /* 1: contacts of the drone's geoms with the floor plane z = 0 are simulated (env_gen.py:14-21,97:
                                      plane vs box / cylinder / sphere, condim 3, pyramidal cone, MuJoCo's default solref / solimp);
                                      0: the floor is out of reach and only would-be contacts are counted (n_near_ground) */
    int32_t per_env_reference;     /* 0: one reference shared by all drones (BaseDroneEnv.py:80) */
    int32_t auto_reset;            /* 1: truncated envs are re-sampled inside the step kernel (native loop) */
    int32_t random_start_pos, random_params;
    double reference[4], start_pos[4];
    double max_distance;
    int64_t max_steps;
    double max_pos_offset;         /* state_difficulty * max_random_offset */
    double angle_sigma[2], vel_sigma[3], ang_vel_sigma[3], pend_rp_sigma[2], pend_vel_sigma[2]; /* already * state_difficulty */
    double param_center[6], param_halfwidth[6], param_difficulty;   /* *_interval in drone_params order */
} DsimConfig;

typedef struct DsimHandle DsimHandle;

/* -- lifetime: BaseDroneEnv.__init__ (:55-149) / close() */
int dsim_abi_version(void);
int dsim_create(const DsimConfig *cfg, int device, DsimHandle **out);
void dsim_destroy(DsimHandle *h);
const char *dsim_last_error(const DsimHandle *h /* NULL: error of the last failed dsim_create */);
int dsim_obs_dim(int obs_id, int pendulum);     /* observation_space.shape[0] actually EMITTED (Q14: 24 for FullStateZvec) */

/* -- model: generate_drone_params (:180-216) + env_gen.make_sim/mjcf_to_mjmodel (env_gen.py:7-133) */
int dsim_regen_params(DsimHandle *h, uint32_t epoch, void *stream);             /* Philox draw + compile, on device */
int dsim_set_params(DsimHandle *h, const double *params_host /*[N][6]*/, void *stream); /* explicit drone_params + compile */
int dsim_get_params(DsimHandle *h, double *params_host /*[N][6]*/);
int dsim_get_consts(DsimHandle *h, double *consts_host /*[N][13]*/);

/* -- reset: reset_model (:296-326), vector_reset (:328-332), reset_at (:334-351); set_state -> mj_forward */
int dsim_reset_all(DsimHandle *h, void *stream);                                /* sample all, num_steps=0, forward, obs */
int dsim_reset_masked(DsimHandle *h, const uint8_t *mask_dev, void *stream);    /* sample where mask!=0, num_steps=0 (obs untouched: Q1) */
int dsim_reset_at(DsimHandle *h, int index, void *stream);
int dsim_forward(DsimHandle *h, int refresh_obs, void *stream);                 /* mj_forward: sensordata (+ obs from current state) */
int dsim_zero_act(DsimHandle *h, void *stream);                                 /* new MjData on regen: act = 0 (Q3) */

/* -- step: vector_step (:259-294) = ctrl remap + mj_step x frame_skip + states + termination + reward + obs */
int dsim_step(DsimHandle *h, const void *actions_dev /* real [N][4] */, void *stream);
/* termination / reward / observation of the CURRENT state for the given raw actions; nothing is advanced, counted or
 * stored (what `terminated_fcn`, `reward_fcn`, `_get_obs` return when called on `self.states`, :275-284) */
int dsim_evaluate(DsimHandle *h, const void *actions_dev /* real [N][4] */, void *stream);
/* same with HOST buffers: the end-to-end path; returns after the outputs are in the host buffers.  Pinned buffers (all of
 * them; actions / obs 16-byte aligned): ONE launch whose bulk loads / stores move actions and obs / reward / truncated over
 * PCIe themselves (zero-copy).  Pageable buffers: staged H2D | kernel | D2H copies, pipelined over page ranges for N >= 65536. */
int dsim_step_host(DsimHandle *h, const float *actions_host /*[N][4]*/, float *obs_host /*[N][obs_dim]*/,
                   float *reward_host /*[N]*/, uint8_t *truncated_host /*[N]*/, void *stream);

/* Scheduling contract for callers that interleave several independent shards (handles) on one stream, e.g. double-buffered
 * rollouts or this repo's bench (R replicas stepped round-robin).  ready != 0 promises, for every later dsim_step(h, actions,
 * stream): the kernel queued just before it in `stream` (a) does not write `actions`, this handle's state or its setpoints,
 * and (b) is either a step kernel of a handle that has inputs_ready set too, or a kernel launched without programmatic
 * stream serialisation.  The step kernel then fetches its first state / action pages and runs the physics of its first page
 * while that predecessor is still draining, and performs the programmatic-dependency wait just before its first store; it
 * releases ITS dependents only after that wait, so an early-starting kernel knows that everything older than its
 * predecessor has completed - the state this handle wrote R launches ago included.  Default 0 (strict): nothing an earlier
 * kernel may have written is touched before the wait - what a single-shard policy -> step -> policy loop needs.
 * Results are bit-identical either way. */
int dsim_set_inputs_ready(DsimHandle *h, int ready);

/* -- reference / setpoints: self.reference (:80), control_reference (:151-172) */
int dsim_set_reference(DsimHandle *h, const double ref[4]);
/* setpoint streams of evaluation.py:135-152 evaluated on the device: env i gets the trajectory point at t + i * phase_step.
 * circle: params = {f, r, h}; step: params = {step_time}; ramp: params = {start_time, duration}; needs per_env_reference */
enum { DSIM_TRAJ_CIRCLE = 0, DSIM_TRAJ_STEP = 1, DSIM_TRAJ_RAMP = 2 };
int dsim_trajectory_reference(DsimHandle *h, int kind, double t, double phase_step, const double params[3],
                              const double start_pos[4] /* step, ramp */, const double end_pos[4], void *stream);
int dsim_control_reference(DsimHandle *h, const void *axes_dev /* real [4][ld] DENSE rows: joystick axes x,y,z,yaw after sign flips */, void *stream);

/* -- state access for callers that poke MjData (BaseDroneEnv.py:342-346) and for parity tests.
 *    Host arrays use the reference's drone-major layout: qpos [N][9|7] (ABSOLUTE positions), qvel [N][8|6], act [N][4], sensordata [N][3] */
int dsim_set_state(DsimHandle *h, const double *qpos, const double *qvel, const double *act, const int32_t *num_steps, void *stream);
int dsim_get_state(DsimHandle *h, double *qpos, double *qvel, double *act, double *sensordata, int32_t *num_steps);
int dsim_compute_states(DsimHandle *h, void *stream);                           /* get_drone_states (:357-380) -> DSIM_BUF_STATES33 */

/* -- zero-copy views for the policy.  *page_rows == 0: dense row-major, element (r, c) = ptr[r * ld + c].
 *    *page_rows == R > 0: PAGED (see the layout note above), element (row r, env i) = ptr[(i / 32) * R * 32 + r * 32 + i % 32]. */
int dsim_buffer(DsimHandle *h, int buf_id, void **dev_ptr, int64_t *rows, int64_t *cols, int64_t *ld, int32_t *dtype, int64_t *page_rows);
int dsim_stats(DsimHandle *h, double out[8], int reset);                        /* episode statistics (device sync) */
int dsim_sync(DsimHandle *h, void *stream);

/* -- caller side of the path: MyBetaDist (distributions.py:6-38) on the policy head's logits.
 *    logits [n][8] (alpha-logits then beta-logits, torch.chunk order), actions [n][4] in (0,1) (what dsim_step consumes),
 *    logp [n] or NULL.  Philox stream keyed by (seed, env_id_offset + i), counter (.., step, ..): pass a new `step` each call. */
int dsim_beta_policy(const void *logits_dev, int n, int precision, uint32_t seed, int64_t env_id_offset, uint32_t step,
                     const uint32_t *step_dev /* NULL, or a device counter ADDED to `step` (CUDA-graph replays) */,
                     int deterministic /* 1: Beta mean (deterministic_sample, :24-26) */, void *actions_dev, void *logp_dev, void *stream);

/* RMA_full inference (models/PPO/RMA/RMA_model.py:48-71,79-116, train_adaptation=False) as one fused tcgen05 kernel:
 * obs [n][22] (16 states + 6 params, LocalFrameRPYParamsEnv rows) + previous action [n][4] -> Beta-head logits [n][8] and
 * value [n].  Weights: a bf16 blob in the UMMA K-major canonical layout and an fp32 constants blob, both produced by
 * mujoco_drone_b200/policy.py::pack_rma_full from the torch module (BatchNorm folded); sizes from dsim_policy_blob_sizes. */
typedef struct DsimPolicy DsimPolicy;
int dsim_policy_blob_sizes(int64_t *weight_elems, int64_t *const_elems);
int dsim_policy_create(int device, const uint16_t *weights_host, const float *consts_host, DsimPolicy **out);
void dsim_policy_destroy(DsimPolicy *h);
int dsim_policy_forward(DsimPolicy *h, const float *obs_dev, const float *prev_action_dev,
                        const uint8_t *reset_mask_dev /* NULL, or [n]: != 0 -> the row's previous action is zero (new episode) */,
                        int n, float *logits_dev, float *value_dev, void *stream);
/* dsim_policy_forward + dsim_beta_policy in ONE launch (RMA_model.py:79-116 followed by distributions.py:6-38): the action row
 * is sampled in the kernel's logits epilogue with the same Philox streams as dsim_beta_policy, so the logits need not be
 * written at all (logits_dev / logp_dev may be NULL).  actions_dev [n][4] may alias prev_action_dev. */
int dsim_policy_forward_sample(DsimPolicy *h, const float *obs_dev, const float *prev_action_dev, const uint8_t *reset_mask_dev, int n,
                               uint32_t seed, uint32_t env_id_offset, uint32_t step, const uint32_t *step_dev, int deterministic,
                               float *logits_dev, float *value_dev, float *actions_dev, float *logp_dev, void *stream);
int dsim_policy_error(DsimPolicy *h);            /* 1: a launch hit a tensor-core barrier timeout (device sync) */

/* RMA_full inference at the REFERENCE's precision (FP32 operands and accumulation, libm tanh): one fused FP32-pipe kernel
 * (csrc/dsim_policy_fp32.cu).  The tcgen05 path above uses bf16 operands (~1e-2 on the logits against the reference's FP32
 * torch module); this one agrees with it to ~1e-6 and is the default of the rollout runner's "fused_fp32" mode.  Weights: one
 * fp32 blob (transposed per layer, BatchNorm folded) from mujoco_drone_b200/policy.py::pack_rma_full_fp32. */
typedef struct DsimPolicy32 DsimPolicy32;
int64_t dsim_policy32_blob_elems(void);
int dsim_policy32_create(int device, const float *weights_host, DsimPolicy32 **out);
void dsim_policy32_destroy(DsimPolicy32 *h);
int dsim_policy32_forward(DsimPolicy32 *h, const float *obs_dev, const float *prev_action_dev, const uint8_t *reset_mask_dev, int n,
                          float *logits_dev, float *value_dev, void *stream);

/* -- instrumentation */
int64_t dsim_debug_guard_check(DsimHandle *h);   /* handles created with DSIM_GUARD=1 in the environment: canary bytes overwritten so far (0 = none), -1 = no canaries */
int64_t dsim_launch_count(const DsimHandle *h);                                 /* kernels launched by this handle */
int dsim_debug_timeline(DsimHandle *h, uint64_t *out /*[npages][8] %globaltimer ns*/, int64_t capacity);   /* needs DSIM_TIMELINE=1 at create */
int dsim_kernel_info(int which /*0 step fp32, 1 step fp64*/, int32_t *regs, int32_t *local_bytes, int32_t *max_threads);

#ifdef __cplusplus
}
#endif
#endif
