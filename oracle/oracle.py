"""ctypes binding of the CPU FP64 oracle (oracle/dsim_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
Physics parity is UNPINNED (no MuJoCo available offline); obs/reward/termination/rotations are
pinned against the reference's own Python through tests/golden/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdsim_oracle.so")

MAXBODY, MAXNV = 5, 8


MAXGEOM, MAXCON = 21, 84


class OrcGeom(C.Structure):
    _fields_ = [("body", C.c_int), ("type", C.c_int), ("size", C.c_double * 3), ("pos", C.c_double * 3), ("yaw", C.c_double)]


class OrcContact(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("dist", C.c_double), ("body", C.c_int), ("geom", C.c_int)]


class OrcModel(C.Structure):
    _fields_ = [
        ("nbody", C.c_int), ("nv", C.c_int), ("nq", C.c_int), ("pendulum", C.c_int),
        ("parent", C.c_int * MAXBODY),
        ("pos", (C.c_double * 3) * MAXBODY),
        ("quat", (C.c_double * 4) * MAXBODY),
        ("mass", C.c_double * MAXBODY),
        ("ipos", (C.c_double * 3) * MAXBODY),
        ("iquat", (C.c_double * 4) * MAXBODY),
        ("inertia", (C.c_double * 3) * MAXBODY),
        ("jnt_type", C.c_int * MAXBODY),
        ("jnt_axis", (C.c_double * 3) * MAXBODY),
        ("dofadr", C.c_int * MAXBODY),
        ("qposadr", C.c_int * MAXBODY),
        ("dof_body", C.c_int * MAXNV),
        ("damping", C.c_double * MAXNV),
        ("site_pos", (C.c_double * 3) * 5),
        ("gear", (C.c_double * 6) * 4),
        ("tau", C.c_double * 4),
        ("timestep", C.c_double), ("density", C.c_double), ("viscosity", C.c_double),
        ("gravity", C.c_double * 3),
        ("params", C.c_double * 6),
        ("ground", C.c_int), ("ngeom", C.c_int),
        ("geom", OrcGeom * MAXGEOM),
        ("invweight0", (C.c_double * 2) * MAXBODY),
    ]


class OrcResetCfg(C.Structure):
    _fields_ = [
        ("start_pos", C.c_double * 4),
        ("max_pos_offset", C.c_double),
        ("angle_sigma", C.c_double * 2),
        ("vel_sigma", C.c_double * 3),
        ("ang_vel_sigma", C.c_double * 3),
        ("pend_rp_sigma", C.c_double * 2),
        ("pend_vel_sigma", C.c_double * 2),
        ("random_start_pos", C.c_int),
        ("pendulum", C.c_int),
        ("param_center", C.c_double * 6),
        ("param_halfwidth", C.c_double * 6),
        ("param_difficulty", C.c_double),
        ("random_params", C.c_int),
    ]


REWARD_IDS = {
    "default_reward_fcn": 0, "distance_reward_fcn": 1, "distance_energy_reward": 2,
    "distance_energy_reward_pendulum_angle": 3, "distance_energy_reward_pendulum_angle2": 4,
    "distance_energy_reward_pendulum_angle3": 5, "distance_energy_reward_pendulum_en": 6,
    "distance_energy_reward_pendulum_en2": 7, "distance_energy_reward_pendulum_en3": 8,
    "distance_energy_reward_pendulum_en4": 9, "distance_time_energy_reward": 10, "reward_1": 11,
    "reward_pendulum_dist": 12, "reward_pendulumDistHeading": 13, "reward_2": 14,
    "reward_2_penergy": 15, "reward_3": 16,
}
OBS_IDS = {
    "BaseDroneEnv": 0, "GlobalFrameRPYEnv": 1, "LocalFramePRYEnv": 2, "LocalFrameFullStateEnv": 3,
    "LocalFrameFullStateZvecEnv": 4, "LocalFramePRYaccEnv": 5, "LocalFramePRYParamsEnv": 6,
    "LocalFramePRYaccParamsEnv": 7, "LocalFrameRPYParamsEnv": 8, "LocalFrameRPYFakeParamsEnv": 9,
    "LocalFrameRPYEnv": 10, "LocalFramePRYaccNoPendEnv": 11, "LocalFramePRYaccParamsNoPendEnv": 12,
    "LocalFrameRmParamsEnv": 13, "LocalFrameZvecEnv": 14,
}
OBS_DIMS = {0: 33, 1: 16, 2: 16, 3: 23, 4: 24, 5: 19, 6: 22, 7: 25, 8: 22, 9: 22, 10: 16, 11: 15,
            12: 21, 13: 28, 14: 17}


def build(force=False):
    """Compile oracle/dsim_oracle.c with the Makefile (gcc, OpenMP)."""
    src = os.path.join(_HERE, "dsim_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp = C.POINTER(C.c_double)
        L.orc_round_prec5.restype = C.c_double
        L.orc_round_prec5.argtypes = [C.c_double]
        L.orc_compile.argtypes = [dp, C.c_int, C.c_double, C.c_int, C.POINTER(OrcModel)]
        L.orc_forward.argtypes = [C.POINTER(OrcModel)] + [dp] * 9
        L.orc_step.argtypes = [C.POINTER(OrcModel), dp, dp, dp, dp, dp, C.c_int]
        L.orc_collide.argtypes = [C.POINTER(OrcModel), dp, C.POINTER(OrcContact)]
        L.orc_collide.restype = C.c_int
        L.orc_forward_contact.argtypes = [C.POINTER(OrcModel)] + [dp] * 7
        L.orc_forward_contact.restype = C.c_int
        L.orc_energy.argtypes = [C.POINTER(OrcModel), dp, dp, dp, dp]
        for f in (L.orc_quat2rpy, L.orc_rpy2quat, L.orc_quat2dcm, L.orc_pendulumrp2quat):
            f.argtypes = [dp, dp]
        L.orc_drone_state.restype = C.c_int
        L.orc_drone_state.argtypes = [C.POINTER(OrcModel), dp, dp, dp, dp, dp, dp]
        L.orc_termination.restype = C.c_int
        L.orc_termination.argtypes = [dp, dp, C.c_double, C.c_int64, C.c_int64]
        L.orc_reward.restype = C.c_double
        L.orc_reward.argtypes = [C.c_int, dp, C.c_int, dp, C.c_int64, dp, C.c_double]
        L.orc_obs.restype = C.c_int
        L.orc_obs.argtypes = [C.c_int, dp, C.c_int, dp, dp]
        u32p = C.POINTER(C.c_uint32)
        L.orc_philox4x32.argtypes = [u32p, u32p, u32p]
        L.orc_sample_state.argtypes = [C.POINTER(OrcResetCfg), C.c_uint32, C.c_uint32, C.c_uint32, dp, dp]
        L.orc_sample_params.argtypes = [C.POINTER(OrcResetCfg), C.c_uint32, C.c_uint32, C.c_uint32, dp]
        L.orc_beta_policy.argtypes = [dp, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, dp, dp]
        L.orc_beta_policy.restype = None
        L.orc_vector_step.argtypes = [C.c_int, C.POINTER(OrcModel), C.c_int, dp, dp, dp, dp,
                                      C.POINTER(C.c_int64), dp, dp, C.c_int, C.c_int, C.c_int,
                                      C.c_double, C.c_int64, dp, C.c_int, dp, C.POINTER(C.c_uint8), C.c_int]
        L.orc_max_threads.restype = C.c_int
        L.orc_vector_step_repeat.argtypes = [C.c_int, C.c_int, C.POINTER(OrcModel), C.c_int, dp, dp, dp, dp, C.POINTER(C.c_int64), dp, C.c_int, dp,
                                             C.c_int, C.c_int, C.c_int, C.c_double, C.c_int64, dp, C.c_int, dp, C.POINTER(C.c_uint8), C.c_int,
                                             C.POINTER(OrcResetCfg), C.c_uint32, C.c_uint32, u32p, C.POINTER(C.c_int64)]
        L.orc_vector_step_repeat.restype = None
        L.orc_step_batch.argtypes = [C.c_int, C.POINTER(OrcModel), dp, dp, dp, dp, dp, C.c_int]
        L.orc_step_batch.restype = None
        L.orc_control_reference.argtypes = [dp, dp, dp]
        L.orc_control_reference.restype = None
        L.orc_reset_truncated.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(OrcResetCfg), C.c_uint32, C.c_uint32, u32p,
                                          C.POINTER(C.c_uint8), dp, dp, C.POINTER(C.c_int64)]
        L.orc_reset_truncated.restype = None
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _arr(x, n=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    if n is not None:
        assert a.size == n, (a.size, n)
    return a


def round_prec5(x):
    return lib().orc_round_prec5(float(x))


def compile_model(params, pendulum=True, frequency=100.0, round_precision=True, ground=False):
    """params: mass, arm_len, motor_force, motor_tau, pendulum_len, weight_mass.  ground: simulate floor contacts."""
    m = OrcModel()
    p = _arr(params, 6)
    lib().orc_compile(_dp(p), int(pendulum), float(frequency), int(round_precision), C.byref(m))
    m.ground = int(ground)
    return m


def collide(m, qpos):
    """Floor contacts of configuration qpos: list of dicts (pos, dist, body, geom)."""
    con = (OrcContact * MAXCON)()
    n = lib().orc_collide(C.byref(m), _dp(_arr(qpos)), con)
    return [dict(pos=np.array(con[i].pos[:]), dist=con[i].dist, body=con[i].body, geom=con[i].geom) for i in range(n)]


def forward_contact(m, qpos, qvel, act, ctrl):
    qpos, qvel, act, ctrl = _arr(qpos), _arr(qvel), _arr(act, 4), _arr(ctrl, 4)
    qacc, qc, sens = np.zeros(m.nv), np.zeros(m.nv), np.zeros(3)
    n = lib().orc_forward_contact(C.byref(m), _dp(qpos), _dp(qvel), _dp(act), _dp(ctrl), _dp(qacc), _dp(qc), _dp(sens))
    return dict(qacc=qacc, qfrc_constraint=qc, sensordata=sens, ncon=n)


def forward(m, qpos, qvel, act, ctrl):
    qpos, qvel, act, ctrl = _arr(qpos), _arr(qvel), _arr(act, 4), _arr(ctrl, 4)
    nv = m.nv
    qacc, act_dot, sens = np.zeros(nv), np.zeros(4), np.zeros(3)
    M, qs = np.zeros((nv, nv)), np.zeros(nv)
    lib().orc_forward(C.byref(m), _dp(qpos), _dp(qvel), _dp(act), _dp(ctrl), _dp(qacc), _dp(act_dot),
                      _dp(sens), _dp(M), _dp(qs))
    return dict(qacc=qacc, act_dot=act_dot, sensordata=sens, M=M, qfrc_smooth=qs)


def step(m, qpos, qvel, act, ctrl, nstep=1):
    """Returns new (qpos, qvel, act, sensordata); inputs are not modified."""
    qpos, qvel, act, ctrl = _arr(qpos).copy(), _arr(qvel).copy(), _arr(act, 4).copy(), _arr(ctrl, 4)
    sens = np.zeros(3)
    lib().orc_step(C.byref(m), _dp(qpos), _dp(qvel), _dp(act), _dp(ctrl), _dp(sens), int(nstep))
    return qpos, qvel, act, sens


def energy(m, qpos, qvel):
    ke, pe = C.c_double(), C.c_double()
    qpos, qvel = _arr(qpos), _arr(qvel)
    lib().orc_energy(C.byref(m), _dp(qpos), _dp(qvel), C.byref(ke), C.byref(pe))
    return ke.value, pe.value


def quat2rpy(q):
    q, o = _arr(q, 4), np.zeros(3)
    lib().orc_quat2rpy(_dp(q), _dp(o))
    return o


def rpy2quat(r):
    r, o = _arr(r, 3), np.zeros(4)
    lib().orc_rpy2quat(_dp(r), _dp(o))
    return o


def quat2dcm(q):
    q, o = _arr(q, 4), np.zeros(9)
    lib().orc_quat2dcm(_dp(q), _dp(o))
    return o.reshape(3, 3)


def pendulumrp2quat(rp):
    rp, o = _arr(rp, 2), np.zeros(4)
    lib().orc_pendulumrp2quat(_dp(rp), _dp(o))
    return o


def drone_state(m, qpos, qvel, act, sens, reference):
    s = np.zeros(40)
    n = lib().orc_drone_state(C.byref(m), _dp(_arr(qpos)), _dp(_arr(qvel)), _dp(_arr(act, 4)),
                              _dp(_arr(sens, 3)), _dp(_arr(reference, 4)), _dp(s))
    return s[:n].copy()


def termination(state, reference, max_distance, num_steps, max_steps):
    return bool(lib().orc_termination(_dp(_arr(state)), _dp(_arr(reference, 4)), float(max_distance),
                                      int(num_steps), int(max_steps)))


def reward(reward_id, state, action, num_steps, reference, max_distance):
    s = _arr(state)
    return lib().orc_reward(int(reward_id), _dp(s), s.size, _dp(_arr(action, 4)), int(num_steps),
                            _dp(_arr(reference, 4)), float(max_distance))


def obs(obs_id, state, reference):
    s, o = _arr(state), np.zeros(40)
    n = lib().orc_obs(int(obs_id), _dp(s), s.size, _dp(_arr(reference, 4)), _dp(o))
    if n < 0:
        raise NameError("observation variant raises in the reference (observation_wrappers.py:448)")
    return o[:n].copy()


def philox4x32(ctr, key):
    c = (C.c_uint32 * 4)(*[int(x) for x in ctr])
    k = (C.c_uint32 * 2)(*[int(x) for x in key])
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32(c, k, o)
    return np.array(list(o), dtype=np.uint32)


def make_reset_cfg(start_pos, max_pos_offset, angle_sigma, vel_sigma, ang_vel_sigma, pend_rp_sigma,
                   pend_vel_sigma, random_start_pos, pendulum, param_center=None, param_halfwidth=None,
                   param_difficulty=0.0, random_params=False):
    c = OrcResetCfg()
    c.start_pos[:] = [float(x) for x in start_pos]
    c.max_pos_offset = float(max_pos_offset)
    c.angle_sigma[:] = [float(x) for x in angle_sigma]
    c.vel_sigma[:] = [float(x) for x in vel_sigma]
    c.ang_vel_sigma[:] = [float(x) for x in ang_vel_sigma]
    c.pend_rp_sigma[:] = [float(x) for x in pend_rp_sigma]
    c.pend_vel_sigma[:] = [float(x) for x in pend_vel_sigma]
    c.random_start_pos = int(random_start_pos)
    c.pendulum = int(pendulum)
    c.param_center[:] = [float(x) for x in (param_center if param_center is not None else [0] * 6)]
    c.param_halfwidth[:] = [float(x) for x in (param_halfwidth if param_halfwidth is not None else [0] * 6)]
    c.param_difficulty = float(param_difficulty)
    c.random_params = int(random_params)
    return c


def sample_state(cfg, seed, env_id, reset_count):
    qpos, qvel = np.zeros(9), np.zeros(8)
    lib().orc_sample_state(C.byref(cfg), int(seed), int(env_id), int(reset_count), _dp(qpos), _dp(qvel))
    n = 2 if cfg.pendulum else 0
    return qpos[:7 + n].copy(), qvel[:6 + n].copy()


def sample_params(cfg, seed, env_id, regen_epoch):
    p = np.zeros(6)
    lib().orc_sample_params(C.byref(cfg), int(seed), int(env_id), int(regen_epoch), _dp(p))
    return p


class CpuVecEnv:
    """Batched CPU vec-env on the oracle (OpenMP over envs) — the timed CPU baseline and the parity
    checker for whole vector_step calls.  State arrays use the reference's drone-major layout."""

    def __init__(self, params, pendulum=True, frequency=100.0, frame_skip=1, round_precision=True, ground=False):
        params = np.atleast_2d(np.asarray(params, dtype=np.float64))
        self.n = params.shape[0]
        self.models = (OrcModel * self.n)()
        L = lib()
        for i in range(self.n):
            p = np.ascontiguousarray(params[i])
            L.orc_compile(_dp(p), int(pendulum), float(frequency), int(round_precision), C.byref(self.models[i]))
            self.models[i].ground = int(ground)
        self.nq, self.nv = self.models[0].nq, self.models[0].nv
        self.frame_skip = frame_skip
        self.qpos = np.zeros((self.n, self.nq))
        self.qpos[:, 3] = 1.0
        self.qvel = np.zeros((self.n, self.nv))
        self.act = np.zeros((self.n, 4))
        self.sens = np.zeros((self.n, 3))
        self.num_steps = np.zeros(self.n, dtype=np.int64)

    def step(self, actions, reference, reward_id, obs_id, max_distance, max_steps, nthreads=0, want_obs=True):
        actions = np.ascontiguousarray(np.asarray(actions, dtype=np.float64).reshape(self.n, 4))
        reference = np.ascontiguousarray(np.asarray(reference, dtype=np.float64))
        per_env = int(reference.ndim == 2)
        od = OBS_DIMS[obs_id] if (obs_id != 0 or self.nq == 9) else 29
        obs = np.zeros((self.n, od)) if want_obs else None
        rew = np.zeros(self.n)
        trunc = np.zeros(self.n, dtype=np.uint8)
        lib().orc_vector_step(self.n, self.models, self.frame_skip, _dp(self.qpos), _dp(self.qvel), _dp(self.act),
                              _dp(self.sens), self.num_steps.ctypes.data_as(C.POINTER(C.c_int64)), _dp(actions),
                              _dp(reference), per_env, int(reward_id), int(obs_id), float(max_distance),
                              int(max_steps), _dp(obs), od, _dp(rew), trunc.ctypes.data_as(C.POINTER(C.c_uint8)),
                              int(nthreads))
        return obs, rew, trunc.astype(bool)

    def step_repeat(self, reps, action_bank, reference, reward_id, obs_id, max_distance, max_steps, cfg, seed, env0, nthreads=0):
        """`reps` x (vector_step + reset_at of the truncated envs) inside ONE C call; returns the number of truncations"""
        if not hasattr(self, "reset_count"):
            self.reset_count = np.zeros(self.n, dtype=np.uint32)
        bank = np.ascontiguousarray(np.asarray(action_bank, dtype=np.float64).reshape(-1, self.n, 4))
        reference = np.ascontiguousarray(np.asarray(reference, dtype=np.float64))
        od = OBS_DIMS[obs_id] if (obs_id != 0 or self.nq == 9) else 29
        obs, rew, trunc, cnt = np.zeros((self.n, od)), np.zeros(self.n), np.zeros(self.n, dtype=np.uint8), np.zeros(1, dtype=np.int64)
        lib().orc_vector_step_repeat(int(reps), self.n, self.models, self.frame_skip, _dp(self.qpos), _dp(self.qvel), _dp(self.act), _dp(self.sens),
                                     self.num_steps.ctypes.data_as(C.POINTER(C.c_int64)), _dp(bank), bank.shape[0], _dp(reference),
                                     int(reference.ndim == 2), int(reward_id), int(obs_id), float(max_distance), int(max_steps), _dp(obs), od,
                                     _dp(rew), trunc.ctypes.data_as(C.POINTER(C.c_uint8)), int(nthreads), C.byref(cfg), int(seed) & 0xFFFFFFFF,
                                     int(env0) & 0xFFFFFFFF, self.reset_count.ctypes.data_as(C.POINTER(C.c_uint32)),
                                     cnt.ctypes.data_as(C.POINTER(C.c_int64)))
        return int(cnt[0])

    def reset_truncated(self, cfg, seed, env0, truncated):
        """RLlib's reset_at() for every truncated env (BaseDroneEnv.py:334-351), reset stream per env like the CUDA path."""
        if not hasattr(self, "reset_count"):
            self.reset_count = np.zeros(self.n, dtype=np.uint32)
        t = np.ascontiguousarray(truncated, dtype=np.uint8)
        lib().orc_reset_truncated(self.n, self.nq, self.nv, C.byref(cfg), int(seed) & 0xFFFFFFFF, int(env0) & 0xFFFFFFFF,
                                  self.reset_count.ctypes.data_as(C.POINTER(C.c_uint32)), t.ctypes.data_as(C.POINTER(C.c_uint8)),
                                  _dp(self.qpos), _dp(self.qvel), self.num_steps.ctypes.data_as(C.POINTER(C.c_int64)))


def control_reference(reference, axes, start_pos):
    """BaseDroneEnv.control_reference (:151-172) for one joystick sample `axes` = (x, y, z, yaw) after the sign flips."""
    r = np.array(reference, dtype=np.float64)
    lib().orc_control_reference(_dp(r), _dp(np.ascontiguousarray(axes, dtype=np.float64)), _dp(np.ascontiguousarray(start_pos, dtype=np.float64)))
    return r


def max_threads():
    return lib().orc_max_threads()


def beta_policy(logits, seed, env0, step, deterministic=False):
    """MyBetaDist (distributions.py:6-38) on [n, 2A] logits -> (actions [n, A], logp [n]); same Philox stream as the kernel."""
    x = np.ascontiguousarray(logits, dtype=np.float64)
    n, a2 = x.shape
    act, lp = np.zeros((n, a2 // 2)), np.zeros(n)
    lib().orc_beta_policy(_dp(x), int(n), int(a2 // 2), int(seed) & 0xFFFFFFFF, int(env0) & 0xFFFFFFFF, int(step), int(bool(deterministic)),
                          _dp(act), _dp(lp))
    return act, lp
