#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE PYTHON (under stubs, see
oracle/ref_stubs.py) in the build container.  /root/reference cannot travel to the GPU box, so the frozen
vectors are committed together with this script.

    python tools/make_golden.py          # rewrites tests/golden/

What is pinned by these fixtures: transformation.py, all 17 rewards.py functions, the 14
observation_wrappers + BaseDroneEnv._get_obs, default_termination_fcn, get_drone_states layout,
sample_state / generate_drone_params distributions, and the vector_step / reset_at / regen protocol of
BaseDroneEnv.vector_step (with the physics backend = this repo's FP64 oracle, because mujoco.mj_step is
not available offline: the PHYSICS itself stays parity-unpinned).
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_stubs  # noqa: E402
from oracle import oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
ref = ref_stubs.load_reference()
T, RW, B, OW = ref.transformation, ref.rewards, ref.BaseDroneEnv, ref.observation_wrappers
rng = np.random.default_rng(20261018)
warnings.filterwarnings("ignore")


def rand_states(n, pend=True):
    """Random 33-vectors (29 without pendulum) in the reference's get_drone_states layout."""
    ref4 = np.array([0.3, -0.2, 15.0, 0.4])
    out = []
    for _ in range(n):
        pos = ref4[:3] + rng.normal(size=3) * 1.5
        rpy = np.array([rng.uniform(-np.pi, np.pi), rng.uniform(-1.5, 1.5), rng.uniform(-np.pi, np.pi)])
        vel, ang = rng.normal(size=3) * 2, rng.normal(size=3) * 3
        prp, pav = rng.normal(size=2) * 0.7, rng.normal(size=2) * 2
        acc, act = rng.normal(size=3) * 5 + [0, 0, 9.81], rng.uniform(0, 1, 4)
        params = np.array([1, .17, 7, .01, 1.2, .3]) * rng.uniform(0.8, 1.2, 6)
        parts = [pos, rpy, vel, ang] + ([prp, pav] if pend else []) + [acc, act, ref4, params]
        out.append(np.concatenate(parts))
    return np.array(out), ref4


# ---------------------------------------------------------------- A. transformation.py
n = 400
quats = rng.normal(size=(n, 4))
quats /= np.linalg.norm(quats, axis=1, keepdims=True)
quats[::7] *= rng.uniform(0.5, 2.0, size=(len(quats[::7]), 1))          # un-normalised inputs are legal
special = []
for pitch in (np.pi / 2, -np.pi / 2, np.pi / 2 - 1e-9, -np.pi / 2 + 1e-9, np.pi / 2 - 1e-5, 0.0):
    for yaw, roll in ((0.3, -0.8), (2.9, 3.0), (-3.1, 0.1)):
        special.append(T.mujoco_rpy2quat(np.array([roll, pitch, yaw])))
special += [np.array([1., 0, 0, 0]), np.array([0., 1, 0, 0]), np.array([0., 0, 1, 0]), np.array([0., 0, 0, 1]), np.array([-1., 0, 0, 0])]
quats = np.concatenate([quats, np.array(special)])
rpy_in = np.column_stack([rng.uniform(-np.pi, np.pi, n), rng.uniform(-np.pi / 2, np.pi / 2, n), rng.uniform(-np.pi, np.pi, n)])
prp_in = rng.normal(size=(n, 2))
np.savez_compressed(
    os.path.join(OUT, "transform.npz"),
    quats=quats, n_special=len(special),
    quat2rpy=np.array([T.mujoco_quat2rpy(q) for q in quats]),
    quat2dcm=np.array([T.mujoco_quat2DCM(q) for q in quats]),
    rpy_in=rpy_in, rpy2quat=np.array([T.mujoco_rpy2quat(r) for r in rpy_in]),
    prp_in=prp_in, pendulumrp2quat=np.array([T.mujoco_pendulumrp2quat(p) for p in prp_in]),
)

# ---------------------------------------------------------------- B. rewards.py
states, ref4 = rand_states(300)
states[:40, :3] = ref4[:3] + rng.normal(size=(40, 3)) * 0.08            # exercise the `pos_err < 0.15` / close_enough branches
states[40:60, :3] = ref4[:3] + rng.normal(size=(20, 3)) * 3.5           # too_far branches
actions = rng.uniform(0, 1, size=(300, 4))
num_steps = rng.integers(0, 1100, size=300)
envns = type("E", (), {})()
envns.reference = ref4
envns.max_distance = 4
rew_out = {}
for name in O.REWARD_IDS:
    f = getattr(RW, name)
    rew_out[name] = np.array([float(f(envns, states[i], actions[i], int(num_steps[i]))) for i in range(300)])
np.savez_compressed(os.path.join(OUT, "rewards.npz"), states=states, actions=actions, num_steps=num_steps,
                    reference=ref4, max_distance=4.0, **{"out_" + k: v for k, v in rew_out.items()})

# ---------------------------------------------------------------- C. observation wrappers
states, ref4 = rand_states(200)
obs_out = {}
for name in O.OBS_IDS:
    cls = B.BaseDroneEnv if name == "BaseDroneEnv" else getattr(OW, name)
    e = object.__new__(cls)
    e.states = list(states)
    e.reference = ref4
    try:
        obs_out[name] = np.array(e._get_obs())
    except NameError:
        obs_out[name] = np.zeros((0,))                                  # observation_wrappers.py:448 raises
np.savez_compressed(os.path.join(OUT, "obs.npz"), states=states, reference=ref4,
                    **{"out_" + k: v for k, v in obs_out.items()})

# ---------------------------------------------------------------- D. termination
states, ref4 = rand_states(400)
states[:100, :3] = ref4[:3] + rng.normal(size=(100, 3)) * 3.0
ns = rng.integers(0, 600, size=400)
ns[::9] = 512
ns[1::9] = 511
envns.reference, envns.max_distance, envns.max_steps = ref4, 4, 512
term = np.array([bool(B.default_termination_fcn(envns, states[i], None, int(ns[i]))) for i in range(400)])
np.savez_compressed(os.path.join(OUT, "termination.npz"), states=states, num_steps=ns, reference=ref4,
                    max_distance=4.0, max_steps=512, out=term)

# ---------------------------------------------------------------- E. get_drone_states layout (pendulum on/off)
gs = {}
for pend in (True, False):
    cfg = dict(B.base_config, num_drones=5, pendulum=pend, reference=[0.5, -1, 14, 0.2])
    e = ref_stubs.make_env(B.BaseDroneEnv, cfg)
    p = 2 * int(pend)
    e.data.qpos[:] = rng.normal(size=e.data.qpos.size)
    e.data.qvel[:] = rng.normal(size=e.data.qvel.size)
    e.data.act[:] = rng.uniform(size=e.data.act.size)
    e.data.sensordata[:] = rng.normal(size=e.data.sensordata.size)
    key = "pend" if pend else "nopend"
    gs.update({key + "_qpos": e.data.qpos.copy(), key + "_qvel": e.data.qvel.copy(), key + "_act": e.data.act.copy(),
               key + "_sens": e.data.sensordata.copy(), key + "_ref": np.array(e.reference, dtype=float),
               key + "_params": np.array([list(d.values()) for d in e.drone_params]),
               key + "_states": np.array(e.get_drone_states())})
np.savez_compressed(os.path.join(OUT, "drone_states.npz"), **gs)

# ---------------------------------------------------------------- F. reset / domain-randomisation distributions
cfg = dict(B.base_config, num_drones=4000, angle_variance=[0.3, 0.2])
e = ref_stubs.make_env(B.BaseDroneEnv, cfg)
params = np.array([list(d.values()) for d in e.drone_params])
samples = [e.sample_state() for _ in range(4000)]
np.savez_compressed(os.path.join(OUT, "sampling.npz"),
                    qpos=np.array([s[0] for s in samples]), qvel=np.array([s[1] for s in samples]), params=params,
                    start_pos=np.array(cfg['start_pos'], dtype=float), state_difficulty=cfg['state_difficulty'],
                    param_difficulty=cfg['param_difficulty'], max_random_offset=cfg['max_random_offset'],
                    angle_variance=np.array(cfg['angle_variance']), vel_variance=np.array(cfg['vel_variance'], dtype=float),
                    ang_vel_variance=np.array(cfg['ang_vel_variance'], dtype=float),
                    pendulum_rp_variance=np.array(cfg['pendulum_rp_variance']),
                    pendulum_ang_vel_variance=np.array(cfg['pendulum_ang_vel_variance']))

# ---------------------------------------------------------------- G. vector_step / reset_at / regen protocol
# Physics backend = the FP64 oracle (mujoco.mj_step is unavailable): what this pins is the reference's
# PYTHON orchestration: ctrl remap, counters, truncation, reward/obs call order, stale reset_at obs (Q1),
# act persistence across resets (Q3), regen semantics (Q8).


def oracle_physics(env, ctrl, n_frames):
    p = 2 * int(bool(env.pendulum))
    nq, nv = 7 + p, 6 + p
    for i in range(env.num_drones):
        m = O.compile_model(list(env.drone_params[i].values()), env.pendulum, env.frequency, True)
        qp, qv, act, sens = O.step(m, env.data.qpos[nq * i:nq * (i + 1)], env.data.qvel[nv * i:nv * (i + 1)],
                                   env.data.act[4 * i:4 * i + 4], ctrl[4 * i:4 * i + 4], n_frames)
        env.data.qpos[nq * i:nq * (i + 1)], env.data.qvel[nv * i:nv * (i + 1)] = qp, qv
        env.data.act[4 * i:4 * i + 4], env.data.sensordata[3 * i:3 * i + 3] = act, sens


def oracle_forward(env):
    p = 2 * int(bool(env.pendulum))
    nq, nv = 7 + p, 6 + p
    for i in range(env.num_drones):
        m = O.compile_model(list(env.drone_params[i].values()), env.pendulum, env.frequency, True)
        f = O.forward(m, env.data.qpos[nq * i:nq * (i + 1)], env.data.qvel[nv * i:nv * (i + 1)],
                      env.data.act[4 * i:4 * i + 4], env.data.ctrl[4 * i:4 * i + 4])
        env.data.sensordata[3 * i:3 * i + 3] = f['sensordata']


def regen_init(self, model, frame_skip, **kw):
    """stand-in for extendedEnv.__init__ on regen: a NEW MjData -> qpos0, zero qvel/act/ctrl/sensordata"""
    self.data.qpos[:] = self.init_qpos
    self.data.qvel[:] = 0
    self.data.act[:] = 0
    self.data.ctrl[:] = 0
    self.data.sensordata[:] = 0


B.mjcf_to_mjmodel = lambda x: None
B.make_sim = lambda *a, **k: None
B.extendedEnv.__init__ = regen_init

N, STEPS = 6, 40
cfg = dict(B.base_config, num_drones=N, max_steps=12, regen_env_at_steps=25, reward_fcn=RW.distance_energy_reward,
           max_distance=1.2, state_difficulty=0.4, param_difficulty=1.0, skip_steps=2)
e = ref_stubs.make_env(OW.LocalFrameRPYParamsEnv, cfg, physics=oracle_physics)
e._forward = oracle_forward
log = {k: [] for k in ("actions", "obs", "rewards", "truncated", "num_steps", "total_steps", "qpos_after", "qvel_after",
                       "act_after", "sens_after", "params", "reset_obs", "reset_idx", "qpos_pre", "qvel_pre", "act_pre")}
obs0, _ = e.vector_reset()
log0 = dict(obs0=np.array(obs0), qpos0=e.data.qpos.copy(), qvel0=e.data.qvel.copy(), act0=e.data.act.copy(),
            sens0=e.data.sensordata.copy(), params0=np.array([list(d.values()) for d in e.drone_params]))
for t in range(STEPS):
    a = rng.uniform(0, 1, size=(N, 4))
    a[:, :] = 0.45 + 0.1 * a if t % 3 else a
    log["qpos_pre"].append(e.data.qpos.copy()); log["qvel_pre"].append(e.data.qvel.copy()); log["act_pre"].append(e.data.act.copy())
    obs, rew, dones, trunc, infos = e.vector_step(list(a))
    assert not any(dones)
    log["actions"].append(a); log["obs"].append(np.array(obs)); log["rewards"].append(np.array(rew, dtype=float))
    log["truncated"].append(np.array(trunc, dtype=bool)); log["num_steps"].append(e.num_steps.copy())
    log["total_steps"].append(e.total_steps); log["qpos_after"].append(e.data.qpos.copy()); log["qvel_after"].append(e.data.qvel.copy())
    log["act_after"].append(e.data.act.copy()); log["sens_after"].append(e.data.sensordata.copy())
    log["params"].append(np.array([list(d.values()) for d in e.drone_params]))
    ro, ri = np.full((N, 22), np.nan), np.zeros(N, dtype=bool)
    for i in range(N):                      # RLlib protocol (Q16): reset_at for every truncated sub-env, ascending
        if trunc[i]:
            ob, _ = e.reset_at(i)
            ro[i], ri[i] = ob, True
    log["reset_obs"].append(ro); log["reset_idx"].append(ri)
np.savez_compressed(os.path.join(OUT, "protocol.npz"), frame_skip=2, max_steps=12, regen=25, max_distance=1.2,
                    frequency=cfg['frequency'], reference=np.array(cfg['reference'], dtype=float),
                    **log0, **{k: np.array(v) for k, v in log.items()})
# error behaviour: wrong action length (mujoco_env_custom.py:200-201)
try:
    e.vector_step(list(np.zeros((N - 1, 4))))
    raise SystemExit("expected ValueError")
except ValueError as ex:
    assert str(ex) == "Action dimension mismatch"

# ---------------------------------------------------------------- trajectory generators (evaluation.py:135-152)
# evaluation.py imports ray / pickle-loads checkpoints at module level: execute only the three generator definitions
import ast
_src = open("/root/reference/evaluation.py").read()
_ns = {"np": np}
for node in ast.parse(_src).body:
    if isinstance(node, ast.FunctionDef) and node.name in ("gen_circle_trajectory", "gen_step_trajectory", "gen_ramp_trajectory"):
        exec(compile(ast.Module([node], []), "evaluation.py", "exec"), _ns)
tc, circ = _ns["gen_circle_trajectory"](T=3, f=0.7, r=1.5, h=14.5)
ts, stp = _ns["gen_step_trajectory"](step_time=1.2, duration=3, start_pos=[0, 0, 15, 0], end_pos=[1, -1, 16, 0.5])
tr, rmp = _ns["gen_ramp_trajectory"](start_time=0.8, duration=3, start_pos=[0, 0, 15, 0], end_pos=[1, -1, 16, 0.5])
np.savez_compressed(os.path.join(OUT, "trajectories.npz"), t_circle=tc, circle=circ, t_step=ts, step=stp, t_ramp=tr, ramp=rmp)

print("golden fixtures written to", OUT)
for f in sorted(os.listdir(OUT)):
    print(f"  {f:24s} {os.path.getsize(os.path.join(OUT, f)) / 1024:8.1f} KiB")
