#!/usr/bin/env python
"""Measured FP32-vs-FP64 deviation of the step kernel, backing every tolerance stated in tests/test_gpu_parity.py.

One vector_step (frame_skip 1 and 2) of the FP32 product kernel on N random states / actions / parameter sets against the
FP64 oracle started from exactly the state the device holds; then obs / reward on identical states (dsim_evaluate), then a
100-step open-loop trajectory.  Prints percentiles of the absolute (or, where stated, relative-to-(1+|x|)) deviation per
quantity next to the tolerance the tests use.      python tools/fp32_error_hist.py > profiles/r02_fp32_error_histogram.txt
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mujoco_drone_b200 as M
from oracle import oracle as O

NOMINAL = np.array([1, 0.17, 7, 0.01, 1.2, 0.3])
PK = ("mass", "arm_len", "motor_force", "motor_tau", "pendulum_len", "weight_mass")
rng = np.random.default_rng(123)
N = 4096


def pct(name, err, tol, unit=""):
    q = np.percentile(err, [50, 90, 99, 99.9, 100])
    print(f"{name:44s} p50 {q[0]:.2e}  p90 {q[1]:.2e}  p99 {q[2]:.2e}  p99.9 {q[3]:.2e}  max {q[4]:.2e}   tolerance {tol:.0e}{unit}   max/tol {q[4] / tol:.2f}")


def rand_inputs(n, scale=1.0):
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    qpos = np.concatenate([np.array([0, 0, 15.0]) + rng.normal(size=(n, 3)), q, rng.normal(size=(n, 2)) * 0.6], axis=1)
    return qpos, rng.normal(size=(n, 8)) * scale, rng.uniform(0, 1, size=(n, 4)), rng.uniform(0, 1, size=(n, 4)), NOMINAL * rng.uniform(0.85, 1.15, size=(n, 6))


print(f"# FP32 step kernel vs FP64 oracle, {N} random envs per row (tools/fp32_error_hist.py), GPU: {torch.cuda.get_device_name(0)}")
for fs in (1, 2):
    qpos, qvel, act, actions, params = rand_inputs(N)
    env = M.BaseDroneEnv(dict(M.base_config, num_drones=N, skip_steps=fs, random_params=False, max_distance=100))
    env.drone_params = [dict(zip(PK, p)) for p in params]
    env.set_state(qpos, qvel, act)
    q0, v0, a0, _, _ = env.get_state()
    env.step_tensor(torch.as_tensor(actions, device="cuda"))
    q1, v1, a1, s1, _ = env.get_state()
    prm = env.drone_params
    a_in = actions.astype(np.float32).astype(np.float64)
    E = {k: [] for k in ("pos", "quat", "hinge", "vel", "omega", "hrate", "sum_rate", "acc", "act")}
    for i in range(N):
        m = O.compile_model(np.array(list(prm[i].values())), True, 100, True)
        oq, ov, oa, osn = O.step(m, q0[i], v0[i], a0[i], 0.1 + 0.9 * a_in[i], fs)
        E["pos"].append(np.abs(q1[i, :3] - oq[:3]).max()); E["quat"].append(np.abs(q1[i, 3:7] - oq[3:7]).max()); E["hinge"].append(np.abs(q1[i, 7:] - oq[7:]).max())
        E["vel"].append((np.abs(v1[i, :3] - ov[:3]) / (1 + np.abs(ov[:3]))).max())
        E["omega"].append((np.abs(v1[i, 3:6] - ov[3:6]) / (1 + np.abs(ov[3:6]))).max())
        E["hrate"].append((np.abs(v1[i, 6:] - ov[6:]) / (1 + np.abs(ov[6:]))).max())
        E["sum_rate"].append(abs((v1[i, 3] + v1[i, 6]) - (ov[3] + ov[6])) / (1 + abs(ov[3]) + abs(ov[6])))
        E["acc"].append((np.abs(s1[i] - osn) / (1 + np.abs(osn))).max()); E["act"].append(np.abs(a1[i] - oa).max())
    print(f"## one vector_step, frame_skip = {fs} (tolerances scale with frame_skip in the tests)")
    pct("|d pos| [m]", E["pos"], 1e-6 * fs); pct("|d quat|", E["quat"], 1e-6 * fs); pct("|d hinge angle| [rad]", E["hinge"], 1e-6 * fs)
    pct("|d v_lin| / (1+|v|)", E["vel"], 1e-5 * fs); pct("|d omega| / (1+|w|)", E["omega"], 1e-4 * fs); pct("|d hinge rate| / (1+|w|)", E["hrate"], 1e-4 * fs)
    pct("|d (omega_x + hinge_x rate)| / (1+|.|)", E["sum_rate"], 6e-5 * fs); pct("|d accelerometer| / (1+|a|)", E["acc"], 1e-4 * fs); pct("|d act|", E["act"], 1e-6)
    env.close()

# obs / reward on identical states
qpos, qvel, act, actions, params = rand_inputs(N, 2.0)
for cls, rew in (("LocalFrameRPYParamsEnv", "distance_energy_reward"), ("BaseDroneEnv", "default_reward_fcn"), ("LocalFrameRmParamsEnv", "distance_energy_reward_pendulum_en3")):
    C = M.BaseDroneEnv if cls == "BaseDroneEnv" else getattr(M.observation_wrappers, cls)
    rfn = getattr(M.rewards, rew, None) or M.rewards.distance_energy_reward
    env = C(dict(M.base_config, num_drones=N, reward_fcn=rfn, random_params=False, max_distance=100))
    env.drone_params = [dict(zip(PK, p)) for p in params]
    env.set_state(qpos, qvel, act)
    q0, v0, a0, s0, _ = env.get_state()
    obs, r, _ = env.evaluate_tensor(torch.as_tensor(actions, device="cuda"))
    obs, r = obs.cpu().numpy().astype(np.float64), r.cpu().numpy().astype(np.float64)
    prm = env.drone_params
    eo, er = [], []
    for i in range(N):
        m = O.compile_model(np.array(list(prm[i].values())), True, 100, True)
        st = O.drone_state(m, q0[i], v0[i], a0[i], s0[i], [0, 0, 15, 0])
        o = O.obs(O.OBS_IDS[cls], st, [0, 0, 15, 0])
        rr = O.reward(O.REWARD_IDS[rfn.__name__], st, actions.astype(np.float32).astype(np.float64)[i], 0, [0, 0, 15, 0], 100.0)
        eo.append((np.abs(obs[i] - o) / (1 + np.abs(o))).max()); er.append(abs(r[i] - rr) / (1 + abs(rr)))
    print(f"## obs / reward on identical states: {cls} + {rfn.__name__}")
    pct("|d obs| / (1+|x|)", eo, 1e-5); pct("|d reward| / (1+|r|)", er, 5e-5)
    env.close()

# 100-step open-loop trajectory
n = 512
qpos, qvel, act, actions, params = rand_inputs(n, 0.5)
env = M.BaseDroneEnv(dict(M.base_config, num_drones=n, random_params=False, max_distance=1e6, max_steps=10 ** 6))
env.drone_params = [dict(zip(PK, p)) for p in params]
env.set_state(qpos, qvel, act)
q0, v0, a0, _, _ = env.get_state()
hov = np.full((n, 4), 0.49) + rng.normal(size=(n, 4)) * 0.03
a_t = torch.as_tensor(hov, device="cuda", dtype=torch.float32)
for _ in range(100):
    env.step_tensor(a_t)
q1, v1, _, _, _ = env.get_state()
prm = env.drone_params
ep = []
for i in range(n):
    m = O.compile_model(np.array(list(prm[i].values())), True, 100, True)
    oq, ov, oa, _ = O.step(m, q0[i], v0[i], a0[i], 0.1 + 0.9 * hov.astype(np.float32).astype(np.float64)[i], 100)
    ep.append(np.abs(q1[i, :3] - oq[:3]).max())
print("## 100-step open-loop trajectory (near-hover actions)")
pct("|d pos| after 100 steps [m]", ep, 1e-4)
