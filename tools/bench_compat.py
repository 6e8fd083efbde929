#!/usr/bin/env python
"""Latency of the RLlib-compatible path (lists of float64 rows, reset_at round trips) at the reference's own shape:
64 drones per env object (train_RMA.py:60-64).  usage: bench_compat.py [num_drones]"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mujoco_drone_b200 as M

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = dict(M.base_config, num_drones=n, reward_fcn=M.rewards.distance_energy_reward, max_steps=1024, state_difficulty=0.3, param_difficulty=1.0)
env = M.observation_wrappers.LocalFrameRPYParamsEnv(cfg)
obs, _ = env.vector_reset()
rng = np.random.default_rng(0)
acts = [list(rng.uniform(0, 1, size=(n, 4))) for _ in range(16)]
for k in range(50):
    obs, rew, dones, trunc, infos = env.vector_step(acts[k % 16])
K, resets = 2000, 0
t0 = time.perf_counter()
for k in range(K):
    obs, rew, dones, trunc, infos = env.vector_step(acts[k % 16])
    for i, t in enumerate(trunc):
        if t:
            env.reset_at(i)
            resets += 1
dt = time.perf_counter() - t0
print(f"compat vector_step, {n} drones: {dt / K * 1e6:.1f} us/call -> {n * K / dt:.3e} env-steps/s per env object ({resets} reset_at calls)")
env.close()
