#!/usr/bin/env python
"""BASELINE.md §4.3 row "reference Python + restated mj_step": the reference's OWN BaseDroneEnv.vector_step / reset_at
(BaseDroneEnv.py:259-351) and observation wrapper, executed unmodified under stub modules (oracle/ref_stubs.py), with the
FP64 C restatement of mj_step plugged in as `do_simulation` (one C call per vector_step over all drones of the env object,
single-threaded like mujoco.mj_step on the reference's N-drone model).  BUILD CONTAINER ONLY (/root/reference is not on the
GPU box); the measured row is recorded in BASELINE.md.

    python tools/bench_reference_python.py [--drones 64] [--seconds 10] [--procs 1]
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(drones, seconds, seed, out):
    from oracle import ref_stubs
    from oracle import oracle as O
    ref = ref_stubs.load_reference()
    B, OW, RW = ref.BaseDroneEnv, ref.observation_wrappers, ref.rewards
    L = O.lib()
    state = {}

    def physics(env, ctrl, n_frames):
        if "models" not in state or state["params"] is not env.drone_params:      # (re)compile on regen, like make_sim -> MjModel
            ms = (O.OrcModel * env.num_drones)()
            for i, d in enumerate(env.drone_params):
                p = np.ascontiguousarray(list(d.values()), dtype=np.float64)
                L.orc_compile(O._dp(p), int(env.pendulum), float(env.frequency), 1, C.byref(ms[i]))
            state["models"], state["params"] = ms, env.drone_params
        c = np.ascontiguousarray(ctrl, dtype=np.float64)
        L.orc_step_batch(env.num_drones, state["models"], O._dp(env.data.qpos), O._dp(env.data.qvel), O._dp(env.data.act), O._dp(c),
                         O._dp(env.data.sensordata), int(n_frames))
    # train_RMA.py:66-75 environment: LocalFrameRPYParamsEnv + distance_energy_reward, param_difficulty 1, state_difficulty 0.3
    cfg = dict(B.base_config, num_drones=drones, reward_fcn=RW.distance_energy_reward, param_difficulty=1.0, state_difficulty=0.3,
               max_steps=1024, regen_env_at_steps=None, seed=seed)
    env = ref_stubs.make_env(OW.LocalFrameRPYParamsEnv, cfg, physics=physics)
    env.vector_reset()
    rng = np.random.default_rng(seed)
    steps, resets, t0 = 0, 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        a = rng.uniform(0, 1, size=(drones, 4))
        obs, rew, dones, trunc, infos = env.vector_step(list(a))
        for i in range(drones):                      # RLlib's reset protocol
            if trunc[i]:
                env.reset_at(i)
                resets += 1
        steps += 1
    dt = time.perf_counter() - t0
    out.put((drones * steps / dt, steps, resets, dt))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--drones", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--procs", type=int, default=1)
    a = ap.parse_args()
    q = mp.Queue()
    ps = [mp.Process(target=worker, args=(a.drones, a.seconds, 42 + k, q)) for k in range(a.procs)]
    [p.start() for p in ps]
    res = [q.get() for _ in ps]
    [p.join() for p in ps]
    total = sum(r[0] for r in res)
    print(json.dumps({"row": "reference Python (BaseDroneEnv.vector_step + reset_at, unmodified, under stubs) + restated mj_step (FP64 C, single thread per env object)",
                      "drones_per_env_object": a.drones, "processes": a.procs, "host_cores": os.cpu_count(),
                      "env_steps_per_s_total": total, "env_steps_per_s_per_process": total / a.procs,
                      "us_per_drone_step": 1e6 * a.procs / total, "vector_steps": [r[1] for r in res], "resets": sum(r[2] for r in res)}))
