#!/usr/bin/env python
"""Memory-safety workload.  compute-sanitizer is closed on this pool, so the check is our own: with DSIM_GUARD=1 every device
buffer of a handle is allocated at its EXACT size between two 4 KB canary regions and `guard_check()` counts overwritten
canary bytes (out-of-bounds device writes: bulk stores of ragged last pages, the observation block, per-env scalars).
Also usable under compute-sanitizer (memcheck / racecheck / synccheck) where that is allowed.  Ragged batch sizes through every step-kernel
instantiation (specialised C2 / C3 / C4, generic, FP64, no pendulum, floor contact), in-kernel resets, evaluate, host entry point, both
dependency modes, the auxiliary kernels and the two policy kernels.      compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mujoco_drone_b200 as M

W = M.observation_wrappers
checked = 0
cases = [(W.LocalFrameRPYParamsEnv, M.rewards.distance_energy_reward, dict(param_difficulty=1.0, random_params=True), 4096 + 33),
         (W.LocalFrameRPYParamsEnv, M.rewards.distance_energy_reward, dict(param_difficulty=1.0, random_params=True), 1),
         (W.LocalFrameRPYParamsEnv, M.rewards.distance_energy_reward, dict(param_difficulty=1.0, random_params=True), 76000 + 7),
         (W.LocalFrameRPYEnv, M.rewards.distance_reward_fcn, dict(per_env_reference=True, random_params=False), 1000),
         (M.BaseDroneEnv, M.rewards.default_reward_fcn, dict(), 97),
         (W.LocalFrameRmParamsEnv, M.rewards.reward_2, dict(skip_steps=2), 65),
         (M.BaseDroneEnv, M.rewards.default_reward_fcn, dict(precision="fp64"), 70),
         (M.BaseDroneEnv, M.rewards.default_reward_fcn, dict(pendulum=False), 33),
         # floor contact: the instantiations with the slow path, every drone within reach of the floor (z = 1 m, 1.2 m pendulum)
         (W.LocalFrameRPYParamsEnv, M.rewards.distance_energy_reward, dict(ground_contact=True, start_pos=[0, 0, 1.0, 0], reference=[0, 0, 1.0, 0], max_distance=100.0), 1000 + 5),
         (M.BaseDroneEnv, M.rewards.default_reward_fcn, dict(ground_contact=True, precision="fp64", start_pos=[0, 0, 1.0, 0], reference=[0, 0, 1.0, 0], max_distance=100.0), 70),
         (M.BaseDroneEnv, M.rewards.default_reward_fcn, dict(ground_contact=True, pendulum=False, start_pos=[0, 0, 0.1, 0], reference=[0, 0, 0.1, 0], max_distance=100.0), 33)]
for ready in (False, True):
    for cls, rew, extra, n in cases:
        cfg = dict(M.base_config, num_drones=n, reward_fcn=rew, auto_reset=True, max_steps=3, max_distance=1.0, inputs_ready=ready)
        cfg.update(extra)
        envs = [cls(dict(cfg, env_id_offset=k * n)) for k in range(2)]
        dt = torch.float64 if extra.get("precision") == "fp64" else torch.float32
        for e in envs:
            e.reset_tensor()
        for t in range(6):
            for e in envs:
                if extra.get("per_env_reference"):
                    e.control_reference_tensor(torch.rand((4, n), device="cuda", dtype=dt) * 2 - 1)
                e.step_tensor(torch.rand((n, 4), device="cuda", dtype=dt))
        envs[0].evaluate_tensor(torch.rand((n, 4), device="cuda", dtype=dt))
        envs[0].get_drone_states()
        if dt == torch.float32:
            envs[0].inputs_ready = False
            envs[0].step_host(np.random.rand(n, 4).astype(np.float32))
        torch.cuda.synchronize()
        assert envs[0].episode_stats()["n_nonfinite"] == 0
        for e in envs:
            g = e.guard_check()
            assert g in (0, -1), (cls.__name__, n, ready, g)
            checked += g == 0
            e.close()
pol = M.policy.make_rma_full().cuda()
obs, prev = torch.randn((300, 22), device="cuda"), torch.rand((300, 4), device="cuda")
f = M.policy.FusedRMAFull(pol, device=0); lg, _ = f(obs, prev); f.check(); f.close()
g = M.policy.FP32RMAFull(pol, device=0); g(obs, prev); g.close()
M.policy.beta_policy(lg, seed=1)
torch.cuda.synchronize()
print(f"sanitize_smoke ok; handles with intact canaries: {checked}")
