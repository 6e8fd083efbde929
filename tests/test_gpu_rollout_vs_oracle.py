"""Whole ROLLOUTS of the CUDA path against the CPU oracle, randomised over the configuration surface: observation wrapper x
reward function x frame_skip x frequency x pendulum x per-env parameters x per-env setpoints, T steps each WITH in-kernel
resets.  The oracle replays the same protocol (orc_vector_step + RLlib's reset_at round trip on the same per-env Philox reset
streams), so in the FP64 build every state, observation, reward, truncation flag and counter of every step must agree to
1e-9 - a wrong branch in any combination of the generic kernel's run-time options shows here.  FP32: first divergence only
through rounding (truncation ties), so the per-step comparison re-synchronises the oracle to the device state each step.
"""
import numpy as np
import pytest

from conftest import has_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]

PKEYS = ("mass", "arm_len", "motor_force", "motor_tau", "pendulum_len", "weight_mass")
PEND_REWARDS = ["default_reward_fcn", "distance_reward_fcn", "distance_energy_reward", "distance_energy_reward_pendulum_angle",
                "distance_energy_reward_pendulum_angle2", "distance_energy_reward_pendulum_angle3", "distance_energy_reward_pendulum_en",
                "distance_energy_reward_pendulum_en2", "distance_energy_reward_pendulum_en3", "distance_energy_reward_pendulum_en4",
                "distance_time_energy_reward", "reward_1", "reward_pendulum_dist", "reward_pendulumDistHeading", "reward_2",
                "reward_2_penergy", "reward_3"]
NOPEND_REWARDS = ["default_reward_fcn", "distance_reward_fcn", "distance_energy_reward", "distance_time_energy_reward"]


def _cases():
    import mujoco_drone_b200 as M
    rng = np.random.default_rng(99)
    pend_wrappers = [k for k in M.observation_wrappers.WRAPPERS if "NoPend" not in k] + ["BaseDroneEnv"]
    out = []
    for k in range(14):
        pend = k % 5 != 4
        wrapper = str(rng.choice(pend_wrappers)) if pend else str(rng.choice(["BaseDroneEnv", "LocalFramePRYaccNoPendEnv"]))
        reward = str(rng.choice(PEND_REWARDS if pend else NOPEND_REWARDS))
        out.append(dict(wrapper=wrapper, reward=reward, pend=pend, frame_skip=int(rng.integers(1, 4)), frequency=int(rng.choice([100, 200, 500])),
                        random_params=bool(rng.integers(0, 2)), per_env_reference=bool(rng.integers(0, 2)), n=int(rng.choice([97, 333, 1025])),
                        max_steps=int(rng.integers(4, 9)), max_distance=float(rng.choice([0.7, 1.2, 2.5])), seed=int(rng.integers(1, 1000)),
                        offset=int(rng.integers(0, 5000)), precision="fp64" if k % 3 else "fp32"))
    return out


@pytest.mark.parametrize("case", _cases(), ids=lambda c: f"{c['wrapper']}-{c['reward']}-fs{c['frame_skip']}-{c['precision']}")
def test_rollout_with_resets_matches_oracle(oracle, case):
    import torch
    import mujoco_drone_b200 as M
    c = case
    n, pend, fp64 = c["n"], c["pend"], c["precision"] == "fp64"
    cls = M.BaseDroneEnv if c["wrapper"] == "BaseDroneEnv" else getattr(M.observation_wrappers, c["wrapper"])
    sd, reference, start = 0.4, [0.3, -0.2, 15.2, 0.5], [0.1, 0.1, 15.0, -0.3]
    cfg = dict(M.base_config, num_drones=n, precision=c["precision"], skip_steps=c["frame_skip"], frequency=c["frequency"], pendulum=pend,
               random_params=c["random_params"], param_difficulty=1.0, per_env_reference=c["per_env_reference"], auto_reset=True,
               max_steps=c["max_steps"], max_distance=c["max_distance"], seed=c["seed"], env_id_offset=c["offset"], state_difficulty=sd,
               reward_fcn=getattr(M.rewards, c["reward"]), reference=reference, start_pos=start, angle_variance=[0.3, 0.2])
    env = cls(cfg)
    env.reset_tensor()
    dt = torch.float64 if fp64 else torch.float32
    if c["per_env_reference"]:                                      # distinct setpoints per env
        g = torch.Generator(device="cuda").manual_seed(c["seed"])
        for _ in range(3):
            env.control_reference_tensor((torch.rand((4, n), device="cuda", generator=g, dtype=dt) * 2 - 1).mul(100).round().div(100))
    params = np.array([list(d.values()) for d in env.drone_params])
    cpu = oracle.CpuVecEnv(params, pend, float(c["frequency"]), c["frame_skip"], True)
    qpos, qvel, act, _, ns = env.get_state()
    cpu.qpos[:], cpu.qvel[:], cpu.act[:] = qpos, qvel, act
    cpu.num_steps[:] = ns
    rc = oracle.make_reset_cfg(start, sd * 2, [sd * 0.3, sd * 0.2], [sd] * 3, [sd] * 3, [0.5 * sd] * 2, [0.5 * sd] * 2, True, pend)
    if c["per_env_reference"]:
        refs = env.rows(M._lib.BUF_REFERENCE)[:, :n].cpu().numpy().astype(np.float64).T.copy()
        refs[:, :3] += np.array(start[:3])
    else:
        refs = np.array(reference, dtype=np.float64)
    rid, oid = oracle.REWARD_IDS[c["reward"]], oracle.OBS_IDS[c["wrapper"]]
    tol = 1e-9 if fp64 else None
    g = torch.Generator(device="cuda").manual_seed(1 + c["seed"])
    n_trunc = 0
    for t in range(12):
        a = torch.rand((n, 4), device="cuda", generator=g, dtype=dt)
        obs, rew, trunc = env.step_tensor(a)
        obs, rew, trunc = obs.cpu().numpy().astype(np.float64), rew.cpu().numpy().astype(np.float64), trunc.cpu().numpy().astype(bool)
        oobs, orew, otr = cpu.step(a.cpu().numpy().astype(np.float64), refs, rid, oid, c["max_distance"], c["max_steps"])
        if fp64:
            assert (trunc == otr).all(), (t, np.nonzero(trunc != otr))
            assert (np.abs(obs - oobs) <= tol * (1 + np.abs(oobs))).all(), (t, np.abs(obs - oobs).max())
            assert (np.abs(rew - orew) <= tol * (1 + np.abs(orew))).all(), t
        else:
            # rounding-level ties of the distance test are the only legitimate disagreement
            d = np.linalg.norm(cpu.qpos[:, :3] - (refs[:, :3] if refs.ndim == 2 else refs[:3]), axis=1)
            tie = np.abs(d - c["max_distance"]) < 1e-5
            assert (trunc == otr)[~tie].all(), t
            ok = trunc == otr
            assert (np.abs(obs - oobs)[ok] <= 3e-4 * c["frame_skip"] * (1 + np.abs(oobs)[ok])).all(), (t, np.abs(obs - oobs)[ok].max())
            assert (np.abs(rew - orew)[ok] <= 3e-4 * c["frame_skip"] * (1 + np.abs(orew)[ok])).all(), t
        cpu.reset_truncated(rc, env.seed_value, c["offset"], trunc)          # RLlib's reset_at round trip, the kernel's own Philox streams
        n_trunc += int(trunc.sum())
        qpos, qvel, act, sens, ns = env.get_state()
        if fp64:
            assert np.abs(qpos - cpu.qpos).max() <= 1e-9 and np.abs(qvel - cpu.qvel).max() <= 1e-8 and np.abs(act - cpu.act).max() <= 1e-10, t
            assert (ns == cpu.num_steps).all()
        else:                                                       # keep the FP64 oracle on the FP32 trajectory: compare step by step
            cpu.qpos[:], cpu.qvel[:], cpu.act[:] = qpos, qvel, act
            cpu.num_steps[:] = ns
            cpu.reset_count[:] = env.rows(M._lib.BUF_RESET_COUNT)[0].cpu().numpy().astype(np.uint32)
    assert n_trunc > n // 2                                         # the reset path was exercised
    assert env.episode_stats()["n_episodes"] == n_trunc
    env.close()
