"""The CPU oracle (oracle/dsim_oracle.c) against fixtures produced by the UNMODIFIED reference Python
(tools/make_golden.py): this is what pins the oracle for everything except the mj_step arithmetic."""
import numpy as np
import pytest

from conftest import golden


def test_quat2rpy_matches_reference(oracle):
    g = golden("transform.npz")
    out = np.array([oracle.quat2rpy(q) for q in g["quats"]])
    ref = g["quat2rpy"]
    ns = int(g["n_special"])
    # regular attitudes: 1e-12; the gimbal-lock set (pitch within 1e-5 of +-pi/2) is ill-conditioned by nature
    np.testing.assert_allclose(out[:-ns], ref[:-ns], atol=1e-12, rtol=0)
    d = np.abs(out[-ns:] - ref[-ns:])
    d = np.minimum(d, 2 * np.pi - d)
    assert d.max() < 1e-6


def test_rpy2quat_dcm_pendulum(oracle):
    g = golden("transform.npz")
    np.testing.assert_allclose(np.array([oracle.rpy2quat(r) for r in g["rpy_in"]]), g["rpy2quat"], atol=1e-15)
    np.testing.assert_allclose(np.array([oracle.quat2dcm(q) for q in g["quats"]]), g["quat2dcm"], atol=1e-14)
    np.testing.assert_allclose(np.array([oracle.pendulumrp2quat(p) for p in g["prp_in"]]), g["pendulumrp2quat"], atol=1e-15)


@pytest.mark.parametrize("name", [
    "default_reward_fcn", "distance_reward_fcn", "distance_energy_reward", "distance_energy_reward_pendulum_angle",
    "distance_energy_reward_pendulum_angle2", "distance_energy_reward_pendulum_angle3",
    "distance_energy_reward_pendulum_en", "distance_energy_reward_pendulum_en2", "distance_energy_reward_pendulum_en3",
    "distance_energy_reward_pendulum_en4", "distance_time_energy_reward", "reward_1", "reward_pendulum_dist",
    "reward_pendulumDistHeading", "reward_2", "reward_2_penergy", "reward_3"])
def test_rewards_match_reference(oracle, name):
    g = golden("rewards.npz")
    rid = oracle.REWARD_IDS[name]
    out = np.array([oracle.reward(rid, g["states"][i], g["actions"][i], int(g["num_steps"][i]), g["reference"],
                                  float(g["max_distance"])) for i in range(len(g["states"]))])
    np.testing.assert_allclose(out, g["out_" + name], rtol=1e-12, atol=1e-11)


def test_obs_wrappers_match_reference(oracle):
    g = golden("obs.npz")
    for name, oid in oracle.OBS_IDS.items():
        ref = g["out_" + name]
        if name == "LocalFramePRYaccParamsNoPendEnv":           # NameError in the reference (Q14)
            assert ref.size == 0
            with pytest.raises(NameError):
                oracle.obs(oid, g["states"][0], g["reference"])
            continue
        out = np.array([oracle.obs(oid, s, g["reference"]) for s in g["states"]])
        assert out.shape == ref.shape, name
        np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-12, err_msg=name)
    assert g["out_LocalFrameFullStateZvecEnv"].shape[1] == 24      # declared 23, emits 24 (Q14)


def test_termination_bit_exact(oracle):
    g = golden("termination.npz")
    out = np.array([oracle.termination(s, g["reference"], float(g["max_distance"]), int(n), int(g["max_steps"]))
                    for s, n in zip(g["states"], g["num_steps"])])
    assert (out == g["out"]).all()
    assert out.any() and (~out).any()


@pytest.mark.parametrize("key,pend", [("pend", True), ("nopend", False)])
def test_drone_state_layout(oracle, key, pend):
    g = golden("drone_states.npz")
    n = g[key + "_params"].shape[0]
    nq, nv = (9, 8) if pend else (7, 6)
    for i in range(n):
        m = oracle.compile_model(g[key + "_params"][i], pend, 100.0, True)
        s = oracle.drone_state(m, g[key + "_qpos"][nq * i:nq * (i + 1)], g[key + "_qvel"][nv * i:nv * (i + 1)],
                               g[key + "_act"][4 * i:4 * i + 4], g[key + "_sens"][3 * i:3 * i + 3], g[key + "_ref"])
        np.testing.assert_allclose(s, g[key + "_states"][i], atol=1e-12)


def test_protocol_replay(oracle):
    """BaseDroneEnv.vector_step orchestration (BaseDroneEnv.py:259-294) replayed on the oracle from the
    pre-step MjData the reference saw: ctrl remap, frame_skip, counters, truncation, reward, obs."""
    g = golden("protocol.npz")
    T, N = g["actions"].shape[:2]
    fs, ms, md = int(g["frame_skip"]), int(g["max_steps"]), float(g["max_distance"])
    for t in range(T):
        regen_step = int(g["total_steps"][t]) == 0          # total_steps reset to 0 -> regen happened (Q8)
        params_pre = g["params"][t - 1] if t > 0 else g["params0"]
        env = oracle.CpuVecEnv(params_pre, True, float(g["frequency"]), fs, True)
        env.qpos[:] = g["qpos_pre"][t].reshape(N, 9)
        env.qvel[:] = g["qvel_pre"][t].reshape(N, 8)
        env.act[:] = g["act_pre"][t].reshape(N, 4)
        prev_ns = g["num_steps"][t - 1].copy() if t > 0 else np.zeros(N, dtype=np.int64)
        if t > 0:
            prev_ns[g["reset_idx"][t - 1]] = 0
        env.num_steps[:] = prev_ns
        obs, rew, trunc = env.step(g["actions"][t], g["reference"], oracle.REWARD_IDS["distance_energy_reward"],
                                   oracle.OBS_IDS["LocalFrameRPYParamsEnv"], md, ms)
        np.testing.assert_allclose(rew, g["rewards"][t], rtol=1e-12, atol=1e-12)
        if regen_step:
            assert g["truncated"][t].all()                   # np.ones(N) on regen
            assert (g["num_steps"][t] == 0).all()
        else:
            assert (trunc == g["truncated"][t]).all()
            np.testing.assert_allclose(obs, g["obs"][t], rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(env.qpos.ravel(), g["qpos_after"][t], atol=1e-14)
            assert (env.num_steps == g["num_steps"][t]).all()
        # Q1: reset_at returns the stale (terminal / post-regen) observation
        idx = g["reset_idx"][t]
        np.testing.assert_allclose(g["reset_obs"][t][idx], g["obs"][t][idx], atol=0)
    assert g["truncated"].any()
    assert (g["total_steps"] == 0).any()


def test_control_reference_matches_reference(oracle):
    """orc_control_reference vs BaseDroneEnv.control_reference (:151-172) run under stubs with a fake joystick
    (tools/make_golden_r2.py): dead-zone edges, yaw wrap, clip - bit for bit"""
    g = golden("control_reference.npz")
    start, ref, last = g["start"], None, -1
    moved = 0
    for a, want, k in zip(g["axes"], g["reference"], g["seq_id"]):
        if k != last:
            ref, last = start.copy(), k
        new = oracle.control_reference(ref, a, start[:3])
        assert np.array_equal(new, want), (k, a, new, want)
        moved += int(np.any(new[:3] != ref[:3]))
        ref = new
    assert 100 < moved < len(g["axes"])              # both branches of the dead zones are exercised
