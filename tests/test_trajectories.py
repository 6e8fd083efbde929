"""Setpoint / trajectory generators (reference evaluation.py:135-152): host mirrors against the reference's own outputs
(tests/golden/trajectories.npz), device evaluation (dsim_trajectory_reference) against the host mirrors."""
import numpy as np
import pytest

from conftest import golden, has_cuda


def test_host_generators_match_reference():
    import mujoco_drone_b200 as M
    g = golden("trajectories.npz")
    t, c = M.trajectories.gen_circle_trajectory(T=3, f=0.7, r=1.5, h=14.5)
    np.testing.assert_array_equal(t, g["t_circle"]); np.testing.assert_array_equal(c, g["circle"])
    t, s = M.trajectories.gen_step_trajectory(step_time=1.2, duration=3, start_pos=[0, 0, 15, 0], end_pos=[1, -1, 16, 0.5])
    np.testing.assert_array_equal(t, g["t_step"]); np.testing.assert_array_equal(s, g["step"])
    t, r = M.trajectories.gen_ramp_trajectory(start_time=0.8, duration=3, start_pos=[0, 0, 15, 0], end_pos=[1, -1, 16, 0.5])
    np.testing.assert_array_equal(t, g["t_ramp"]); np.testing.assert_allclose(r, g["ramp"], rtol=0, atol=1e-15)


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_device_trajectories_match_host(precision):
    import mujoco_drone_b200 as M
    n, dt = 300, 0.01
    env = M.BaseDroneEnv(dict(M.base_config, num_drones=n, per_env_reference=True, precision=precision, start_pos=[0.5, -0.5, 15, 0]))
    env.vector_reset()
    tol = 1e-12 if precision == "fp64" else 2e-6
    cases = [("circle", dict(f=0.7, r=1.5, h=14.5), M.trajectories.gen_circle_trajectory(T=3, f=0.7, r=1.5, h=14.5)[1]),
             ("step", dict(step_time=1.2, start_pos=[0, 0, 15, 0], end_pos=[1, -1, 16, 0.5]),
              M.trajectories.gen_step_trajectory(step_time=1.2, duration=3, start_pos=[0, 0, 15, 0], end_pos=[1, -1, 16, 0.5])[1]),
             ("ramp", dict(start_time=0.8, duration=3, start_pos=[0, 0, 15, 0], end_pos=[1, -1, 16, 0.5]),
              M.trajectories.gen_ramp_trajectory(start_time=0.8, duration=3, start_pos=[0, 0, 15, 0], end_pos=[1, -1, 16, 0.5])[1])]
    for kind, kw, host in cases:
        tr = M.trajectories.TrajectoryReference(env, kind, phase_step=dt, **kw)    # env i is i steps ahead: one launch covers 300 samples
        tr.advance(0.0)
        states = np.array(env.get_drone_states())                                  # reference columns 23:27 of the 33-row
        np.testing.assert_allclose(states[:, 23:27], host[:n], atol=tol * 20, rtol=0)
        obs, rew, _, trunc, _ = env.vector_step(list(np.full((n, 4), 0.5)))
        assert np.isfinite(np.array(obs)).all() and len(trunc) == n
    env.close()
