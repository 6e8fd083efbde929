"""Invariant / known-answer checks of the oracle's mj_step restatement.
Physics parity against real MuJoCo is UNPINNED; these tests guard the derivation itself."""
import numpy as np

NOMINAL = [1, 0.17, 7, 0.01, 1.2, 0.3]


def _rand_state(rng, scale=1.0):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    qpos = np.concatenate([[0.3, -0.2, 15.0], q, rng.normal(size=2) * 0.6])
    qvel = rng.normal(size=8) * scale
    return qpos, qvel


def test_known_answer_constants(oracle):
    """SURVEY.md A.2 table (derived by hand from env_gen.py:7-73)."""
    m = oracle.compile_model(NOMINAL, True, 100, False)
    assert abs(m.mass[2] - 1.0) < 1e-12 and abs(m.mass[3] - 0.01) < 1e-15 and abs(m.mass[4] - 0.54) < 1e-12
    assert abs(m.ipos[2][2] - 0.0024) < 1e-12
    np.testing.assert_allclose(sorted(m.inertia[2]), [8.935118e-3, 8.935118e-3, 1.768190e-2], rtol=2e-7)
    np.testing.assert_allclose(sorted(m.inertia[4]), [8.99281e-4, 0.07769778, 0.07769778], rtol=2e-7)
    assert abs(m.ipos[4][2] + 0.9333333333) < 1e-9
    assert abs(abs(m.site_pos[0][0]) - 0.170208) < 1e-6
    mr = oracle.compile_model(NOMINAL, True, 100, True)            # "%.5g" stage (env_gen.py:129)
    np.testing.assert_allclose(sorted(mr.inertia[2]), [8.934968e-3, 8.934971e-3, 1.768160e-2], rtol=2e-7)
    assert mr.site_pos[0][0] == 0.17021 and mr.site_pos[0][1] == -0.17021
    assert oracle.round_prec5(0.0166666666) == 0.016667 and oracle.round_prec5(123456.0) == 1.2346e5


def test_hover_equilibrium(oracle):
    m = oracle.compile_model(NOMINAL, True, 100, False)
    hov = sum(m.mass) * 9.81 / (4 * 7)
    assert abs(hov - 0.54305) < 1e-5
    f = oracle.forward(m, [0, 0, 15, 1, 0, 0, 0, 0, 0], np.zeros(8), [hov] * 4, [hov] * 4)
    assert np.abs(f["qacc"]).max() < 1e-12
    np.testing.assert_allclose(f["sensordata"], [0, 0, 9.81], atol=1e-12)
    assert np.abs(f["act_dot"]).max() < 1e-12


def test_mass_matrix_structure(oracle):
    rng = np.random.default_rng(1)
    m = oracle.compile_model(NOMINAL, True, 100, True)
    for _ in range(10):
        qpos, qvel = _rand_state(rng)
        M = oracle.forward(m, qpos, qvel, [0.5] * 4, [0.5] * 4)["M"]
        assert np.abs(M - M.T).max() < 1e-15
        assert np.linalg.eigvalsh(M).min() > 1e-4
        np.testing.assert_allclose(M[:3, :3], sum(m.mass) * np.eye(3), atol=1e-14)
        assert abs(M[6, 7]) < 1e-15                      # hinge axes: x of C, y of C -> orthogonal about a common point


def test_yaw_torque_sign(oracle):
    """(-1)^k gear on the yaw axis (env_gen.py:62): motors 0 and 2 give positive yaw acceleration."""
    m = oracle.compile_model(NOMINAL, False, 100, True)
    for k, sgn in enumerate((1, -1, 1, -1)):
        act = np.zeros(4)
        act[k] = 1.0
        f = oracle.forward(m, [0, 0, 15, 1, 0, 0, 0], np.zeros(6), act, act)
        assert np.sign(f["qacc"][5]) == sgn


def test_energy_and_momentum_conservation(oracle):
    rng = np.random.default_rng(2)
    m = oracle.compile_model(NOMINAL, True, 2000, False)
    m.damping[6] = m.damping[7] = 0.0
    m.density = m.viscosity = 0.0
    qpos, qvel = _rand_state(rng, 2.0)
    act = np.zeros(4)
    ke0, pe0 = oracle.energy(m, qpos, qvel)
    M0 = oracle.forward(m, qpos, qvel, act, act)["M"]
    p0 = M0[:3] @ qvel
    T = 2000                                         # 1 s
    qp, qv, a, _ = oracle.step(m, qpos, qvel, act, np.zeros(4), T)
    ke1, pe1 = oracle.energy(m, qp, qv)
    assert abs((ke1 + pe1) - (ke0 + pe0)) < 2e-2 * (abs(ke0) + 1)          # O(h) drift of semi-implicit Euler
    p1 = oracle.forward(m, qp, qv, a, a)["M"][:3] @ qv
    np.testing.assert_allclose(p1 - p0, sum(m.mass) * np.array([0, 0, -9.81]) * T * m.timestep, atol=2e-2)
    # halving the time step halves the drift (first-order integrator) -> the bias forces are consistent with M
    m2 = oracle.compile_model(NOMINAL, True, 4000, False)
    m2.damping[6] = m2.damping[7] = 0.0
    m2.density = m2.viscosity = 0.0
    qp2, qv2, _, _ = oracle.step(m2, qpos, qvel, act, np.zeros(4), 2 * T)
    ke2, pe2 = oracle.energy(m2, qp2, qv2)
    d1, d2 = abs((ke1 + pe1) - (ke0 + pe0)), abs((ke2 + pe2) - (ke0 + pe0))
    assert d2 < 0.7 * d1 + 1e-9


def test_accelerometer_free_fall_and_spin(oracle):
    m = oracle.compile_model(NOMINAL, False, 100, True)
    m.density = m.viscosity = 0.0
    f = oracle.forward(m, [0, 0, 15, 1, 0, 0, 0], np.zeros(6), np.zeros(4), np.zeros(4))
    np.testing.assert_allclose(f["sensordata"], 0, atol=1e-12)            # free fall: proper acceleration 0
    np.testing.assert_allclose(f["qacc"][:3], [0, 0, -9.81], atol=1e-12)
    # pure spin about body z with COM on the z axis: centripetal term at the site (0,0,-0.0125) vanishes on the axis
    qvel = np.array([0, 0, 0, 0, 0, 5.0])
    f = oracle.forward(m, [0, 0, 15, 1, 0, 0, 0], qvel, np.zeros(4), np.zeros(4))
    np.testing.assert_allclose(f["sensordata"], 0, atol=1e-10)


def test_implicit_damping_differs_only_through_hinges(oracle):
    """mj_EulerSkip: (M + h*diag(B)) qacc = qfrc_smooth; with zero hinge damping the step is explicit."""
    rng = np.random.default_rng(3)
    qpos, qvel = _rand_state(rng)
    m = oracle.compile_model(NOMINAL, True, 100, True)
    f = oracle.forward(m, qpos, qvel, [0.5] * 4, [0.6] * 4)
    qp, qv, act, sens = oracle.step(m, qpos, qvel, [0.5] * 4, [0.6] * 4, 1)
    h = m.timestep
    H = f["M"] + h * np.diag(list(m.damping))
    qacc_i = np.linalg.solve(H, f["qfrc_smooth"])
    np.testing.assert_allclose(qv, qvel + h * qacc_i, atol=1e-12)
    np.testing.assert_allclose(act, 0.5 + h * (0.6 - 0.5) / m.tau[0], atol=1e-15)
    np.testing.assert_allclose(qp[:3], qpos[:3] + h * qv[:3], atol=1e-15)
    np.testing.assert_allclose(sens, f["sensordata"], atol=0)              # sensor = pre-integration, explicit qacc
    assert np.abs(qacc_i - f["qacc"]).max() > 1e-6


def test_philox_known_answer(oracle):
    """Random123 KAT for philox4x32-10."""
    out = oracle.philox4x32([0, 0, 0, 0], [0, 0])
    assert [hex(x) for x in out] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    out = oracle.philox4x32([0xffffffff] * 4, [0xffffffff] * 2)
    assert [hex(x) for x in out] == ['0x408f276d', '0x41c83b0e', '0xa20bc7c6', '0x6d5451fd']
    out = oracle.philox4x32([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])
    assert [hex(x) for x in out] == ['0xd16cfe09', '0x94fdcceb', '0x5001e420', '0x24126ea1']


def test_sampler_distribution_matches_reference(oracle):
    """Philox-based sample_state / generate_drone_params vs 4000 draws of the reference's own
    BaseDroneEnv.sample_state (PCG64): same distribution (two-sample KS), not the same stream."""
    from scipy import stats
    from conftest import golden
    g = golden("sampling.npz")
    sd = float(g["state_difficulty"])
    cfg = oracle.make_reset_cfg(g["start_pos"], sd * float(g["max_random_offset"]), sd * g["angle_variance"],
                                sd * g["vel_variance"], sd * g["ang_vel_variance"], sd * g["pendulum_rp_variance"],
                                sd * g["pendulum_ang_vel_variance"], True, True,
                                [1, 0.17, 7, 0.01, 1.2, 0.3], [0.1, 0.02, 1, 0.0025, 0.2, 0.05],
                                float(g["param_difficulty"]), True)
    n = 4000
    S = [oracle.sample_state(cfg, 42, i, 0) for i in range(n)]
    qpos, qvel = np.array([s[0] for s in S]), np.array([s[1] for s in S])
    for k in (0, 1, 2, 7, 8):
        assert stats.ks_2samp(qpos[:, k], g["qpos"][:, k]).pvalue > 1e-3, k
    for k in range(8):
        assert stats.ks_2samp(qvel[:, k], g["qvel"][:, k]).pvalue > 1e-3, k
    rpy = np.array([oracle.quat2rpy(q) for q in qpos[:, 3:7]])
    rpy_ref = np.array([oracle.quat2rpy(q) for q in g["qpos"][:, 3:7]])
    for k in range(3):
        assert stats.ks_2samp(rpy[:, k], rpy_ref[:, k]).pvalue > 1e-3, k
    r = np.linalg.norm(qpos[:, :3] - g["start_pos"][:3], axis=1)
    assert r.max() <= sd * float(g["max_random_offset"]) + 1e-12
    P = np.array([oracle.sample_params(cfg, 42, i, 0) for i in range(n)])
    for k in range(6):
        assert stats.ks_2samp(P[:, k], g["params"][:, k]).pvalue > 1e-3, k
    # streams are keyed by (seed, env, epoch): reproducible and distinct
    assert np.array_equal(oracle.sample_state(cfg, 42, 7, 3)[0], oracle.sample_state(cfg, 42, 7, 3)[0])
    assert not np.array_equal(oracle.sample_state(cfg, 42, 7, 3)[0], oracle.sample_state(cfg, 42, 7, 4)[0])
