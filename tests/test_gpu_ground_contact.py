"""Floor contact (env_gen.py:14-21,97) - the CUDA slow path (csrc/dsim_contact.cuh, body-frame coordinates,
dense 8 x 8 Newton) against the oracle's MuJoCo-layout restatement (oracle/dsim_oracle.c: world-frame Jacobians from cdof).
Both restate MuJoCo's contact model from its documentation; like the rest of the physics it is UNPINNED against MuJoCo itself
(tests/test_mujoco_pin.py switches on when `import mujoco` works), so what is asserted here is CUDA == oracle plus the
invariants a contact model must have (tests/test_ground_contact.py holds the oracle-side ones)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NOMINAL = np.array([1, 0.17, 7, 0.01, 1.2, 0.3])
KEYS = ("mass", "arm_len", "motor_force", "motor_tau", "pendulum_len", "weight_mass")


def _mk(**over):
    import mujoco_drone_b200 as M
    cfg = dict(M.base_config)
    cfg.update(dict(start_pos=[0, 0, 1.0, 0], reference=[0, 0, 1.0, 0], max_distance=100, random_params=False, ground_contact=True))
    cfg.update(over)
    return M.BaseDroneEnv(cfg)


def _near_floor(rng, n, pend, zmax=1.5):
    """random attitudes (any, incl. upside down), heights from 'core body in the floor' to 'pendulum tip just above it'"""
    q = rng.normal(size=(n, 4))
    q[: n // 4] = [1, 0, 0, 0] + 0.05 * rng.normal(size=(n // 4, 4))          # a quarter almost level
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    pos = np.stack([rng.normal(size=n), rng.normal(size=n), rng.uniform(0.0, zmax, size=n)], axis=1)
    qpos = np.concatenate([pos, q] + ([rng.normal(size=(n, 2)) * 0.5] if pend else []), axis=1)
    qvel = rng.normal(size=(n, 8 if pend else 6)) * 0.5
    act = rng.uniform(0, 1, size=(n, 4))
    actions = rng.uniform(0, 1, size=(n, 4))
    params = NOMINAL * rng.uniform(0.85, 1.15, size=(n, 6))
    if not pend:
        params[:, 4:] = 0
    return qpos, qvel, act, actions, params


def _set(env, qpos, qvel, act, params):
    env.drone_params = [dict(zip(KEYS, p)) for p in params]
    env.set_state(qpos, qvel, act, None)


def test_collision_geometry_table_matches_oracle_model(oracle):
    """the per-env geometry the device compile writes (csrc/dsim_contact.cuh GeoRow, FP64 + "%.5g") vs the geoms of the oracle's model"""
    import mujoco_drone_b200 as M
    rng = np.random.default_rng(2)
    n = 70
    params = NOMINAL * rng.uniform(0.6, 1.4, size=(n, 6))
    env = _mk(num_drones=n, precision="fp64")
    env.drone_params = [dict(zip(KEYS, p)) for p in params]
    G = env.tensor(M._lib.BUF_GEOMETRY).cpu().numpy()
    assert G.shape == (39, n)
    for i in range(n):
        m = oracle.compile_model(params[i], True, 100, True)
        geoms = [m.geom[k] for k in range(m.ngeom)]
        core, front, pole, weight = geoms[0], geoms[9], geoms[15], geoms[16]
        assert np.allclose([G[0, i], G[1, i]], [core.size[0], core.size[2]], rtol=0, atol=1e-15)
        assert np.allclose(G[2:5, i], [front.pos[0], front.size[0], front.size[1]], rtol=0, atol=1e-15)
        for a in range(4):
            arm, motor, prop = geoms[1 + 2 * a], geoms[2 + 2 * a], geoms[10 + a]
            assert np.allclose([G[5, i], G[6, i], G[7, i], G[8, i], G[9, i]], [arm.size[0], arm.size[1], prop.size[0], motor.pos[2], prop.pos[2]], rtol=0, atol=1e-15)
            assert np.allclose([G[10 + a, i], G[14 + a, i], G[18 + a, i], G[22 + a, i]], [arm.pos[0], arm.pos[1], motor.pos[0], motor.pos[1]], rtol=0, atol=1e-15)
            assert motor.pos[0] == prop.pos[0] and motor.pos[1] == prop.pos[1]
            assert np.allclose([G[26 + a, i], G[30 + a, i]], [np.cos(arm.yaw), np.sin(arm.yaw)], rtol=0, atol=1e-15)
        assert G[34, i] == 1.0
        assert np.allclose(G[35:39, i], [pole.size[1], pole.pos[2], weight.size[0], weight.pos[2]], rtol=0, atol=1e-15)
    env.close()


@pytest.mark.parametrize("precision,frame_skip,pend", [("fp64", 1, True), ("fp64", 2, False), ("fp32", 1, True), ("fp32", 1, False)])
def test_contact_step_matches_oracle(oracle, precision, frame_skip, pend):
    import torch
    rng = np.random.default_rng(11)
    n = 256
    qpos, qvel, act, actions, params = _near_floor(rng, n, pend, zmax=1.5 if pend else 0.3)
    env = _mk(num_drones=n, precision=precision, skip_steps=frame_skip, pendulum=pend)
    _set(env, qpos, qvel, act, params)
    qpos_d, qvel_d, act_d, sens0, _ = env.get_state()             # sens0: mj_forward after set_state, contact stage included
    env.step_tensor(torch.as_tensor(actions, device="cuda"))
    qp, qv, ac, sens, ns = env.get_state()
    a_in = actions.astype(np.float32).astype(np.float64) if precision == "fp32" else actions
    prm = env.drone_params
    # FP32: a contact force is K d(r) r / R with r the penetration (~1e-4 .. 1e-2 m) computed from a height of ~1 m: its
    # relative error is eps * 1 m / r, i.e. up to 1e-3 where the penetration is shallow, times h / m on the velocity
    # measured maxima over 8 seeds x 256 states (tools/gpu_ground_seeds.py, profiles/r02d_ground_contact.txt): FP64 1.8e-14 /
    # 1.0e-12 / 4.2e-13; FP32 8.3e-6 (hinge angles) / 5.5e-5 / 9.5e-5 -> FP32 tolerances 3-4x those
    tol = dict(pos=1e-11, vel=1e-9, acc=1e-7) if precision == "fp64" else dict(pos=3e-5, vel=2e-4, acc=3e-4)
    touching = 0
    worst = dict(pos=0.0, vel=0.0, acc=0.0)
    for i in range(n):
        p = np.array(list(prm[i].values()))
        m = oracle.compile_model(p, pend, 100, True, ground=True)
        ncon = len(oracle.collide(m, qpos_d[i]))
        touching += ncon > 0
        oqp, oqv, oact, osens = oracle.step(m, qpos_d[i], qvel_d[i], act_d[i], 0.1 + 0.9 * a_in[i], frame_skip)
        worst["pos"] = max(worst["pos"], np.abs(qp[i] - oqp).max())
        worst["vel"] = max(worst["vel"], (np.abs(qv[i] - oqv) / (1 + np.abs(oqv))).max())
        worst["acc"] = max(worst["acc"], (np.abs(sens[i] - osens) / (1 + np.abs(osens))).max())
        f0 = oracle.forward_contact(m, qpos_d[i], qvel_d[i], act_d[i], np.zeros(4))
        assert f0["ncon"] == ncon
        worst["acc"] = max(worst["acc"], (np.abs(sens0[i] - f0["sensordata"]) / (1 + np.abs(f0["sensordata"]))).max())
    print(precision, frame_skip, pend, "touching", touching, "of", n, worst)
    assert touching > n // 4                                      # the scenario does exercise the contact path
    for k in worst:
        assert worst[k] <= tol[k] * frame_skip, (k, worst[k])
    env.close()


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_drop_and_settle_matches_oracle(oracle, precision):
    """400 steps from 0.3 m above the floor: free fall, impact, bounce, rest.  FP64: the whole trajectory; FP32: the resting pose
    (a chaotic bounce amplifies rounding, the rest state is an attractor)."""
    import torch
    rng = np.random.default_rng(5)
    n = 64
    params = NOMINAL * rng.uniform(0.85, 1.15, size=(n, 6))
    params[:, 4:] = 0
    qpos = np.zeros((n, 7)); qpos[:, 2] = 0.3; qpos[:, 3] = 1
    tilt = rng.normal(size=(n, 3)) * 0.1
    qpos[:, 4:7] = tilt; qpos[:, 3:7] /= np.linalg.norm(qpos[:, 3:7], axis=1, keepdims=True)
    qvel = np.zeros((n, 6)); act = np.zeros((n, 4))
    env = _mk(num_drones=n, precision=precision, pendulum=False, max_steps=100000)
    _set(env, qpos, qvel, act, params)
    qpos_d, qvel_d, act_d, _, _ = env.get_state()
    zero = torch.full((n, 4), -1.0, device="cuda", dtype=torch.float64 if precision == "fp64" else torch.float32)   # ctrl = clamp(0.1 - 0.9) = 0
    T = 400
    for _ in range(T):
        env.step_tensor(zero)
    qp, qv, ac, sens, ns = env.get_state()
    assert np.isfinite(qp).all() and np.isfinite(qv).all()
    prm = env.drone_params
    for i in range(n):
        m = oracle.compile_model(np.array(list(prm[i].values())), False, 100, True, ground=True)
        oqp, oqv, oact, osens = oracle.step(m, qpos_d[i], qvel_d[i], act_d[i], np.zeros(4), T)
        if precision == "fp64":
            assert np.abs(qp[i] - oqp).max() < 1e-7 and np.abs(qv[i] - oqv).max() < 1e-6
        else:
            assert abs(qp[i, 2] - oqp[2]) < 2e-5                              # rest height: core box half height - penetration
            assert np.abs(qv[i]).max() < 5e-3 and np.abs(oqv).max() < 1e-3   # at rest
            assert np.abs(sens[i] - [0, 0, 9.81]).max() < 0.05               # the floor carries the weight
    env.close()


def test_far_from_the_floor_ground_contact_changes_nothing():
    """ground_contact=True with every drone out of reach of the floor: the instantiation with the slow path compiled in takes none
    of it and reproduces ground_contact=False (another instantiation of the same source: equal up to FP32 contraction order
    over the 20 steps; frame_skip 2: both runs take a generic instantiation)."""
    import torch
    n = 4128
    outs = []
    for ground in (False, True):
        import mujoco_drone_b200 as M
        cfg = dict(M.base_config)
        cfg.update(dict(num_drones=n, random_params=True, ground_contact=ground, seed=3, skip_steps=2))
        env = M.BaseDroneEnv(cfg)
        env.reset_tensor()
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        for _ in range(20):
            obs, rew, trunc = env.step_tensor(torch.rand((n, 4), device="cuda", generator=g))
        outs.append((obs.clone(), rew.clone(), trunc.clone(), env.get_state()))
        env.close()
    assert torch.equal(outs[0][2], outs[1][2])
    assert torch.allclose(outs[0][0], outs[1][0], rtol=1e-4, atol=1e-4) and torch.allclose(outs[0][1], outs[1][1], rtol=1e-4, atol=1e-4)
    for a, b in zip(outs[0][3], outs[1][3]):
        assert np.allclose(a, b, rtol=1e-4, atol=1e-4)


def test_many_drones_land_and_stay_on_the_floor():
    """Size-independent properties at a full batch: random drops, every drone ends up resting on the floor - finite, no
    geom deeper than the soft-contact penetration allows, velocities gone, accelerometer reads |g|."""
    import torch
    import mujoco_drone_b200 as M
    n = 32768
    cfg = dict(M.base_config)
    cfg.update(dict(num_drones=n, start_pos=[0, 0, 1.6, 0], reference=[0, 0, 1.6, 0], max_distance=1000, max_steps=10**6, random_params=True,
                    ground_contact=True, angle_variance=[0.5, 0.5], vel_variance=[0.5, 0.5, 0.5], ang_vel_variance=[1, 1, 1],
                    pendulum_rp_variance=[0.3, 0.3], max_random_offset=0.3, seed=9))
    env = M.BaseDroneEnv(cfg)
    env.reset_tensor()
    off = torch.full((n, 4), -1.0, device="cuda")
    for _ in range(1200):
        obs, rew, trunc = env.step_tensor(off)
    qp, qv, ac, sens, ns = env.get_state()
    assert np.isfinite(qp).all() and np.isfinite(qv).all() and np.isfinite(sens).all()
    assert not trunc.any().item()
    # body origin: not under the floor, and either down on it or - a real, statically stable outcome of ~1.5 % of these drops - hanging
    # upside down from the top of the pole, which stands on its flat weight box (hinge angles at +-pi)
    assert (qp[:, 2] > -0.01).all() and (qp[:, 2] < 1.2 * 1.1 + 0.2).all()
    perched = qp[:, 2] > 0.5
    assert perched.mean() < 0.05 and (np.abs(np.abs(qp[perched, 7:9]).max(axis=1) - np.pi) < 0.05).all()
    moving = np.abs(qv).max(axis=1) > 0.05
    assert moving.mean() < 0.01, moving.mean()                              # (a few may still be rocking)
    settled = np.abs(qv).max(axis=1) < 1e-3
    g = np.linalg.norm(sens, axis=1)
    print("settled", settled.mean(), "perched", perched.mean(), "max | |acc| - g | of the settled", np.abs(g[settled] - 9.81).max())
    assert settled.mean() > 0.9
    assert np.percentile(np.abs(g[settled] - 9.81), 99.9) < 0.05            # the floor carries the weight
    st = env.episode_stats()
    assert st["n_nonfinite"] == 0
    env.close()


def test_before_the_first_reset_the_drones_sit_on_the_spawn_grid():
    """SURVEY Q17 (env_gen.py:114-124): MjData starts at qpos0 = make_sim's spawn grid, 0.15 m above the floor; a vector_step issued
    there (RLlib never does: it resets first) sends the pendulum into the floor and truncates every drone at once."""
    import torch
    import mujoco_drone_b200 as M
    n = 70
    env = M.BaseDroneEnv(dict(M.base_config, num_drones=n, ground_contact=True))
    qp, qv, ac, sens, ns = env.get_state()
    sz = int(np.ceil(np.sqrt(n)))
    steps = (np.arange(sz) - (sz - 1) / 2) * 0.5
    xpos, ypos, zpos = np.meshgrid(steps, steps, [0.15])
    want = np.stack([xpos.flat[:n], ypos.flat[:n], zpos.flat[:n]], axis=1)
    assert np.abs(qp[:, :3] - want).max() < 2e-6                   # FP32 offsets from start_pos z = 15
    assert np.array_equal(qp[:, 3:], np.tile([1, 0, 0, 0, 0, 0], (n, 1))) and not qv.any() and not ac.any()
    obs, rew, trunc = env.step_tensor(torch.rand((n, 4), device="cuda"))
    assert trunc.all().item() and torch.isfinite(obs).all().item() and torch.isfinite(rew).all().item()
    qp2 = env.get_state()[0]
    assert (qp2[:, 2] > 0.15).all()                                 # the floor pushed them up (the weight starts 1.1 m under it)
    env.close()
