"""Round-2 GPU parity tests (through the C ABI): rows of SURVEY.md §8 that had no checker of their own in round 1.

  a-11  control_reference kernel vs the reference's own outputs (tests/golden/control_reference.npz: dead-zone edges, yaw
        wrap, +-(5,5,6) clip) - dead-zone decisions exact in both precisions, setpoints <= 1e-12 (FP64) / 2e-6 (FP32)
  a-1   the C3-specialised step kernel (per-env setpoints) vs orc_vector_step(per_env_ref=1)
  a-7   gimbal-lock specials of transform.npz through dsim_set_state -> dsim_compute_states ON THE GPU
  e     one handle of 2N envs == two handles of N envs at env_id_offset 0 / N, bit for bit (multi-GPU determinism)
  b     several step-kernel instantiations from one handle (step -> evaluate; set_params uniform <-> per-env)
"""
import numpy as np
import pytest

from conftest import golden, has_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]

NOMINAL = np.array([1, 0.17, 7, 0.01, 1.2, 0.3])
PKEYS = ("mass", "arm_len", "motor_force", "motor_tau", "pendulum_len", "weight_mass")


def _mk(cls_name="BaseDroneEnv", **over):
    import mujoco_drone_b200 as M
    cls = M.BaseDroneEnv if cls_name == "BaseDroneEnv" else getattr(M.observation_wrappers, cls_name)
    cfg = dict(M.base_config)
    cfg.update(over)
    return cls(cfg)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_control_reference_matches_reference_outputs(precision):
    """every joystick sequence of the fixture runs in its own env column: env k follows sequence k sample by sample"""
    import torch
    g = golden("control_reference.npz")
    start, axes, want, sid = g["start"], g["axes"], g["reference"], g["seq_id"]
    nseq = int(sid.max()) + 1
    seqs = [axes[sid == k] for k in range(nseq)]
    refs = [want[sid == k] for k in range(nseq)]
    T = max(len(s) for s in seqs)
    env = _mk("LocalFrameRPYEnv", num_drones=nseq, precision=precision, per_env_reference=True, reference=list(start), start_pos=list(start))
    dt = torch.float64 if precision == "fp64" else torch.float32
    tol = 1e-12 if precision == "fp64" else 2e-6
    import mujoco_drone_b200 as M
    prev = np.tile(start, (nseq, 1))
    prev_dev = prev.copy()
    for t in range(T):
        a = np.zeros((4, nseq))
        for k in range(nseq):
            if t < len(seqs[k]):
                a[:, k] = seqs[k][t]
        env.control_reference_tensor(torch.as_tensor(a, dtype=dt, device="cuda"))
        r = env.rows(M._lib.BUF_REFERENCE)[:, :nseq].cpu().numpy().astype(np.float64).T       # offsets from start_pos
        r[:, :3] += start[:3]
        for k in range(nseq):
            exp = refs[k][t] if t < len(seqs[k]) else prev[k]
            assert np.abs(r[k] - exp).max() <= tol * (1 + np.abs(exp).max()), (precision, k, t, r[k], exp)
            # the dead-zone decision itself: a setpoint either moved or it did not, exactly like the reference's
            moved_ref = np.abs(exp - prev[k]) > 0
            moved = np.abs(r[k] - prev_dev[k]) > (0 if precision == "fp64" else 1e-5)     # smallest real move: 0.1 * 0.01
            clipped = (np.abs(exp[:3] - start[:3]) >= np.array([5, 5, 6]) - 1e-9)
            assert (moved[:3] == moved_ref[:3])[~clipped].all(), (precision, k, t)
            prev[k], prev_dev[k] = exp, r[k]
    env.close()


def test_c3_specialised_kernel_matches_oracle_vector_step(oracle):
    """BASELINE config 3 instantiation step_kernel<float, true, LOCAL_RPY, distance_reward, per-env setpoints> against the
    oracle's whole vector_step with per-env references (obs, reward, truncated, state), FP32 tolerances of this file's header"""
    import torch
    import mujoco_drone_b200 as M
    rng = np.random.default_rng(21)
    n = 2048 + 17
    env = _mk("LocalFrameRPYEnv", num_drones=n, per_env_reference=True, reward_fcn=M.rewards.distance_reward_fcn, random_params=False,
              state_difficulty=0.4, max_steps=512, max_distance=4)
    env.reset_tensor()
    # distinct per-env setpoints, then two joystick updates
    axes = np.round(rng.uniform(-1, 1, size=(4, n)), 2)
    for _ in range(3):
        env.control_reference_tensor(torch.as_tensor(axes, dtype=torch.float32, device="cuda"))
    refs = env.rows(M._lib.BUF_REFERENCE)[:, :n].cpu().numpy().astype(np.float64).T.copy()
    refs[:, :3] += np.array([0, 0, 15.0])
    qpos, qvel, act, _, ns0 = env.get_state()
    cpu = oracle.CpuVecEnv(np.tile(NOMINAL, (n, 1)), True, 100.0, 1, True)
    cpu.qpos[:], cpu.qvel[:], cpu.act[:] = qpos, qvel, act
    cpu.num_steps[:] = ns0
    a = rng.uniform(0, 1, size=(n, 4)).astype(np.float32)
    obs, rew, trunc = env.step_tensor(torch.as_tensor(a, device="cuda"))
    oobs, orew, otr = cpu.step(a.astype(np.float64), refs, oracle.REWARD_IDS["distance_reward_fcn"], oracle.OBS_IDS["LocalFrameRPYEnv"], 4.0, 512)
    obs, rew, trunc = obs.cpu().numpy().astype(np.float64), rew.cpu().numpy().astype(np.float64), trunc.cpu().numpy().astype(bool)
    qp, qv, _, _, ns = env.get_state()
    assert np.abs(qp[:, :3] - cpu.qpos[:, :3]).max() <= 2e-6 and np.abs(qp[:, 3:] - cpu.qpos[:, 3:]).max() <= 2e-6
    assert (np.abs(qv - cpu.qvel) <= 1e-4 * (1 + np.abs(cpu.qvel))).all()
    assert (np.abs(obs - oobs) <= 2e-4 * (1 + np.abs(oobs))).all(), np.abs(obs - oobs).max()
    assert (np.abs(rew - orew) <= 2e-4 * (1 + np.abs(orew))).all()
    near = np.abs(np.linalg.norm(cpu.qpos[:, :3] - refs[:, :3], axis=1) - 4.0) < 1e-5       # rounding-level ties only
    assert (trunc == otr)[~near].all() and (ns == cpu.num_steps).all()
    env.close()


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_gimbal_lock_specials_on_the_gpu(precision):
    """transform.npz quaternions (random, un-normalised, and the pitch = +-pi/2 specials) through dsim_set_state ->
    dsim_compute_states: rows 3:6 of get_drone_states are mujoco_quat2rpy (transformation.py:5-8).  In gimbal lock only
    yaw -/+ roll is defined: scipy pins the third angle to 0; FP32 may land a hair outside the 1e-7 lock window, so the
    specials are compared through the rotation matrix they encode; generic quaternions angle by angle."""
    from oracle import oracle as O
    g = golden("transform.npz")
    quats, want, nspec = g["quats"], g["quat2rpy"], int(g["n_special"])
    n = len(quats)
    env = _mk(num_drones=n, precision=precision, random_params=False)
    qpos = np.zeros((n, 9))
    qpos[:, 2] = 15.0
    qpos[:, 3:7] = quats
    env.set_state(qpos, np.zeros((n, 8)))
    st = np.array(env.get_drone_states())
    rpy = st[:, 3:6]
    tol = 1e-9 if precision == "fp64" else 2e-5
    generic = np.ones(n, dtype=bool)
    generic[n - nspec:] = False
    generic &= np.abs(np.abs(want[:, 1]) - np.pi / 2) > 1e-3
    d = np.abs((rpy - want + np.pi) % (2 * np.pi) - np.pi)
    assert d[generic].max() <= tol, d[generic].max()
    if precision == "fp64":                                                   # same algorithm, same branches: everything matches
        assert d.max() <= 1e-9, d.max()
    for i in np.where(~generic)[0]:
        R1 = O.quat2dcm(O.rpy2quat(rpy[i]))
        R2 = O.quat2dcm(O.rpy2quat(want[i]))
        assert np.abs(R1 - R2).max() <= (1e-9 if precision == "fp64" else 5e-4), (i, rpy[i], want[i])   # sqrt(eps_fp32) at the pole
    env.close()


def test_one_handle_of_2n_equals_two_handles_with_id_offsets():
    """Philox streams are keyed by the GLOBAL env id, so a shard boundary changes nothing: states, observations,
    rewards, truncation flags and drawn parameters of envs [0, 2N) in one handle are bit-identical to handles [0, N) + [N, 2N)"""
    import torch
    import mujoco_drone_b200 as M
    N = 1024 + 32
    kw = dict(state_difficulty=0.4, param_difficulty=1.0, random_params=True, max_steps=9, max_distance=1.5, auto_reset=True,
              reward_fcn=M.rewards.distance_energy_reward, seed=5)
    whole = _mk("LocalFrameRPYParamsEnv", num_drones=2 * N, env_id_offset=1000, **kw)
    parts = [_mk("LocalFrameRPYParamsEnv", num_drones=N, env_id_offset=1000 + k * N, **kw) for k in range(2)]
    o = whole.reset_tensor()
    po = [p.reset_tensor() for p in parts]
    assert torch.equal(o, torch.cat(po))
    g = torch.Generator(device="cuda").manual_seed(17)
    for t in range(25):                                       # several episode ends and in-kernel resets per env
        a = torch.rand((2 * N, 4), device="cuda", generator=g)
        ow, rw, tw = whole.step_tensor(a)
        outs = [p.step_tensor(a[k * N:(k + 1) * N].contiguous()) for k, p in enumerate(parts)]
        assert torch.equal(ow, torch.cat([x[0] for x in outs])), t
        assert torch.equal(rw, torch.cat([x[1] for x in outs])) and torch.equal(tw, torch.cat([x[2] for x in outs])), t
    sw = whole.get_state()
    sp = [p.get_state() for p in parts]
    for j in range(5):
        assert np.array_equal(sw[j], np.concatenate([s[j] for s in sp]))
    pw = np.array([list(d.values()) for d in whole.drone_params])
    pp = np.concatenate([np.array([list(d.values()) for d in p.drone_params]) for p in parts])
    assert np.array_equal(pw, pp)
    assert whole.episode_stats()["n_episodes"] == sum(p.episode_stats()["n_episodes"] for p in parts) > 0
    whole.close()
    [p.close() for p in parts]


def test_fresh_process_step_then_evaluate_and_params_toggle():
    """the > 48 KB dynamic shared-memory opt-in is per kernel INSTANTIATION.  A fresh process that launches the
    specialised C4 kernel first and the generic one second (dsim_evaluate; a different CFG after set_params) must not
    fail with cudaErrorInvalidValue.  Runs in a subprocess so no earlier test has opted any kernel in."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
import mujoco_drone_b200 as M
n = 4096
cfg = dict(M.base_config, num_drones=n, reward_fcn=M.rewards.distance_energy_reward, param_difficulty=1.0, random_params=True, auto_reset=True)
env = M.observation_wrappers.LocalFrameRPYParamsEnv(cfg)
env.reset_tensor()
a = torch.rand((n, 4), device="cuda")
o1 = env.step_tensor(a)[0].clone()                      # specialised <float, true, RPY_PARAMS, 2, 5>
o2, r2, t2 = env.evaluate_tensor(a)                     # generic <float, true, -1, -1>
torch.cuda.synchronize()
assert torch.allclose(o1, o2, atol=1e-5), (o1 - o2).abs().max()
p = [dict(zip(("mass", "arm_len", "motor_force", "motor_tau", "pendulum_len", "weight_mass"), [1, .17, 7, .01, 1.2, .3]))] * n
env.drone_params = p                                    # uniform -> per_env_consts = 0 -> generic CFG
env.step_tensor(a)
q = [dict(d, mass=1.0 + 0.0001 * i) for i, d in enumerate(p)]
env.drone_params = q                                    # per-env again -> specialised
env.step_tensor(a)
torch.cuda.synchronize()
c = M.BaseDroneEnv(dict(M.base_config, num_drones=64, random_params=False))     # C2 instantiation, then its evaluate
c.reset_tensor(); c.step_tensor(a[:64].contiguous()); c.evaluate_tensor(a[:64].contiguous())
torch.cuda.synchronize()
print("OK")
''' % root
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


def test_step_host_rejects_wrong_output_buffers():
    env = _mk("LocalFrameRPYParamsEnv", num_drones=64)
    env.reset_tensor()
    a = np.random.rand(64, 4).astype(np.float32)
    with pytest.raises(ValueError):
        env.step_host(a, obs_out=np.empty((64, env.obs_dim), np.float64))
    with pytest.raises(ValueError):
        env.step_host(a, obs_out=np.empty((63, env.obs_dim), np.float32))
    with pytest.raises(ValueError):
        env.step_host(a, reward_out=np.empty(128, np.float32)[::2])
    with pytest.raises(ValueError):
        env.evaluate_tensor(__import__("torch").zeros((63, 4), device="cuda"))
    env.close()


@pytest.mark.parametrize("n", [4096 + 32, 131072])
def test_inputs_ready_mode_is_bit_identical_to_strict_stepping(n):
    """dsim_set_inputs_ready: R shards interleaved on one stream (eagerly and as a replayed CUDA graph) prefetch their pages
    and run their first page's physics before the programmatic-dependency wait.  Everything they produce must equal strict
    stepping bit for bit - a stale early read of state written R launches ago would show up here (small grids are the
    dangerous case: several kernels fit on the GPU at once)."""
    import torch
    import mujoco_drone_b200 as M
    R = 3
    kw = dict(state_difficulty=0.4, param_difficulty=1.0, random_params=True, max_steps=11, max_distance=1.5, auto_reset=True,
              reward_fcn=M.rewards.distance_energy_reward, seed=7)
    sets = []
    for ready in (False, True):
        envs = [_mk("LocalFrameRPYParamsEnv", num_drones=n, env_id_offset=r * n, inputs_ready=ready, **kw) for r in range(R)]
        for e in envs:
            assert e.inputs_ready == ready
            e.reset_tensor()
        sets.append(envs)
    g = torch.Generator(device="cuda").manual_seed(3)
    bank = torch.rand((5, n, 4), device="cuda", generator=g)

    def run(envs, k0, k):
        for i in range(k0, k0 + k):
            envs[i % R].step_tensor(bank[i % 5])
    for envs in sets:                                        # eager, back to back
        run(envs, 0, 4 * R)
    torch.cuda.synchronize()
    graphs = []
    for envs in sets:                                        # graph replay: kernels back to back with no host gaps
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run(envs, 0, R)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            run(envs, 0, 5 * R)
        graphs.append(gr)
    for _ in range(8):
        for gr in graphs:
            gr.replay()
    torch.cuda.synchronize()
    for a, b in zip(*sets):
        sa, sb = a.get_state(), b.get_state()
        assert all(np.array_equal(x, y) for x, y in zip(sa, sb))
        assert torch.equal(a.obs_tensor, b.obs_tensor) and torch.equal(a.reward_tensor, b.reward_tensor) and torch.equal(a.truncated_tensor, b.truncated_tensor)
        assert a.episode_stats()["n_episodes"] == b.episode_stats()["n_episodes"] > 0
    for envs in sets:
        [e.close() for e in envs]


def test_no_out_of_bounds_device_writes_canaries():
    """compute-sanitizer is closed on the pool: own check instead.  DSIM_GUARD=1 allocates every device buffer of a handle at
    its exact size between two 4 KB canary regions; tools/sanitize_smoke.py drives ragged sizes (1, 33, 65, 97, 1000, 4129,
    76007 envs) through every step-kernel instantiation, both dependency modes, in-kernel resets, evaluate, the host entry
    point and the auxiliary kernels, and asserts that no canary byte changed."""
    import os
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DSIM_GUARD="1")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "sanitize_smoke.py")], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    m = re.search(r"handles with intact canaries: (\d+)", r.stdout)
    assert m and int(m.group(1)) >= 40, r.stdout[-500:]


def test_two_envs_per_lane_kernel_matches_the_default_kernel():
    """csrc/dsim_step_x2.cuh (DSIM_X2=1: two envs per lane, physics on the packed FP32 forms) is measured slower than the default
    and stays opt-in; this keeps it honest: same C4 rollout, fresh processes, outputs within FP32 round-off of the default kernel
    (both are held to the oracle by the tolerance tests when selected: tools/gpu_x2.sh runs the parity suite through it)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
import mujoco_drone_b200 as M
n = 4128 + 32
cfg = dict(M.base_config, num_drones=n, reward_fcn=M.rewards.distance_energy_reward, auto_reset=True, param_difficulty=1.0, max_steps=40, seed=5)
env = M.observation_wrappers.LocalFrameRPYParamsEnv(cfg)
env.reset_tensor()
g = torch.Generator(device="cuda"); g.manual_seed(3)
tot = 0
for t in range(60):
    obs, rew, trunc = env.step_tensor(torch.rand((n, 4), device="cuda", generator=g))
    tot += int(trunc.sum())
np.save(sys.argv[1], np.concatenate([obs.cpu().numpy().ravel(), rew.cpu().numpy().ravel(), trunc.cpu().numpy().astype(np.float32).ravel(), [tot]]))
''' % root
    import tempfile
    import numpy as np
    outs = []
    with tempfile.TemporaryDirectory() as td:
        for x2 in ("0", "1"):
            f = os.path.join(td, f"o{x2}.npy")
            r = subprocess.run([sys.executable, "-c", code, f], capture_output=True, text=True, timeout=600, env=dict(os.environ, DSIM_X2=x2))
            assert r.returncode == 0, r.stderr[-2000:]
            outs.append(np.load(f))
    a, b = outs
    assert a[-1] == b[-1] and a[-1] > 0                           # same episodes ended (in-kernel resets exercised)
    assert np.isfinite(b).all() and np.abs(a - b).max() < 2e-3    # 60 chaotic steps apart by FP32 contraction order only
