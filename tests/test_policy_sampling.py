"""MyBetaDist (reference distributions.py:6-38) — the action-sampling step on the caller side of the env path.

CPU: the oracle's sampler is pinned against the published definition: torch.distributions.Beta (what RLlib's TorchBeta
wraps) for log-probabilities / means, scipy.stats.beta for the sample distribution.
GPU: the CUDA kernel (dsim_beta_policy, through the C ABI) against the oracle on the same Philox stream."""
import ctypes as C

import numpy as np
import pytest

from conftest import has_cuda


def _ref_alpha_beta(x):
    import torch
    t = torch.clamp(torch.as_tensor(x, dtype=torch.float64), -50, 50)          # distributions.py:12-16
    t = torch.log(torch.exp(t) + 1.0) + 1.0
    a, b = torch.chunk(t, 2, dim=-1)
    return a, b


def test_oracle_logp_and_mean_match_torch_beta(oracle):
    import torch
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(size=(500, 8)) * 3, rng.uniform(-80, 80, size=(100, 8))])
    a, b = _ref_alpha_beta(x)
    dist = torch.distributions.Beta(concentration1=a, concentration0=b)
    act, lp = oracle.beta_policy(x, seed=7, env0=0, step=3)
    assert ((act > 0) & (act < 1)).all()
    ref_lp = dist.log_prob(torch.clamp(torch.as_tensor(act), 1e-2, 1 - 1e-2)).sum(-1).numpy()   # MyBetaDist.logp (:19-22)
    np.testing.assert_allclose(lp, ref_lp, rtol=1e-10, atol=1e-9)
    mean, _ = oracle.beta_policy(x, seed=7, env0=0, step=3, deterministic=True)
    np.testing.assert_allclose(mean, dist.mean.numpy(), rtol=1e-12)                          # deterministic_sample (:24-26)


def test_oracle_sample_distribution_is_beta(oracle):
    from scipy import stats
    for xa, xb in [(-50.0, -50.0), (0.0, 1.5), (3.0, -2.0), (6.0, 6.0)]:
        n = 40000
        x = np.tile(np.array([xa] * 4 + [xb] * 4), (n, 1))
        a, b = _ref_alpha_beta(x[:1])
        act, _ = oracle.beta_policy(x, seed=11, env0=1000, step=0)
        ks = stats.kstest(act[:, 0], stats.beta(float(a[0, 0]), float(b[0, 0])).cdf)
        assert ks.pvalue > 1e-3, (xa, xb, ks)
        assert abs(np.corrcoef(act[:, 0], act[:, 1])[0, 1]) < 0.02           # independent action dimensions
    a1, _ = oracle.beta_policy(x, seed=11, env0=1000, step=0)
    a2, _ = oracle.beta_policy(x, seed=11, env0=1000, step=1)
    a3, _ = oracle.beta_policy(x[:10], seed=11, env0=1005, step=0)
    assert not np.array_equal(a1, a2)                                        # a new step draws new numbers
    np.testing.assert_array_equal(a3[:5], a1[5:10])                          # streams are keyed by the GLOBAL env id


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_kernel_matches_oracle(oracle, precision):
    import torch
    import mujoco_drone_b200 as M
    rng = np.random.default_rng(3)
    n = 5003
    x = np.concatenate([rng.normal(size=(n - 200, 8)) * 2.5, rng.uniform(-70, 70, size=(200, 8))])
    dt = torch.float64 if precision == "fp64" else torch.float32
    logits = torch.as_tensor(x, device="cuda", dtype=dt)
    x_dev = logits.cpu().numpy().astype(np.float64)
    for det in (False, True):
        act, lp = M.policy.beta_policy(logits, seed=5, env_id_offset=123, step=9, deterministic=det)
        oa, olp = oracle.beta_policy(x_dev, seed=5, env0=123, step=9, deterministic=det)
        a, l = act.cpu().numpy().astype(np.float64), lp.cpu().numpy().astype(np.float64)
        assert ((a > 0) & (a < 1)).all()
        if precision == "fp64":
            np.testing.assert_allclose(a, oa, rtol=1e-11, atol=1e-13)
            np.testing.assert_allclose(l, olp, rtol=1e-9, atol=1e-9)
        else:
            # FP32: identical accept/reject decisions except on rounding-level ties of the squeeze test
            close = np.abs(a - oa) <= 2e-5 * (1 + np.abs(oa))
            assert close.mean() > 0.998, close.mean()
            rows = close.all(axis=1)
            assert (np.abs(l[rows] - olp[rows]) <= 2e-3 * (1 + np.abs(olp[rows]))).all()


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
def test_kernel_sample_distribution_and_torch_logp():
    import torch
    from scipy import stats
    import mujoco_drone_b200 as M
    n = 200000
    x = torch.tensor([0.3, -1.0, 2.0, 5.0, 1.0, 0.0, -3.0, 5.0], device="cuda").repeat(n, 1)
    act, lp = M.policy.beta_policy(x, seed=1, env_id_offset=0, step=0)
    a, b = _ref_alpha_beta(x[:1].cpu().numpy())
    a_np = act.cpu().numpy()
    for k in range(4):
        ks = stats.kstest(a_np[:50000, k], stats.beta(float(a[0, k]), float(b[0, k])).cdf)
        assert ks.pvalue > 1e-3, (k, ks)
    dist = torch.distributions.Beta(a.to("cuda").float(), b.to("cuda").float())
    ref = dist.log_prob(torch.clamp(act, 1e-2, 1 - 1e-2)).sum(-1)
    assert (lp - ref).abs().max() < 2e-3
    with pytest.raises(ValueError):
        M.policy.beta_policy(x[:, :6], seed=1, env_id_offset=0, step=0)


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
def test_rollout_runner_graph_equals_eager_and_replays_through_the_env():
    """the GPU-resident rollout loop: (i) CUDA-graph replay == eager stepping, (ii) the recorded (obs, action) pairs are
    exactly what the env produces when the recorded actions are replayed step by step, (iii) prev_actions protocol"""
    import torch
    import mujoco_drone_b200 as M
    n, T = 300, 12

    def mk():
        cfg = dict(M.base_config, num_drones=n, auto_reset=True, max_steps=7, max_distance=2.0, param_difficulty=1.0,
                   reward_fcn=M.rewards.distance_energy_reward, seed=3)
        return M.observation_wrappers.LocalFrameRPYParamsEnv(cfg)
    pol = M.policy.make_rma_full()
    out = {}
    for mode in (False, True):
        env = mk()
        r = M.rollout.RolloutRunner(env, pol, horizon=T, seed=9, use_graph=mode)
        if not mode:
            r.warm_up()                   # the graph capture runs the same 3 unrecorded steps first
        b = r.run()
        assert env.total_steps == T + r.WARMUP_STEPS
        out[mode] = {k: v.clone() for k, v in b.items()}
        assert b["obs"].shape == (T + 1, n, 22) and b["actions"].shape == (T, n, 4) and b["truncated"].dtype == torch.uint8
        assert ((b["actions"] > 0) & (b["actions"] < 1)).all() and torch.isfinite(b["values"]).all() and torch.isfinite(b["action_logp"]).all()
        assert b["truncated"].sum() > 0
        ds = r.to_reference_dataset(b)
        assert ds["o"].shape == (n, T, 22) and ds["a"].shape == (n, T, 4) and ds["t"].shape == (n, T) and ds["z"].shape == (n, 6)
        # (ii) replay the recorded actions through a fresh env from the same reset
        env2 = mk()
        if mode:
            continue                      # the replay check runs on the eager rollout
        r2 = M.rollout.RolloutRunner(env2, pol, horizon=T, seed=9, use_graph=False)
        r2.warm_up()                      # same start as the recorded rollout
        assert torch.equal(r2._obs_cur, b["obs"][0])
        for t in range(T):
            o, rew, tr = env2.step_tensor(b["actions"][t])
            assert torch.equal(o, b["obs"][t + 1]) and torch.equal(rew, b["rewards"][t]) and torch.equal(tr, b["truncated"][t])
        env2.close()
        env.close()
    # (i) graph replay == eager stepping, bit for bit (same policy, seeds and warm-up steps)
    for k in out[False]:
        assert torch.equal(out[True][k], out[False][k]), k


def test_umma_weight_packing_layout():
    """host packer: element (n, k) of a [N, K] matrix lands at ((k // 8) * N + n) * 8 + k % 8 (UMMA K-major canonical layout)"""
    import torch
    import mujoco_drone_b200 as M
    w = torch.arange(16 * 32, dtype=torch.float32).reshape(16, 32) / 8.0        # exactly representable in bf16? use small ints
    w = torch.arange(16 * 32, dtype=torch.float32).reshape(16, 32) % 251
    p = M.policy._pack_umma_kmajor(w, 16, 32).view(torch.bfloat16).float()
    for n, k in [(0, 0), (3, 7), (5, 8), (15, 31), (9, 20)]:
        assert p[((k // 8) * 16 + n) * 8 + k % 8].item() == w[n, k].item()
    blob, c = M.policy.pack_rma_full(M.policy.make_rma_full())
    assert blob.numel() == 92160 and c.numel() == 1408


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
@pytest.mark.parametrize("schedule", ["pingpong", "shared_epilogue"])
def test_fused_tcgen05_policy_matches_torch_fp32(schedule, monkeypatch):
    """the fused tcgen05 RMA_full kernel vs the plain PyTorch FP32 module (same weights, BatchNorm with non-trivial running
    statistics).  Tolerance: bf16 operands with FP32 accumulation through 5 layers -> 3e-2 absolute on O(1) logits.
    Both schedules of the kernel (default ping-pong; DSIM_MLP_V2=1: epilogues shared by all warps + a dedicated MMA warp)."""
    import torch
    import mujoco_drone_b200 as M
    monkeypatch.setenv("DSIM_MLP_V2", "1" if schedule == "shared_epilogue" else "0")
    torch.manual_seed(0)
    model = M.policy.make_rma_full().cuda()
    bn = model.hidden[4]
    with torch.no_grad():
        bn.running_mean.copy_(torch.randn(128, device="cuda") * 0.2)
        bn.running_var.copy_(torch.rand(128, device="cuda") + 0.5)
        bn.weight.copy_(torch.rand(128, device="cuda") + 0.5)
        bn.bias.copy_(torch.randn(128, device="cuda") * 0.1)
        for lin in (model.logits[2], model.value_branch[4]):
            lin.bias.copy_(torch.randn_like(lin.bias) * 0.3)
    fused = M.policy.FusedRMAFull(model, device=0)
    for n in (1, 127, 128, 300, 4096 + 37, 150000):
        obs = torch.randn((n, 22), device="cuda")
        obs[:, 16:] = torch.tensor([1.0, 0.17, 7.0, 0.01, 1.2, 0.3], device="cuda") * (1 + 0.1 * torch.randn((n, 6), device="cuda"))
        prev = torch.rand((n, 4), device="cuda")
        with torch.no_grad():
            ref_l, ref_v = model(obs, prev)
        lg, val = fused(obs, prev)
        fused.check()
        assert torch.isfinite(lg).all() and torch.isfinite(val).all()
        assert (lg - ref_l).abs().max().item() < 3e-2, (n, (lg - ref_l).abs().max().item())
        assert (val - ref_v).abs().max().item() < 3e-2, (n, (val - ref_v).abs().max().item())
    fused.close()


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
@pytest.mark.parametrize("deterministic", [False, True])
def test_fused_policy_sampling_equals_separate_sampling_kernel(deterministic):
    """dsim_policy_forward_sample == dsim_policy_forward followed by dsim_beta_policy on its logits: same Philox streams
    (seed, global env id, step, device step counter), same device function -> same actions and log-probabilities; also
    in place (actions written over the previous-action rows) and with the reset mask."""
    import torch
    import mujoco_drone_b200 as M
    torch.manual_seed(3)
    model = M.policy.make_rma_full().cuda()
    with torch.no_grad():
        model.logits[2].bias.copy_(torch.randn(8, device="cuda"))
        model.logits[2].weight.mul_(4.0)
    fused = M.policy.FusedRMAFull(model, device=0)
    ctr = torch.full((1,), 5, dtype=torch.int32, device="cuda")
    for n in (1, 200, 128 * 301 + 5):
        obs = torch.randn((n, 22), device="cuda")
        prev = torch.rand((n, 4), device="cuda")
        mask = (torch.rand((n,), device="cuda") < 0.3).to(torch.uint8)
        lg, val = fused(obs, prev, reset_mask=mask)
        act_ref, lp_ref = M.policy.beta_policy(lg, seed=11, env_id_offset=1000, step=7, deterministic=deterministic, step_tensor=ctr)
        act, lp, val2, lg2 = fused.sample(obs, prev, seed=11, env_id_offset=1000, step=7, deterministic=deterministic, reset_mask=mask,
                                          step_tensor=ctr, want_logits=True)
        fused.check()
        assert torch.equal(lg, lg2) and torch.equal(val, val2)
        assert (act - act_ref).abs().max().item() < 1e-6 and (lp - lp_ref).abs().max().item() < 1e-4
        inplace = prev.clone()
        act3, lp3, _, none = fused.sample(obs, inplace, seed=11, env_id_offset=1000, step=7, deterministic=deterministic, reset_mask=mask,
                                          step_tensor=ctr, actions_out=inplace)
        assert none is None and act3.data_ptr() == inplace.data_ptr()
        assert torch.equal(act3, act) and torch.equal(lp3, lp)
    fused.close()


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
@pytest.mark.parametrize("fuse_sampling", [False, True])
def test_rollout_runner_with_fused_policy_kernel(fuse_sampling):
    """config-5 loop with the hand-written tcgen05 policy kernel, CUDA-graph replayed: finite, in-range, and its logits
    agree with the torch FP32 module evaluated on the recorded observations / previous actions"""
    import torch
    import mujoco_drone_b200 as M
    n, T = 1000, 6
    cfg = dict(M.base_config, num_drones=n, auto_reset=True, max_steps=64, param_difficulty=1.0, reward_fcn=M.rewards.distance_energy_reward)
    env = M.observation_wrappers.LocalFrameRPYParamsEnv(cfg)
    pol = M.policy.make_rma_full()
    r = M.rollout.RolloutRunner(env, pol, horizon=T, seed=1, policy_dtype="fused", use_graph=True, fuse_sampling=fuse_sampling)
    b = r.run()
    r._fused.check()
    assert ((b["actions"] > 0) & (b["actions"] < 1)).all() and torch.isfinite(b["values"]).all() and torch.isfinite(b["action_logp"]).all()
    with torch.no_grad():
        prev = torch.zeros((n, 4), device="cuda")
        for t in range(T):
            _, v = pol(b["obs"][t], prev)
            assert (v - b["values"][t]).abs().max().item() < 3e-2
            prev = b["actions"][t] * (b["truncated"][t] == 0).float().unsqueeze(1)
    env.close()


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
@pytest.mark.parametrize("mode", ["fp32", "fused"])
def test_rollout_history_windows(mode):
    """32-step (state, previous action) windows (RMA_model.py:41-43): rebuilt on the host from the recorded rollout"""
    import torch
    import mujoco_drone_b200 as M
    n, T, L = 200, 14, 6
    cfg = dict(M.base_config, num_drones=n, auto_reset=True, max_steps=5, param_difficulty=1.0, reward_fcn=M.rewards.distance_energy_reward)
    env = M.observation_wrappers.LocalFrameRPYParamsEnv(cfg)
    r = M.rollout.RolloutRunner(env, M.policy.make_rma_full(), horizon=T, seed=2, policy_dtype=mode, use_graph=False, history_len=L)
    b = r.run()
    h = r.history().cpu().numpy()
    obs, act, tr = b["obs"].cpu().numpy(), b["actions"].cpu().numpy(), b["truncated"].cpu().numpy().astype(bool)
    exp = np.zeros((n, L, 20), dtype=np.float32)
    for i in range(n):
        rows = []
        for t in range(T):                                     # row t = (obs[t][:16], action[t-1] or 0 at an episode start)
            start = t == 0 or tr[t - 1, i]
            prev = np.zeros(4, np.float32) if start else act[t - 1, i]
            if start:
                rows = []
            rows.append(np.concatenate([obs[t, i, :16], prev]))
        rows = rows[-L:]
        exp[i, L - len(rows):] = np.array(rows)
    np.testing.assert_allclose(h, exp, atol=1e-6)
    env.close()
