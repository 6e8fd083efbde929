"""§8 f-1 against the REFERENCE's own classes: tests/golden/policy_ref.npz holds a state_dict, inputs and the outputs of
RMA_full (models/PPO/RMA/RMA_model.py:19-136) and MyBetaDist (distributions.py:6-38), executed unmodified in the build
container (tools/make_golden_policy.py; RLlib's SlimFC / TorchModelV2 / TorchBeta are stubbed there, ray is absent).

CPU: the torch mirror (policy.make_rma_full + load_reference_state_dict) and the oracle's MyBetaDist against those outputs.
GPU: the fused tcgen05 kernel and the Beta kernel against the same outputs.
Tolerances: torch mirror 1e-5; oracle Beta (FP64) 1e-4 on logp (torch computes lgamma in FP32), 1e-6 on the mean;
fused kernel, bf16 operands / FP32 accumulate: logits and value <= 3e-2 (measured ~1e-2: bf16 has 8 bits of mantissa and
the reference's policy is FP32 - stated, not hidden); Beta kernel FP32: logp 2e-3 (1 + |logp|), mean 1e-5.
"""
import numpy as np
import pytest
import torch

from conftest import golden, has_cuda


def _model():
    import mujoco_drone_b200 as M
    g = golden("policy_ref.npz")
    m = M.policy.make_rma_full()
    M.policy.load_reference_state_dict(m, {k[3:]: g[k] for k in g.files if k.startswith("sd_")})
    return m.eval(), g


def test_torch_mirror_equals_reference_class():
    m, g = _model()
    with torch.no_grad():
        logits, value = m(torch.from_numpy(g["obs"]), torch.from_numpy(g["prev_action"]))
    assert np.abs(logits.numpy() - g["logits"]).max() <= 1e-5
    assert np.abs(value.numpy() - g["value"]).max() <= 1e-5
    with pytest.raises(KeyError):
        import mujoco_drone_b200 as M
        M.policy.load_reference_state_dict(M.policy.make_rma_full(), {"bogus.weight": np.zeros(3)})


def test_oracle_beta_matches_reference_mybetadist(oracle):
    g = golden("policy_ref.npz")
    det, _ = oracle.beta_policy(g["logits"].astype(np.float64), seed=1, env0=0, step=0, deterministic=True)
    assert np.abs(det - g["beta_det"]).max() <= 1e-6
    # logp of given actions: the oracle's logp is evaluated at its own sample, so rebuild log Beta(x; a, b) from the
    # oracle's alpha / beta definition (clamp +-50, softplus + 1) and compare with the reference's clamp-to-[0.01, 0.99] logp
    from scipy.special import betaln
    x = np.clip(g["beta_x"].astype(np.float64), 1e-2, 1 - 1e-2)
    lg = np.clip(g["logits"].astype(np.float64), -50, 50)
    ab = np.log(np.exp(lg) + 1.0) + 1.0
    a, b = ab[:, :4], ab[:, 4:]
    np.testing.assert_allclose(a, g["beta_alpha"], rtol=2e-6)
    np.testing.assert_allclose(b, g["beta_beta"], rtol=2e-6)
    logp = ((a - 1) * np.log(x) + (b - 1) * np.log1p(-x) - betaln(a, b)).sum(1)
    assert np.abs(logp - g["beta_logp"]).max() <= 1e-4 * (1 + np.abs(g["beta_logp"]).max())
    # the oracle's own logp at its deterministic sample (= the mean, clamped) against the same closed form
    act, lp = oracle.beta_policy(g["logits"].astype(np.float64), seed=1, env0=0, step=0, deterministic=True)
    xm = np.clip(act, 1e-2, 1 - 1e-2)
    want = ((a - 1) * np.log(xm) + (b - 1) * np.log1p(-xm) - betaln(a, b)).sum(1)
    assert np.abs(lp - want).max() <= 1e-9 * (1 + np.abs(want).max())


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
def test_fused_kernel_matches_reference_class_outputs():
    import mujoco_drone_b200 as M
    m, g = _model()
    fused = M.policy.FusedRMAFull(m.cuda(), device=0)
    obs, prev = torch.from_numpy(g["obs"]).cuda(), torch.from_numpy(g["prev_action"]).cuda()
    logits, value = fused(obs, prev)
    fused.check()
    el = np.abs(logits.cpu().numpy() - g["logits"]).max()
    ev = np.abs(value.cpu().numpy() - g["value"]).max()
    print(f"fused tcgen05 RMA_full vs the reference class: max |d logits| {el:.2e}, max |d value| {ev:.2e} (bf16 operands)")
    assert el <= 3e-2 and ev <= 3e-2, (el, ev)
    # reset mask: rows flagged as episode starts see a zero previous action, exactly like the fixture rows with prev = 0
    mask = torch.zeros(len(g["obs"]), dtype=torch.uint8, device="cuda")
    mask[::9] = 1
    lg2, _ = fused(obs, torch.rand_like(prev) * mask[:, None] + prev * (1 - mask[:, None]), reset_mask=mask)
    assert torch.equal(lg2[::9], logits[::9])
    fused.close()


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
def test_beta_kernel_matches_reference_mybetadist():
    import mujoco_drone_b200 as M
    from scipy.special import betaln
    g = golden("policy_ref.npz")
    lg = torch.from_numpy(g["logits"]).cuda()
    det, lp = M.policy.beta_policy(lg, seed=3, deterministic=True)
    assert np.abs(det.cpu().numpy() - g["beta_det"]).max() <= 1e-5
    a, b = g["beta_alpha"].astype(np.float64), g["beta_beta"].astype(np.float64)
    xm = np.clip(det.cpu().numpy().astype(np.float64), 1e-2, 1 - 1e-2)
    want = ((a - 1) * np.log(xm) + (b - 1) * np.log1p(-xm) - betaln(a, b)).sum(1)
    assert (np.abs(lp.cpu().numpy() - want) <= 2e-3 * (1 + np.abs(want))).all()


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
@pytest.mark.parametrize("n", [512, 1, 131])
def test_fp32_fused_kernel_matches_reference_class_outputs(n):
    """the FP32-faithful fused kernel (csrc/dsim_policy_fp32.cu) against the reference class's own outputs: <= 1e-4 on logits and
    value (measured ~1e-6: FP32 operands, FP32 accumulation, libm tanh; only the summation order differs from torch)"""
    import mujoco_drone_b200 as M
    m, g = _model()
    net = M.policy.FP32RMAFull(m, device=0)
    obs, prev = torch.from_numpy(g["obs"][:n]).cuda(), torch.from_numpy(g["prev_action"][:n]).cuda()
    logits, value = net(obs, prev)
    el = np.abs(logits.cpu().numpy() - g["logits"][:n]).max()
    ev = np.abs(value.cpu().numpy() - g["value"][:n]).max()
    print(f"FP32 fused RMA_full vs the reference class: max |d logits| {el:.2e}, max |d value| {ev:.2e}")
    assert el <= 1e-4 and ev <= 1e-4, (el, ev)
    mask = torch.zeros(n, dtype=torch.uint8, device="cuda")
    mask[::9] = 1
    lg2, _ = net(obs, torch.rand_like(prev) * mask[:, None] + prev * (1 - mask[:, None]), reset_mask=mask)
    assert torch.equal(lg2[::9], logits[::9])
    net.close()


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
def test_rollout_runner_with_fp32_fused_policy_equals_torch_fp32_policy():
    """RolloutRunner(policy_dtype='fused_fp32') against the library-GEMM FP32 policy on the same seeds: same actions up to FP32
    summation order (the Beta sampler is a continuous function of the logits except at rejection boundaries)"""
    import mujoco_drone_b200 as M
    n, T = 300, 6
    pol = M.policy.make_rma_full()
    out = {}
    for mode in ("fp32", "fused_fp32"):
        cfg = dict(M.base_config, num_drones=n, auto_reset=True, max_steps=7, max_distance=2.0, param_difficulty=1.0,
                   reward_fcn=M.rewards.distance_energy_reward, seed=3)
        env = M.observation_wrappers.LocalFrameRPYParamsEnv(cfg)
        r = M.rollout.RolloutRunner(env, pol, horizon=T, seed=9, use_graph=(mode == "fused_fp32"), policy_dtype=mode)
        if mode == "fp32":
            r.warm_up()
        b = r.run()
        out[mode] = {k: v.clone() for k, v in b.items()}
        env.close()
    same = (out["fp32"]["actions"][0] - out["fused_fp32"]["actions"][0]).abs() < 1e-4
    assert same.float().mean() > 0.99                                  # first step: identical inputs
    assert (out["fp32"]["values"][0] - out["fused_fp32"]["values"][0]).abs().max() < 1e-4
