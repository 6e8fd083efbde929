#!/usr/bin/env python
"""tests/golden/policy_ref.npz: outputs of the REFERENCE's own policy classes, executed in the build container.

`RMA_full` (models/PPO/RMA/RMA_model.py:19-136) and `MyBetaDist` (distributions.py:6-38) are imported UNMODIFIED from
/root/reference.  ray is not installed, so the four RLlib names they build on are replaced by stand-ins that restate
RLlib's documented behaviour (ray 2.x, rllib/models/torch/misc.py and torch_action_dist.py):
  TorchModelV2            stores the constructor arguments, `view_requirements = {}`
  SlimFC(in, out, initializer, activation_fn, bias_init=0.0)   `self._model = nn.Sequential(nn.Linear, [nn.Tanh])`, weight
                          through `initializer`, bias constant 0 -> state_dict keys `<name>._model.0.weight / .bias`
  normc_initializer(std)  N(0,1) weight, every output row scaled to norm `std`
  TorchBeta / TorchDistributionWrapper   `.inputs`, `.model`; `entropy()` / `kl()` of `self.dist`
Everything else (layer list, BatchNorm, the concatenation order of states / previous action / embedding, the clamp and
softplus of MyBetaDist, logp clamping) is the reference's code running as written.

Run through tools/make_golden_r2.py (which also installs the gymnasium stubs).
"""
import os
import sys
import types

import numpy as np


def install_rllib_policy_stubs():
    import torch
    import torch.nn as nn
    from oracle import ref_stubs
    ref_stubs.install_stubs()
    M = ref_stubs._mod

    class TorchModelV2:
        def __init__(self, obs_space, action_space, num_outputs, model_config, name):
            self.obs_space, self.action_space, self.num_outputs, self.model_config, self.name = obs_space, action_space, num_outputs, model_config, name
            self.view_requirements = {}

    class SlimFC(nn.Module):
        def __init__(self, in_size, out_size, initializer=None, activation_fn=None, use_bias=True, bias_init=0.0):
            super().__init__()
            layers = []
            linear = nn.Linear(in_size, out_size, bias=use_bias)
            if initializer is None:
                initializer = nn.init.xavier_uniform_
            initializer(linear.weight)
            if use_bias:
                nn.init.constant_(linear.bias, bias_init)
            layers.append(linear)
            if activation_fn == 'tanh':
                layers.append(nn.Tanh())
            elif activation_fn is not None:
                raise ValueError(activation_fn)
            self._model = nn.Sequential(*layers)

        def forward(self, x):
            return self._model(x)

    class AppendBiasLayer(nn.Module):
        pass

    def normc_initializer(std=1.0):
        def initializer(tensor):
            tensor.data.normal_(0, 1)
            tensor.data *= std / torch.sqrt(tensor.data.pow(2).sum(1, keepdim=True))
        return initializer

    class TorchDistributionWrapper:
        def __init__(self, inputs, model):
            if not isinstance(inputs, torch.Tensor):
                inputs = torch.from_numpy(inputs)
            self.inputs, self.model = inputs, model

        def entropy(self):
            return self.dist.entropy()

        def kl(self, other):
            return torch.distributions.kl.kl_divergence(self.dist, other.dist)

        def sample(self):
            self.last_sample = self.dist.sample()
            return self.last_sample

    class TorchBeta(TorchDistributionWrapper):
        pass

    class ViewRequirement:
        def __init__(self, data_col=None, shift=0, space=None, **kw):
            self.data_col, self.shift, self.space = data_col, shift, space

    class SampleBatch(dict):
        OBS, PREV_ACTIONS, ACTIONS = "obs", "prev_actions", "actions"

    class ModelCatalog:
        register_custom_model = staticmethod(lambda *a, **k: None)
        register_custom_action_dist = staticmethod(lambda *a, **k: None)

    M("ray.rllib.models.torch.torch_modelv2", TorchModelV2=TorchModelV2)
    M("ray.rllib.models.torch.misc", SlimFC=SlimFC, AppendBiasLayer=AppendBiasLayer, normc_initializer=normc_initializer)
    M("ray.rllib.models.torch.torch_action_dist", TorchBeta=TorchBeta, TorchDistributionWrapper=TorchDistributionWrapper)
    M("ray.rllib.models.torch", torch_modelv2=sys.modules["ray.rllib.models.torch.torch_modelv2"], misc=sys.modules["ray.rllib.models.torch.misc"])
    M("ray.rllib.models", ModelCatalog=ModelCatalog, torch=sys.modules["ray.rllib.models.torch"])
    M("ray.rllib.policy.view_requirement", ViewRequirement=ViewRequirement)
    M("ray.rllib.policy.sample_batch", SampleBatch=SampleBatch)
    M("ray.rllib.policy", view_requirement=sys.modules["ray.rllib.policy.view_requirement"], sample_batch=sys.modules["ray.rllib.policy.sample_batch"])
    M("ray.rllib.utils.annotations", override=lambda cls: (lambda f: f))
    M("ray.rllib.utils.typing", Dict=dict, TensorType=object, List=list, ModelConfigDict=dict)
    M("ray.rllib.utils", annotations=sys.modules["ray.rllib.utils.annotations"], typing=sys.modules["ray.rllib.utils.typing"])


def load_reference_policy_classes():
    install_rllib_policy_stubs()
    from oracle import ref_stubs
    if ref_stubs.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_stubs.REFERENCE_ROOT)
    import importlib
    rma = importlib.import_module("models.PPO.RMA.RMA_model")
    dists = importlib.import_module("distributions")
    return rma, dists


def gen_policy_ref(out_dir):
    import torch
    rma, dists = load_reference_policy_classes()
    torch.manual_seed(20261019)
    cfg = {"custom_model_config": {"num_states": 16, "num_params": 6, "num_actions": 4, "param_embed_dim": 8,
                                   "train_adaptation": False, "adapt_seq_len": 32}}          # train_RMA.py:48-54 with the adaptation phase off
    model = rma.RMA_full(None, None, 8, cfg, "rma_full")
    # a trained network has non-trivial biases and BatchNorm statistics: perturb them so the BatchNorm fold of the fused kernel is exercised
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("bias"):
                p.add_(0.1 * torch.randn_like(p))
        bn = model._hidden_layers[2]
        bn.running_mean.copy_(0.2 * torch.randn(128))
        bn.running_var.copy_(0.5 + torch.rand(128))
        bn.weight.copy_(1.0 + 0.2 * torch.randn(128))
        bn.bias.copy_(0.1 * torch.randn(128))
    n = 512
    rng = np.random.default_rng(5)
    obs = np.concatenate([rng.normal(size=(n, 16)) * np.array([1.5] * 3 + [0.5] * 3 + [1.5] * 3 + [2.0] * 3 + [0.5] * 2 + [1.0] * 2),
                          np.array([1, .17, 7, .01, 1.2, .3]) * rng.uniform(0.8, 1.2, size=(n, 6))], axis=1).astype(np.float32)
    prev = rng.uniform(0, 1, size=(n, 4)).astype(np.float32)
    prev[::9] = 0.0                                                   # first step of an episode: zero previous action
    with torch.no_grad():
        logits, _ = model.forward({"obs_history": torch.from_numpy(obs), "action_history": torch.from_numpy(prev), "is_training": False}, [], None)
        value = model.value_function()
        z = model.z
        dist = dists.MyBetaDist(logits, model)
        det = dist.deterministic_sample()
        x = torch.from_numpy(rng.uniform(0, 1, size=(n, 4)).astype(np.float32))
        x[::7] = 0.001                                                # below the 1e-2 clamp of logp
        x[3::11] = 0.9995
        logp = dist.logp(x)
        ent = dist.entropy()
        alpha, beta = dist.dist.concentration1, dist.dist.concentration0
    sd = {"sd_" + k: v.detach().cpu().numpy() for k, v in model.state_dict().items() if not k.startswith("adaptation_module")}
    np.savez_compressed(os.path.join(out_dir, "policy_ref.npz"), obs=obs, prev_action=prev, logits=logits.numpy(), value=value.numpy(), z=z.numpy(),
                        beta_x=x.numpy(), beta_det=det.numpy(), beta_logp=logp.numpy(), beta_entropy=ent.numpy(),
                        beta_alpha=alpha.numpy(), beta_beta=beta.numpy(), **sd)
    print("policy_ref.npz:", n, "rows;", len(sd), "state_dict tensors:", sorted(sd)[:4], "...")


if __name__ == "__main__":
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    gen_policy_ref(os.path.join(ROOT, "tests", "golden"))
