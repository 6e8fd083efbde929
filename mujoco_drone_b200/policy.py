"""Caller side of the env path: the policy head the reference trains and the action distribution it
samples from, GPU-resident so that obs -> logits -> actions -> vector_step never leaves the device.

* `beta_policy`   — `MyBetaDist` (distributions.py:6-38): sampling / mean / log-probability, ONE CUDA kernel
                    (csrc/dsim_policy.cuh, through the C ABI `dsim_beta_policy`).  No CPU fallback.
* `RMAFull`       — `RMA_full` (models/PPO/RMA/RMA_model.py:19-136) inference graph in plain torch (the reference's class
                    needs ray's TorchModelV2; the layers, sizes, activations and initialisers are the same): parameter
                    encoder 6 -> 32 -> E, trunk (S + A + E) -> 256 -> 128 + BatchNorm, logits 128 -> 128 -> 2A,
                    value 128 -> 128 -> 128 -> 1.  The dense layers are library GEMMs (cuBLAS through torch).
"""
import ctypes as C

from . import _lib


def beta_policy(logits, seed, env_id_offset=0, step=0, deterministic=False, actions_out=None, logp_out=None, want_logp=True, step_tensor=None):
    """logits: CUDA tensor [n, 8] float32/float64 (alpha-logits then beta-logits).  Returns (actions [n, 4], logp [n] | None).
    `step_tensor`: optional CUDA int32 scalar tensor added to `step` on the device (advance it inside a CUDA graph)."""
    import torch
    if logits.dim() != 2 or logits.shape[1] != 8:
        raise ValueError("beta_policy expects [n, 8] logits (4 actions: alpha-logits then beta-logits)")
    if not logits.is_cuda:
        raise RuntimeError("beta_policy runs on the GPU only (there is no CPU fallback)")
    if logits.dtype not in (torch.float32, torch.float64):
        logits = logits.float()
    logits = logits.contiguous()
    n = logits.shape[0]
    L = _lib.load()
    act = actions_out if actions_out is not None else torch.empty((n, 4), dtype=logits.dtype, device=logits.device)
    lp = logp_out if logp_out is not None else (torch.empty((n,), dtype=logits.dtype, device=logits.device) if want_logp else None)
    stream = C.c_void_p(torch.cuda.current_stream(logits.device).cuda_stream)
    rc = L.dsim_beta_policy(C.c_void_p(logits.data_ptr()), n, _lib.FP64 if logits.dtype == torch.float64 else _lib.FP32,
                            int(seed) & 0xFFFFFFFF, int(env_id_offset), int(step) & 0xFFFFFFFF,
                            C.c_void_p(step_tensor.data_ptr()) if step_tensor is not None else None, int(bool(deterministic)),
                            C.c_void_p(act.data_ptr()), C.c_void_p(lp.data_ptr()) if lp is not None else None, stream)
    if rc != _lib.OK:
        raise _lib.DsimError(rc, "dsim_beta_policy failed")
    return act, lp


def _normc_(w, std):
    """ray.rllib normc_initializer: N(0,1) columns normalised to `std` over the input dimension."""
    import torch
    with torch.no_grad():
        w.normal_(0, 1)
        w *= std / torch.sqrt(w.pow(2).sum(1, keepdim=True))


def make_rma_full(num_states=16, num_params=6, num_actions=4, param_embed_dim=8, num_outputs=8, seed=42):
    """RMA_full with train_adaptation=False (models/PPO/RMA/RMA_model.py:48-71,79-109), random init like the reference
    (xavier_normal_ weights, zero biases, normc(0.01) for the last value layer)."""
    import torch
    from torch import nn

    class RMAFull(nn.Module):
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(seed)

            def fc(i, o, act, normc=None):
                lin = nn.Linear(i, o)
                if normc is None:
                    std = (2.0 / (i + o)) ** 0.5                                  # xavier_normal_, gain 1
                    with torch.no_grad():
                        lin.weight.copy_(torch.randn((o, i), generator=g) * std)
                else:
                    with torch.no_grad():
                        lin.weight.copy_(torch.randn((o, i), generator=g))
                        lin.weight.mul_(normc / torch.sqrt(lin.weight.pow(2).sum(1, keepdim=True)))
                nn.init.zeros_(lin.bias)
                return [lin, nn.Tanh()] if act else [lin]
            self.num_states, self.num_params, self.num_actions = num_states, num_params, num_actions
            self.param_encoder = nn.Sequential(*fc(num_params, 32, True), *fc(32, param_embed_dim, False))           # :48-51
            hid = num_states + num_actions + param_embed_dim
            self.hidden = nn.Sequential(*fc(hid, 256, True), *fc(256, 128, True), nn.BatchNorm1d(128))                # :55-60
            self.logits = nn.Sequential(*fc(128, 128, True), *fc(128, num_outputs, False))                            # :62-65
            self.value_branch = nn.Sequential(*fc(128, 128, True), *fc(128, 128, True), *fc(128, 1, False, normc=0.01))  # :67-71
            self.eval()                                                            # inference: BatchNorm uses running statistics (:88)

        def forward(self, obs, prev_action):
            """obs [n, S + P] (wrapper rows, e.g. LocalFrameRPYParamsEnv: 16 states + 6 params), prev_action [n, A].
            Returns (logits [n, 2A], value [n])."""
            s, e = obs[:, :self.num_states], obs[:, -self.num_params:]                                               # :94-96
            z = self.param_encoder(e)                                                                                 # :104
            f = self.hidden(torch.cat((s, prev_action, z), dim=-1)) if self.num_actions else self.hidden(torch.cat((s, z), dim=-1))
            return self.logits(f), self.value_branch(f).squeeze(1)                                                    # :106, :111-116
    return RMAFull()
