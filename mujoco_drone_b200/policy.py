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


# the reference module's state_dict names (RLlib SlimFC keeps its nn.Linear at `_model.0`) -> this module's
_REF_KEYS = {"param_encoder.0._model.0": "param_encoder.0", "param_encoder.1._model.0": "param_encoder.2",
             "_hidden_layers.0._model.0": "hidden.0", "_hidden_layers.1._model.0": "hidden.2", "_hidden_layers.2": "hidden.4",
             "_logits.0._model.0": "logits.0", "_logits.1._model.0": "logits.2",
             "_value_branch.0._model.0": "value_branch.0", "_value_branch.1._model.0": "value_branch.2", "_value_branch.2._model.0": "value_branch.4"}


def load_reference_state_dict(model, state_dict):
    """Copy a checkpoint of the REFERENCE's RMA_full (models/PPO/RMA/RMA_model.py:48-71; keys as saved by RLlib, e.g.
    `_hidden_layers.0._model.0.weight`) into an RMAFull built by make_rma_full; the adaptation module (unused with
    train_adaptation=False) is skipped.  Values may be torch tensors or numpy arrays."""
    import torch
    own = model.state_dict()
    seen = set()
    for k, v in state_dict.items():
        if k.startswith("adaptation_module"):
            continue
        prefix, leaf = k.rsplit(".", 1)
        if prefix not in _REF_KEYS:
            raise KeyError(f"unexpected key in the reference state_dict: {k}")
        name = _REF_KEYS[prefix] + "." + leaf
        t = torch.as_tensor(v)
        if tuple(own[name].shape) != tuple(t.shape):
            raise ValueError(f"{k}: shape {tuple(t.shape)} does not match {tuple(own[name].shape)}")
        own[name].copy_(t)
        seen.add(name)
    missing = [k for k in own if k not in seen]
    if missing:
        raise KeyError(f"reference state_dict lacks {missing}")
    return model


# ----------------------------------------------------------------------------------------------------------------------
# Fused tcgen05 inference of RMA_full (csrc/dsim_policy_mlp.cu, C ABI dsim_policy_*)

def _pack_umma_kmajor(w, n_pad, k_pad):
    """torch Linear weight [N, K] -> bf16 bits in the UMMA K-major no-swizzle canonical layout [K/8][N][8]
    (core matrices of 8 rows x 16 bytes; element (n, k) at ((k // 8) * N + n) * 8 + k % 8)."""
    import torch
    n, k = w.shape
    wp = torch.zeros((n_pad, k_pad), dtype=torch.float32)
    wp[:n, :k] = w.detach().float().cpu()
    wp = wp.reshape(n_pad, k_pad // 8, 8).permute(1, 0, 2).contiguous().to(torch.bfloat16)
    return wp.view(torch.int16).flatten()


def pack_rma_full(model):
    """(weights int16[W_ELEMS], consts float32[C_ELEMS]) for dsim_policy_create from an RMAFull module (eval mode).
    BatchNorm1d (running statistics) is an affine map h -> scale * h + shift: folded into the two layers that consume it."""
    import torch
    if model.num_states != 16 or model.num_params != 6 or model.num_actions != 4:
        raise ValueError("the fused kernel is specialised for RMA_full with 16 states, 6 params, 4 actions (train_RMA.py:47-53)")
    enc1, enc2 = model.param_encoder[0], model.param_encoder[2]
    h1, h2, bn = model.hidden[0], model.hidden[2], model.hidden[4]
    l1, l2 = model.logits[0], model.logits[2]
    v1, v2, v3 = model.value_branch[0], model.value_branch[2], model.value_branch[4]
    if enc2.out_features != 8 or h1.out_features != 256 or h2.out_features != 128 or l2.out_features != 8:
        raise ValueError("unexpected RMA_full layer sizes")
    with torch.no_grad():
        scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).float().cpu()
        shift = (bn.bias - bn.running_mean * bn.weight / torch.sqrt(bn.running_var + bn.eps)).float().cpu()
        w3 = l1.weight.float().cpu() * scale[None, :]
        b3 = l1.bias.float().cpu() + l1.weight.float().cpu() @ shift
        wv = v1.weight.float().cpu() * scale[None, :]
        bv = v1.bias.float().cpu() + v1.weight.float().cpu() @ shift
        blob = torch.cat([
            _pack_umma_kmajor(h1.weight, 256, 32),
            _pack_umma_kmajor(h2.weight, 128, 256),
            _pack_umma_kmajor(torch.cat([w3, wv], 0), 256, 128),
            _pack_umma_kmajor(l2.weight, 16, 128),
            _pack_umma_kmajor(v2.weight, 128, 128),
        ])
        c = torch.zeros(1408, dtype=torch.float32)
        f = lambda t: t.detach().float().cpu().flatten()
        c[0:256] = f(h1.bias); c[256:384] = f(h2.bias); c[384:512] = b3; c[512:640] = bv
        c[640:648] = f(l2.bias); c[656:784] = f(v2.bias); c[784:912] = f(v3.weight); c[912] = f(v3.bias)[0]
        c[913:913 + 192] = f(enc1.weight); c[1105:1137] = f(enc1.bias); c[1137:1137 + 256] = f(enc2.weight); c[1393:1401] = f(enc2.bias)
    return blob.contiguous(), c.contiguous()


class FusedRMAFull:
    """RMA_full forward as ONE tcgen05 kernel: `logits, value = fused(obs, prev_action)` on CUDA float32 tensors
    ([n, 22], [n, 4]) -> ([n, 8], [n]).  bf16 operands, FP32 accumulation.  No CPU fallback."""

    def __init__(self, model, device=0):
        import torch
        self._torch = torch
        L = self._L = _lib.load()
        we, ce = C.c_int64(), C.c_int64()
        L.dsim_policy_blob_sizes(C.byref(we), C.byref(ce))
        blob, consts = pack_rma_full(model)
        assert blob.numel() == we.value and consts.numel() == ce.value, (blob.numel(), we.value, consts.numel(), ce.value)
        self.device = torch.device("cuda", int(device))
        h = C.c_void_p()
        rc = L.dsim_policy_create(int(device), C.c_void_p(blob.data_ptr()), C.c_void_p(consts.data_ptr()), C.byref(h))
        if rc != _lib.OK:
            raise _lib.DsimError(rc, "dsim_policy_create failed (needs a CUDA device: there is no CPU fallback)")
        self._h = h

    def __call__(self, obs, prev_action, logits_out=None, value_out=None, reset_mask=None):
        torch = self._torch
        n = obs.shape[0]
        if obs.shape != (n, 22) or prev_action.shape != (n, 4) or obs.dtype != torch.float32 or prev_action.dtype != torch.float32:
            raise ValueError("FusedRMAFull expects float32 obs [n, 22] and prev_action [n, 4]")
        obs, prev_action = obs.contiguous(), prev_action.contiguous()
        logits = logits_out if logits_out is not None else torch.empty((n, 8), dtype=torch.float32, device=obs.device)
        value = value_out if value_out is not None else torch.empty((n,), dtype=torch.float32, device=obs.device)
        stream = C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream)
        if reset_mask is not None and (reset_mask.dtype != torch.uint8 or reset_mask.numel() != n or not reset_mask.is_contiguous()):
            raise ValueError("reset_mask must be a contiguous uint8 tensor [n]")
        rc = self._L.dsim_policy_forward(self._h, C.c_void_p(obs.data_ptr()), C.c_void_p(prev_action.data_ptr()),
                                         C.c_void_p(reset_mask.data_ptr()) if reset_mask is not None else None, n,
                                         C.c_void_p(logits.data_ptr()), C.c_void_p(value.data_ptr()), stream)
        if rc != _lib.OK:
            raise _lib.DsimError(rc, "dsim_policy_forward failed")
        return logits, value

    def sample(self, obs, prev_action, seed, env_id_offset=0, step=0, deterministic=False, actions_out=None, logp_out=None,
               value_out=None, logits_out=None, reset_mask=None, step_tensor=None, want_logits=False):
        """forward + Beta-head sampling in one launch -> (actions [n, 4], logp [n], value [n], logits [n, 8] | None).
        `actions_out` may be the `prev_action` tensor itself (in-place update).  Draws the numbers `beta_policy` draws."""
        torch = self._torch
        n = obs.shape[0]
        if obs.shape != (n, 22) or prev_action.shape != (n, 4) or obs.dtype != torch.float32 or prev_action.dtype != torch.float32:
            raise ValueError("FusedRMAFull expects float32 obs [n, 22] and prev_action [n, 4]")
        if not (obs.is_contiguous() and prev_action.is_contiguous()):
            raise ValueError("FusedRMAFull.sample expects contiguous inputs")
        mk = lambda shape: torch.empty(shape, dtype=torch.float32, device=obs.device)
        act = actions_out if actions_out is not None else mk((n, 4))
        lp = logp_out if logp_out is not None else mk((n,))
        value = value_out if value_out is not None else mk((n,))
        logits = logits_out if logits_out is not None else (mk((n, 8)) if want_logits else None)
        if reset_mask is not None and (reset_mask.dtype != torch.uint8 or reset_mask.numel() != n or not reset_mask.is_contiguous()):
            raise ValueError("reset_mask must be a contiguous uint8 tensor [n]")
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        rc = self._L.dsim_policy_forward_sample(self._h, ptr(obs), ptr(prev_action), ptr(reset_mask), n, int(seed) & 0xFFFFFFFF,
                                                int(env_id_offset) & 0xFFFFFFFF, int(step) & 0xFFFFFFFF, ptr(step_tensor),
                                                int(bool(deterministic)), ptr(logits), ptr(value), ptr(act), ptr(lp),
                                                C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream))
        if rc != _lib.OK:
            raise _lib.DsimError(rc, "dsim_policy_forward_sample failed")
        return act, lp, value, logits

    def check(self):
        """raise if any launch hit a tensor-core barrier timeout (device sync)"""
        if self._L.dsim_policy_error(self._h) != 0:
            raise RuntimeError("fused RMA_full kernel reported a barrier timeout: results invalid")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.dsim_policy_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ----------------------------------------------------------------------------------------------------------------------
# FP32-faithful fused inference of RMA_full (csrc/dsim_policy_fp32.cu, C ABI dsim_policy32_*)

def pack_rma_full_fp32(model):
    """float32 blob for dsim_policy32_create: per layer the TRANSPOSED weight [K][N] then the bias, BatchNorm1d (eval) folded
    into the two layers that consume it (the same algebra as pack_rma_full, in FP32)."""
    import torch
    if model.num_states != 16 or model.num_params != 6 or model.num_actions != 4:
        raise ValueError("the fused kernels are specialised for RMA_full with 16 states, 6 params, 4 actions (train_RMA.py:47-53)")
    enc1, enc2 = model.param_encoder[0], model.param_encoder[2]
    h1, h2, bn = model.hidden[0], model.hidden[2], model.hidden[4]
    l1, l2 = model.logits[0], model.logits[2]
    v1, v2, v3 = model.value_branch[0], model.value_branch[2], model.value_branch[4]
    f = lambda t: t.detach().double().cpu()
    with torch.no_grad():
        scale = f(bn.weight) / torch.sqrt(f(bn.running_var) + bn.eps)
        shift = f(bn.bias) - f(bn.running_mean) * scale
        w3 = torch.cat([f(l1.weight) * scale[None, :], f(v1.weight) * scale[None, :]], 0)            # [256][128]
        b3 = torch.cat([f(l1.bias) + f(l1.weight) @ shift, f(v1.bias) + f(v1.weight) @ shift])
        parts = [f(h1.weight).t(), f(h1.bias), f(h2.weight).t(), f(h2.bias), w3.t(), b3, f(v2.weight).t(), f(v2.bias),
                 f(l2.weight).t(), f(l2.bias), f(v3.weight).flatten(), f(v3.bias), f(enc1.weight), f(enc1.bias), f(enc2.weight), f(enc2.bias)]
        blob = torch.cat([x.contiguous().flatten() for x in parts]).float().contiguous()
    return blob


class FP32RMAFull:
    """RMA_full forward as ONE fused FP32 kernel (FP32 operands / accumulation, libm tanh: the reference's precision).
    `logits, value = net(obs, prev_action)` on CUDA float32 tensors ([n, 22], [n, 4]) -> ([n, 8], [n]).  No CPU fallback."""

    def __init__(self, model, device=0):
        import torch
        self._torch = torch
        L = self._L = _lib.load()
        blob = pack_rma_full_fp32(model)
        assert blob.numel() == L.dsim_policy32_blob_elems(), (blob.numel(), L.dsim_policy32_blob_elems())
        self.device = torch.device("cuda", int(device))
        h = C.c_void_p()
        rc = L.dsim_policy32_create(int(device), C.c_void_p(blob.data_ptr()), C.byref(h))
        if rc != _lib.OK:
            raise _lib.DsimError(rc, "dsim_policy32_create failed (needs a CUDA device: there is no CPU fallback)")
        self._h = h

    def __call__(self, obs, prev_action, logits_out=None, value_out=None, reset_mask=None):
        torch = self._torch
        n = obs.shape[0]
        if obs.shape != (n, 22) or prev_action.shape != (n, 4) or obs.dtype != torch.float32 or prev_action.dtype != torch.float32:
            raise ValueError("FP32RMAFull expects float32 obs [n, 22] and prev_action [n, 4]")
        obs, prev_action = obs.contiguous(), prev_action.contiguous()
        logits = logits_out if logits_out is not None else torch.empty((n, 8), dtype=torch.float32, device=obs.device)
        value = value_out if value_out is not None else torch.empty((n,), dtype=torch.float32, device=obs.device)
        if reset_mask is not None and (reset_mask.dtype != torch.uint8 or reset_mask.numel() != n or not reset_mask.is_contiguous()):
            raise ValueError("reset_mask must be a contiguous uint8 tensor [n]")
        rc = self._L.dsim_policy32_forward(self._h, C.c_void_p(obs.data_ptr()), C.c_void_p(prev_action.data_ptr()),
                                           C.c_void_p(reset_mask.data_ptr()) if reset_mask is not None else None, n,
                                           C.c_void_p(logits.data_ptr()), C.c_void_p(value.data_ptr()),
                                           C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream))
        if rc != _lib.OK:
            raise _lib.DsimError(rc, "dsim_policy32_forward failed")
        return logits, value

    def check(self):
        return None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.dsim_policy32_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
