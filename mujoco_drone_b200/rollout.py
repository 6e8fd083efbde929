"""GPU-resident rollout loop: policy -> MyBetaDist -> vector_step, T steps, nothing leaves the device
(what RLlib's sampler / the reference's rollout.py:55-86 do with Python lists and one MuJoCo call per step).

    runner = RolloutRunner(env, policy, horizon=32)
    batch = runner.run()            # dict of [T, N, ...] device tensors, GAE-ready

Per step: RMA_full forward (library GEMMs) -> `dsim_beta_policy` (one kernel) -> `dsim_step` (one kernel) — or, with
policy_dtype="fused", ONE tcgen05 kernel for forward + sampling (`dsim_policy_forward_sample`) -> `dsim_step` — with the env's
in-kernel auto-reset; `prev_actions` is zeroed where an episode just ended, like RLlib's view requirement at an episode
start.  With `use_graph=True` one step is captured in a CUDA graph and replayed (the sampling kernel reads its step
counter from device memory, the step kernel's work-stealing counter re-arms itself)."""
import numpy as np


class RolloutRunner:
    def __init__(self, env, policy, horizon, seed=0, policy_dtype="fp32", use_graph=True, deterministic=False, history_len=0,
                 history_states=16, fuse_sampling=False):
        import torch
        if not env.auto_reset:
            raise ValueError("RolloutRunner needs an env created with auto_reset=True (the native loop)")
        self.torch, self.env, self.policy, self.T = torch, env, policy, int(horizon)
        self.N, self.D = env.num_drones, env.obs_dim
        self.seed, self.deterministic = int(seed), bool(deterministic)
        self.dev = env._device
        dt = env.obs_tensor.dtype
        if dt != torch.float32:
            raise ValueError("RolloutRunner drives the FP32 product path")
        self.policy_dtype = policy_dtype
        # "fused" policy only: draw the action inside the policy kernel (2 launches per step instead of 3, no logits round trip).
        # Off by default: at 524288 envs the sampling code runs at the policy kernel's 8 warps/SM and costs more there
        # (+47 us) than the separate 64-warps/SM sampling launch (34 us); measured 276.6 vs 264.5 us per step on a B200.
        self.fuse_sampling = bool(fuse_sampling)
        T, N, D = self.T, self.N, self.D
        z = dict(device=self.dev)
        self.obs = torch.zeros((T + 1, N, D), dtype=dt, **z)
        self.actions = torch.zeros((T, N, 4), dtype=dt, **z)
        self.rewards = torch.zeros((T, N), dtype=dt, **z)
        self.truncated = torch.zeros((T, N), dtype=torch.uint8, **z)
        self.values = torch.zeros((T + 1, N), dtype=dt, **z)
        self.logp = torch.zeros((T, N), dtype=dt, **z)
        self.prev_actions = torch.zeros((N, 4), dtype=dt, **z)
        # static single-step buffers (graph capture needs fixed addresses)
        self._obs_cur = torch.zeros((N, D), dtype=dt, **z)
        self._act = torch.zeros((N, 4), dtype=dt, **z)
        self._logp = torch.zeros((N,), dtype=dt, **z)
        self._val = torch.zeros((N,), dtype=dt, **z)
        self._step_ctr = torch.zeros((1,), dtype=torch.int32, **z)
        # 32-step (state, previous action) windows of the adaptation module / state estimators (RMA_model.py:41-43 obs_history
        # "-31:0", action_history "-32:-1"; StateEstimatorLSTM.py:239-244): a time-major device ring buffer, one row per step
        self.history_len, self.history_states = int(history_len), int(history_states)
        if self.history_len:
            self._hist = torch.zeros((self.history_len, N, self.history_states + 4), dtype=dt, **z)
            self._hist_age = torch.zeros((N,), dtype=torch.int32, **z)      # steps since the env's episode started (caps the valid window)
            self._hist_pos = 0
        self.policy = policy.to(self.dev)
        self._fused = None
        if policy_dtype == "fused":                    # hand-written tcgen05 kernel (csrc/dsim_policy_mlp.cu) instead of library GEMMs
            from .policy import FusedRMAFull
            self._fused = FusedRMAFull(self.policy, device=env.device_index)
            self._logits = torch.zeros((N, 8), dtype=dt, **z)
        elif policy_dtype == "fused_fp32":             # hand-written FP32-pipe kernel (csrc/dsim_policy_fp32.cu): the reference's precision
            from .policy import FP32RMAFull
            if fuse_sampling:
                raise ValueError("fuse_sampling needs policy_dtype='fused' (the tcgen05 kernel)")
            self._fused = FP32RMAFull(self.policy, device=env.device_index)
            self._logits = torch.zeros((N, 8), dtype=dt, **z)
        self.use_graph, self._graph = bool(use_graph), None
        self.total_steps = 0
        if policy_dtype == "tf32":
            torch.backends.cuda.matmul.allow_tf32 = True
        # the policy reads the env's observation buffer in place: the step kernel overwrites it only after the policy and the
        # sampling kernel of the same step have run (stream order)
        self._obs_cur = env.reset_tensor()
        self._mask = env.truncated_tensor
        self._mask.fill_(1)                             # every env starts an episode: zero previous action

    # one policy + sample + env step on the static buffers
    def _forward(self):
        torch = self.torch
        if self._fused is not None:   # reads the env's own observation / truncated buffers and last step's actions: no copies, no masking kernels
            return self._fused(self._obs_cur, self._act, logits_out=self._logits, value_out=self._val, reset_mask=self._mask)
        with torch.no_grad():
            if self.policy_dtype == "bf16":
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    logits, value = self.policy(self._obs_cur, self.prev_actions)
                logits, value = logits.float(), value.float()
            else:
                logits, value = self.policy(self._obs_cur, self.prev_actions)
        return logits, value

    def _one_step(self):
        from .policy import beta_policy
        if self._fused is not None and self.fuse_sampling:   # forward + sampling in one launch; actions updated in place
            self._fused.sample(self._obs_cur, self._act, self.seed, self.env.env_id_offset, 0, self.deterministic, actions_out=self._act,
                               logp_out=self._logp, value_out=self._val, reset_mask=self._mask, step_tensor=self._step_ctr)
        else:
            logits, value = self._forward()
            if value.data_ptr() != self._val.data_ptr():
                self._val.copy_(value)
            beta_policy(logits, self.seed, self.env.env_id_offset, 0, self.deterministic, actions_out=self._act, logp_out=self._logp,
                        step_tensor=self._step_ctr)
        self._step_ctr.add_(1)
        obs, rew, trunc = self.env.step_tensor(self._act)
        self._obs_next, self._rew, self._trunc = obs, rew, trunc
        if self._fused is None:   # first action of a new episode sees zeros as its previous action
            self.prev_actions.copy_(self._act * (trunc == 0).to(self._act.dtype).unsqueeze(1))

    WARMUP_STEPS = 3

    def warm_up(self, steps=None):
        """`steps` UNRECORDED policy + sample + env steps (default WARMUP_STEPS = what the CUDA-graph capture runs first:
        lazy init, cuBLAS workspaces, shared-memory opt-in).  They advance the env, the sampling step counter and
        env.total_steps like recorded steps do; an eager runner that calls warm_up() once is bit-identical to a graph runner."""
        for _ in range(self.WARMUP_STEPS if steps is None else steps):
            self._one_step()

    def _capture(self):
        torch = self.torch
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            self.warm_up()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        ts = self.env.total_steps
        with torch.cuda.graph(g):
            self._one_step()
        self.env.total_steps = ts                # the captured pass queued nothing: only replays advance the env
        self._graph = g

    def step(self, t):
        """advance one env-step and record it at row t of the rollout buffers"""
        if self.use_graph and self._graph is None:
            self._capture()
        self.obs[t].copy_(self._obs_cur)
        if self.history_len:
            self._record_history()
        if self._graph is not None:
            self._graph.replay()
            self.env.total_steps += 1            # what step_tensor does on the eager path
            self.env._states_cache = None; self.env._states_frozen = False
        else:
            self._one_step()
        self.values[t].copy_(self._val)
        self.actions[t].copy_(self._act)
        self.logp[t].copy_(self._logp)
        self.rewards[t].copy_(self._rew)
        self.truncated[t].copy_(self._trunc)
        self.total_steps += 1

    def _record_history(self):
        """append (current state part of the observation, PREVIOUS action) of every env; the previous action of an episode's
        first step is zero and the env's valid window restarts there (RLlib zero-pads views before an episode start)"""
        torch = self.torch
        fresh = self._mask if self._fused is not None else None
        prev = self._act if self._fused is not None else self.prev_actions
        if fresh is None:                                       # torch policies: prev_actions is already zeroed at episode starts
            started = (self.prev_actions.abs().sum(1) == 0)
        else:
            started = fresh != 0
        row = self._hist[self._hist_pos]
        row[:, :self.history_states] = self._obs_cur[:, :self.history_states]
        row[:, self.history_states:] = torch.where(started.unsqueeze(1), torch.zeros_like(prev), prev)
        self._hist_age = torch.where(started, torch.ones_like(self._hist_age), torch.clamp(self._hist_age + 1, max=self.history_len))
        self._hist_pos = (self._hist_pos + 1) % self.history_len

    def history(self):
        """[N, history_len, states + 4] windows, oldest first, ending at the most recently recorded step; entries from before
        the env's current episode are zero (what RLlib's ViewRequirement hands the adaptation module / estimators)"""
        torch = self.torch
        if not self.history_len:
            raise ValueError("create the runner with history_len > 0")
        order = [(self._hist_pos + k) % self.history_len for k in range(self.history_len)]
        h = self._hist[order].permute(1, 0, 2).clone()           # [N, L, F], oldest first
        k = torch.arange(self.history_len, device=self.dev).unsqueeze(0)
        valid = k >= (self.history_len - self._hist_age.unsqueeze(1))
        return h * valid.unsqueeze(2).to(h.dtype)

    def run(self):
        """T steps; returns the [T(+1), N, ...] device tensors (bootstrap value of the last observation included)."""
        for t in range(self.T):
            self.step(t)
        self.obs[self.T].copy_(self._obs_cur)
        _, v = self._forward()
        self.values[self.T].copy_(v)
        if self._fused is not None:
            self._fused.check()                  # a tensor-core barrier timeout invalidates the whole batch: raise, never return it
        return dict(obs=self.obs, actions=self.actions, rewards=self.rewards, truncated=self.truncated, values=self.values,
                    action_logp=self.logp)

    def to_reference_dataset(self, batch=None):
        """the dump format of the reference's rollout.py:68-85: {'z': params, 'o': observations, 'a': actions, 't': truncated}
        as host numpy arrays, env-major like its per-drone lists"""
        b = batch or dict(obs=self.obs, actions=self.actions, truncated=self.truncated)
        o = b["obs"][:self.T].permute(1, 0, 2).cpu().numpy().astype(np.float64)
        return {"z": np.array([list(d.values()) for d in self.env.drone_params]), "o": o,
                "a": b["actions"].permute(1, 0, 2).cpu().numpy().astype(np.float64),
                "t": b["truncated"].permute(1, 0).cpu().numpy().astype(bool)}
