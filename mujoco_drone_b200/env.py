"""BaseDroneEnv: drop-in for the reference's `environments/BaseDroneEnv.py` on top of libdronesim_b200.so.

Same surface as the reference (RLlib `VectorEnv`): `vector_reset`, `vector_step`, `reset_at`, `reset_model`,
`get_drone_states`, `observation_space`, `action_space`, `num_drones`, `reference`, `states`, `drone_params`,
same `config` keys and `.get` fall-backs (BaseDroneEnv.py:60-106), same error behaviour.  The compat methods
move host lists/arrays; the `*_tensor` methods are the native loop: device tensors, zero-copy, no sync.

Every env-step is ONE CUDA kernel launch (csrc/dsim_kernels.cu::step_kernel).  There is no CPU path.
"""
import ctypes as C
import types
import warnings

import numpy as np

from . import _lib
from .rewards import default_reward_fcn, resolve_reward

try:  # the real RLlib / gymnasium bases when they are installed, small shims otherwise
    from ray.rllib.env.vector_env import VectorEnv as _VectorEnv
except Exception:  # pragma: no cover - ray is not in this image
    class _VectorEnv:
        def __init__(self, observation_space, action_space, num_envs):
            self.observation_space = observation_space
            self.action_space = action_space
            self.num_envs = num_envs

try:
    from gymnasium.spaces import Box
except Exception:  # pragma: no cover
    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float64, seed=None):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

        def sample(self):
            lo = np.broadcast_to(np.asarray(self.low, dtype=np.float64), self.shape)
            hi = np.broadcast_to(np.asarray(self.high, dtype=np.float64), self.shape)
            return np.random.uniform(np.where(np.isfinite(lo), lo, -1), np.where(np.isfinite(hi), hi, 1)).astype(self.dtype)


def default_termination_fcn(env, state, action, num_steps):
    """Token for BaseDroneEnv.py:12-16; evaluated in FP64 inside the step kernel."""
    raise NotImplementedError("default_termination_fcn is evaluated inside the fused CUDA step kernel")


# BaseDroneEnv.py:19-50 — identical keys and values (callables are the device tokens)
base_config = {'seed': 42, 'frequency': 100, 'skip_steps': 1, 'reference': [0, 0, 15, 0], 'start_pos': [0, 0, 15, 0],
               'max_distance': 4, 'random_start_pos': True, 'random_params': True, 'pendulum': True,
               'state_difficulty': 0.4, 'param_difficulty': 0.1, 'max_random_offset': 2, 'rp_variance': [0.8, 0.8],
               'vel_variance': [1, 1, 1], 'ang_vel_variance': [1, 1, 1], 'mass_interval': [1, 0.1],
               'arm_len_interval': [0.17, 0.02], 'motor_force_interval': [7, 1], 'motor_tau_interval': [0.01, 0.0025],
               'pendulum_length_interval': [1.2, 0.2], 'weight_mass_interval': [0.3, 0.05],
               'pendulum_rp_variance': [0.5, 0.5], 'pendulum_ang_vel_variance': [0.5, 0.5],
               'reward_fcn': default_reward_fcn, 'terminated_fcn': default_termination_fcn, 'max_steps': 512,
               'regen_env_at_steps': None, 'train_vis': 0, 'window_title': 'mujoco', 'controlled': False, 'mocaps': 1}

PARAM_KEYS = ('mass', 'arm_len', 'motor_force', 'motor_tau', 'pendulum_len', 'weight_mass')
_NP_STR = {_lib.DT_F32: '<f4', _lib.DT_F64: '<f8', _lib.DT_I32: '<i4', _lib.DT_U8: '|u1', _lib.DT_U32: '<i4', _lib.DT_I64: '<i8'}
_ITEM = {_lib.DT_F32: 4, _lib.DT_F64: 8, _lib.DT_I32: 4, _lib.DT_U8: 1, _lib.DT_U32: 4, _lib.DT_I64: 8}


class _DevView:
    """Minimal __cuda_array_interface__ carrier so torch can wrap handle-owned device memory without a copy."""

    def __init__(self, ptr, shape, strides, dtype, owner):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'strides': tuple(strides) if strides else None,
                                         'typestr': _NP_STR[dtype], 'data': (int(ptr), False), 'version': 3}
        self._owner = owner


class BaseDroneEnv(_VectorEnv):
    OBS_ID = 0                 # DSIM_OBS_BASE: raw get_drone_states rows (BaseDroneEnv.py:353-355)
    DECLARED_OBS = None        # observation_space size the reference class DECLARES (None: 27|23 + 6)

    def __init__(self, config, **kwargs):
        import torch
        self._torch = torch
        L = self._L = _lib.load()
        g = config.get
        self.width, self.height = 640, 480
        # --- reference keys (BaseDroneEnv.py:60-106) with the reference's own fall-backs
        self.controlled = g('controlled', False)
        self.render_mode = None                                   # headless: viewer / joystick are out of scope
        self.window_title = g('window_title', 'mujoco')
        self.mocaps = g('mocaps', 1)
        self.skip_steps = g('skip_steps', 1)
        self.frame_skip = self.skip_steps
        self.frequency = g('frequency', 200)
        self._reference = np.array(g('reference', [0, 0, 0, 0]), dtype=np.float64)
        self.num_drones = g('num_drones', 1)
        self.pendulum = g('pendulum', True)
        self.mass_interval = np.array(g('mass_interval', [1.35, 0.15]))
        self.arm_len_interval = np.array(g('arm_len_interval', [0.17, 0.02]))
        self.motor_force_interval = np.array(g('motor_force_interval', [7.5, 1.5]))
        self.motor_tau_interval = np.array(g('motor_tau_interval', [0.003, 0.002]))
        self.pendulum_length_interval = np.array(g('pendulum_length_interval', [1.2, 0.3]))
        self.weight_mass_interval = np.array(g('weight_mass_interval', [0.2, 0.1]))
        self.state_difficulty = g('state_difficulty', 0.1)
        self.param_difficulty = g('param_difficulty', 0.1)
        self.random_start_pos = g('random_start_pos', False)
        self.random_params = g('random_params', False)
        self.regen_env_at_steps = g('regen_env_at_steps', None)
        self.start_pos = g('start_pos', list(self._reference))
        self.max_distance = g('max_distance', 1)
        self.reward_fcn = g('reward_fcn', default_reward_fcn)
        self.terminated_fcn = g('terminated_fcn', default_termination_fcn)
        self.max_steps = g('max_steps', 512)
        self.max_pos_offset = self.state_difficulty * g('max_random_offset', 0)
        self.angle_variance = self.state_difficulty * np.array(g('angle_variance', [0, 0]))      # Q4: 'rp_variance' is never read
        self.ang_vel_variance = self.state_difficulty * np.array(g('ang_vel_variance', [0, 0, 0]))
        self.vel_variance = self.state_difficulty * np.array(g('vel_variance', [0, 0, 0]))
        self.pendulum_rp_variance = self.state_difficulty * np.array(g('pendulum_rp_variance', [0, 0]))
        self.pendulum_ang_vel_variance = self.state_difficulty * np.array(g('pendulum_ang_vel_variance', [0, 0]))
        self.total_steps = 0
        if getattr(self.terminated_fcn, '__name__', None) != 'default_termination_fcn':
            raise NotImplementedError("only default_termination_fcn (BaseDroneEnv.py:12-16) runs inside the step kernel")
        self.reward_id = resolve_reward(self.reward_fcn)
        # --- extension keys of this implementation
        self.device_index = int(g('device', torch.cuda.current_device() if torch.cuda.is_available() else 0))
        self.precision = g('precision', 'fp32')
        self.auto_reset = bool(g('auto_reset', False))
        self.per_env_reference = bool(g('per_env_reference', False))
        self.env_id_offset = int(g('env_id_offset', 0))
        self.round_precision = bool(g('round_precision', True))
        self._inputs_ready = bool(g('inputs_ready', False))
        # floor contact (env_gen.py:14-21,97): off by default - the floor is out of reach in the training configs (z = 15 m,
        # truncation at 4 m) and would-be contacts are only counted (episode_stats()['n_near_ground']); on: simulated
        gc = g('ground_contact', None)
        if gc is None:
            # not asked for: off (the specialised in-air kernels run), but say so when a drone could get down to the floor before it
            # is truncated - lowest setpoint (control_reference clips the joystick setpoint to start_pos[2] - 6, :167-170) minus
            # max_distance, against the longest drone (pendulum tip <= 2.5 m below the body).  BASELINE configs: 15 - 6 - 4 = 5 m.
            lowest = min(float(self._reference[2]), float(self.start_pos[2]) - (6.0 if (self.controlled or self.per_env_reference) else 0.0)) - float(self.max_distance)
            if lowest < 2.5:
                warnings.warn("the floor (z = 0) is within reach of this configuration but 'ground_contact' is not set: floor contacts are "
                              "only COUNTED (episode_stats()['n_near_ground']), not simulated; pass ground_contact=True to simulate them", stacklevel=2)
            gc = False
        self.ground_contact = bool(gc)
        self._regen_epoch = 0
        # seed: reference uses config.get('worker_index', -1) + 1 + seed (BaseDroneEnv.py:113, Q5)
        self.seed_value = int(g('worker_index', -1) + 1 + g('seed', 1))

        self.num_params = 6
        self.num_states = 27 if self.pendulum else 23
        declared = self.DECLARED_OBS if self.DECLARED_OBS is not None else self.num_states + self.num_params
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=(declared,), dtype=np.float64)
        self.action_space = Box(low=0, high=1, shape=(4,), dtype=np.float64)
        self.metadata = {"render_modes": ["human", "rgb_array", "depth_array"], "render_fps": self.frequency // self.skip_steps}

        cfg = _lib.DsimConfig()
        cfg.struct_size = C.sizeof(_lib.DsimConfig)
        cfg.abi_version = _lib.ABI_VERSION
        cfg.num_envs = int(self.num_drones)
        cfg.precision = _lib.FP64 if self.precision == 'fp64' else _lib.FP32
        cfg.env_id_offset = self.env_id_offset
        cfg.seed = self.seed_value & 0xFFFFFFFF
        cfg.pendulum = int(bool(self.pendulum))
        cfg.frame_skip = int(self.skip_steps)
        cfg.round_precision = int(self.round_precision)
        cfg.frequency = float(self.frequency)
        cfg.obs_id = int(self.OBS_ID)
        cfg.reward_id = int(self.reward_id)
        cfg.per_env_reference = int(self.per_env_reference)
        cfg.ground_contact = int(self.ground_contact)
        cfg.auto_reset = int(self.auto_reset)
        cfg.random_start_pos = int(bool(self.random_start_pos))
        cfg.random_params = int(bool(self.random_params))
        cfg.reference[:] = [float(x) for x in self._reference]
        cfg.start_pos[:] = [float(x) for x in self.start_pos]
        cfg.max_distance = float(self.max_distance)
        cfg.max_steps = int(self.max_steps)
        cfg.max_pos_offset = float(self.max_pos_offset)
        cfg.angle_sigma[:] = [float(x) for x in self.angle_variance]
        cfg.vel_sigma[:] = [float(x) for x in self.vel_variance]
        cfg.ang_vel_sigma[:] = [float(x) for x in self.ang_vel_variance]
        cfg.pend_rp_sigma[:] = [float(x) for x in self.pendulum_rp_variance]
        cfg.pend_vel_sigma[:] = [float(x) for x in self.pendulum_ang_vel_variance]
        iv = [self.mass_interval, self.arm_len_interval, self.motor_force_interval, self.motor_tau_interval,
              self.pendulum_length_interval, self.weight_mass_interval]
        cfg.param_center[:] = [float(x[0]) for x in iv]
        cfg.param_halfwidth[:] = [float(x[1]) for x in iv]
        cfg.param_difficulty = float(self.param_difficulty)
        self._cfg = cfg
        h = C.c_void_p()
        rc = L.dsim_create(C.byref(cfg), self.device_index, C.byref(h))
        if rc != _lib.OK:
            raise _lib.DsimError(rc, (L.dsim_last_error(None) or b"").decode())
        self._h = h
        if self._inputs_ready:
            self.inputs_ready = True
        self._np_dtype = np.float64 if cfg.precision == _lib.FP64 else np.float32
        self.obs_dim = L.dsim_obs_dim(self.OBS_ID, cfg.pendulum)
        self._views = {}
        self._device = torch.device('cuda', self.device_index)
        n, d = self.num_drones, self.obs_dim
        pin = dict(pin_memory=True)
        tdt = torch.float64 if cfg.precision == _lib.FP64 else torch.float32
        self._h_actions = torch.empty((n, 4), dtype=tdt, **pin)
        self._h_obs = torch.empty((n, d), dtype=tdt, **pin)
        self._h_reward = torch.empty((n,), dtype=tdt, **pin)
        self._h_trunc = torch.empty((n,), dtype=torch.uint8, **pin)
        self._d_actions = torch.empty((n, 4), dtype=tdt, device=self._device)
        # numpy views / raw pointers of the pinned staging buffers (the compat vector_step path)
        self._np_actions, self._np_obs = self._h_actions.numpy(), self._h_obs.numpy()
        self._p_actions, self._p_obs = self._h_actions.data_ptr(), self._h_obs.data_ptr()
        self._p_reward, self._p_trunc = self._h_reward.data_ptr(), self._h_trunc.data_ptr()
        self._last_obs = None
        self._sensor_stale = False
        self._states_cache = None; self._states_frozen = False
        _VectorEnv.__init__(self, self.observation_space, self.action_space, self.num_drones)

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(self._torch.cuda.current_stream(self._device).cuda_stream)

    def _ck(self, rc):
        return _lib.check(self._h, rc)

    def tensor(self, buf_id):
        """Zero-copy torch view of a handle-owned device buffer (include/dronesim_b200.h DSIM_BUF_*).
        Dense buffers come back as [cols] / [rows, cols].  PAGED buffers (state, num_steps, params, consts,
        reference, ep_return) come back as [rows, npages, 32]: env i is element [..., i // 32, i % 32]."""
        if buf_id in self._views:
            return self._views[buf_id]
        ptr, rows, cols, ld, dt, pr = C.c_void_p(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32(), C.c_int64()
        self._ck(self._L.dsim_buffer(self._h, buf_id, C.byref(ptr), C.byref(rows), C.byref(cols), C.byref(ld), C.byref(dt), C.byref(pr)))
        item = _ITEM[dt.value]
        if pr.value:
            T = _lib.TILE
            view = _DevView(ptr.value, (rows.value, ld.value // T, T), (T * item, pr.value * T * item, item), dt.value, self)
        elif rows.value == 1 and buf_id not in (_lib.BUF_OBS, _lib.BUF_STATES33):      # vectors; [1, D] stays 2-D for num_drones == 1
            view = _DevView(ptr.value, (cols.value,), None, dt.value, self)
        else:
            view = _DevView(ptr.value, (rows.value, cols.value), (ld.value * item, item), dt.value, self)
        t = self._torch.as_tensor(view, device=self._device)
        self._views[buf_id] = t
        return t

    def rows(self, buf_id):
        """Copy of a PAGED buffer as a dense [rows, num_drones] tensor."""
        t = self.tensor(buf_id)
        return t.reshape(t.shape[0], -1)[:, :self.num_drones]

    def write_rows(self, buf_id, row0, values):
        """values [k, num_drones] -> rows row0..row0+k of a PAGED buffer (in place on the device)."""
        t = self.tensor(buf_id)
        v = self._torch.as_tensor(values, device=self._device).to(t.dtype)
        k, n, T = v.shape[0], self.num_drones, _lib.TILE
        full, rem = n // T, n % T
        if full:
            t[row0:row0 + k, :full, :] = v[:, :full * T].reshape(k, full, T)
        if rem:
            t[row0:row0 + k, full, :rem] = v[:, full * T:]

    @property
    def obs_tensor(self):
        return self.tensor(_lib.BUF_OBS)

    @property
    def reward_tensor(self):
        return self.tensor(_lib.BUF_REWARD)

    @property
    def truncated_tensor(self):
        return self.tensor(_lib.BUF_TRUNCATED)

    @property
    def state_tensor(self):
        """[21, npages, 32] zero-copy view of the state rows qpos / qvel / act (see `tensor`; sensordata: BUF_SENSORDATA)."""
        return self.tensor(_lib.BUF_STATE)

    @property
    def num_steps_tensor(self):
        """[num_drones] int copy of BaseDroneEnv.num_steps."""
        return self.rows(_lib.BUF_NUM_STEPS)[0]

    @property
    def reference_tensor(self):
        """[4, npages, 32] zero-copy view of the per-env setpoints."""
        return self.tensor(_lib.BUF_REFERENCE)

    # ------------------------------------------------------------------ native (device tensor) loop
    def step_tensor(self, actions):
        """actions: CUDA tensor [N,4] in [0,1] (policy output).  Returns (obs, reward, truncated) device views that are
        overwritten by the next step.  One kernel launch, no host sync."""
        if tuple(actions.shape) != (self.num_drones, 4):
            raise ValueError("Action dimension mismatch")                       # mujoco_env_custom.py:200-201
        if actions.dtype != self._d_actions.dtype or not actions.is_contiguous() or actions.device != self._device:
            actions = actions.to(device=self._device, dtype=self._d_actions.dtype).contiguous()
        self._ck(self._L.dsim_step(self._h, C.c_void_p(actions.data_ptr()), self._stream()))
        self.total_steps += 1
        self._sensor_stale = False
        self._states_cache = None; self._states_frozen = False
        return self.obs_tensor, self.reward_tensor, self.truncated_tensor

    def evaluate_tensor(self, actions):
        """termination / reward / observation of the CURRENT state (no physics, no counters): what the reference's
        terminated_fcn / reward_fcn / _get_obs return on `self.states`."""
        if tuple(actions.shape) != (self.num_drones, 4):
            raise ValueError("Action dimension mismatch")
        actions = actions.to(device=self._device, dtype=self._d_actions.dtype).contiguous()
        self._ck(self._L.dsim_evaluate(self._h, C.c_void_p(actions.data_ptr()), self._stream()))
        return self.obs_tensor, self.reward_tensor, self.truncated_tensor

    def reset_tensor(self):
        self._ck(self._L.dsim_reset_all(self._h, self._stream()))
        self._sensor_stale = False                                          # dsim_reset_all ends with mj_forward
        self._states_cache = None; self._states_frozen = False
        return self.obs_tensor

    def reset_masked(self, mask):
        """mask: CUDA uint8/bool tensor [N]; re-samples the flagged envs (what RLlib does through reset_at)."""
        m = mask.to(device=self._device, dtype=self._torch.uint8).contiguous()
        self._ck(self._L.dsim_reset_masked(self._h, C.c_void_p(m.data_ptr()), self._stream()))
        self._sensor_stale = True
        self._states_cache = None; self._states_frozen = False

    def step_host(self, actions, obs_out=None, reward_out=None, trunc_out=None):
        """End-to-end path with HOST float32 arrays.  Pinned arrays (e.g. views of torch `pin_memory()` tensors): one kernel
        that reads the actions and writes obs / reward / truncated over PCIe itself; pageable arrays: staged copies."""
        a = np.ascontiguousarray(actions, dtype=np.float32)
        if a.shape != (self.num_drones, 4):
            raise ValueError("Action dimension mismatch")
        obs_out = np.empty((self.num_drones, self.obs_dim), np.float32) if obs_out is None else obs_out
        reward_out = np.empty(self.num_drones, np.float32) if reward_out is None else reward_out
        trunc_out = np.empty(self.num_drones, np.uint8) if trunc_out is None else trunc_out
        # the library writes through raw pointers (from the GPU, over PCIe, on the pinned path): anything but the exact
        # C-contiguous layout would be a silent out-of-bounds / garbled write
        for name, arr, shape, dt in (("obs_out", obs_out, (self.num_drones, self.obs_dim), np.float32),
                                     ("reward_out", reward_out, (self.num_drones,), np.float32),
                                     ("trunc_out", trunc_out, (self.num_drones,), np.uint8)):
            if not (isinstance(arr, np.ndarray) and arr.dtype == dt and arr.shape == shape and arr.flags['C_CONTIGUOUS'] and arr.flags['WRITEABLE']):
                raise ValueError(f"{name} must be a writable C-contiguous {np.dtype(dt).name} array of shape {shape}")
        ptr = lambda x: x.__array_interface__['data'][0]                     # (ndarray.ctypes builds a helper object per access)
        self._ck(self._L.dsim_step_host(self._h, ptr(a), ptr(obs_out), ptr(reward_out), ptr(trunc_out), self._stream()))
        self.total_steps += 1
        self._sensor_stale = False
        self._states_cache = None; self._states_frozen = False
        return obs_out, reward_out, trunc_out

    def episode_stats(self, reset=False):
        out = np.zeros(8)
        self._ck(self._L.dsim_stats(self._h, out.ctypes.data_as(C.POINTER(C.c_double)), int(reset)))
        return dict(sum_return=out[0], sum_length=out[1], n_episodes=out[2], n_nonfinite=out[3], n_near_ground=out[4])

    @property
    def inputs_ready(self):
        """dsim_set_inputs_ready: promise that the kernel queued right before every step_tensor() call does not write this
        env's actions / state / setpoints (several independent env shards interleaved on one stream)."""
        return self._inputs_ready

    @inputs_ready.setter
    def inputs_ready(self, value):
        self._inputs_ready = bool(value)
        self._ck(self._L.dsim_set_inputs_ready(self._h, int(self._inputs_ready)))

    # ------------------------------------------------------------------ reference attributes
    @property
    def reference(self):
        return self._reference

    @reference.setter
    def reference(self, value):                                               # assigned by evaluation.py:48,66
        self._reference = np.array(value, dtype=np.float64)
        if hasattr(self, '_h'):
            self._ck(self._L.dsim_set_reference(self._h, self._reference.ctypes.data_as(C.POINTER(C.c_double))))
            if self.per_env_reference:
                r = self.reference_tensor
                off = self._reference.copy()
                off[:3] -= np.asarray(self.start_pos[:3], dtype=np.float64)
                r[:, :, :] = self._torch.as_tensor(off, dtype=r.dtype, device=self._device)[:, None, None]
            self._states_cache = None; self._states_frozen = False

    @property
    def states(self):
        """`self.states` of the reference (:148,273,325): refreshed by vector_step / reset_model, NOT by reset_at (Q1).
        Computed lazily; reset_at freezes the pre-reset rows first so the staleness is preserved."""
        if self._states_cache is None:
            if self._states_frozen:      # rows computed on the device before a reset_at changed the state: only the read-back was deferred
                self._states_cache = list(self.tensor(_lib.BUF_STATES33).to('cpu').numpy().astype(np.float64))
            else:
                self._states_cache = self.get_drone_states()
        return self._states_cache

    @property
    def num_steps(self):
        return self.num_steps_tensor.to('cpu').numpy().astype(np.int64)         # np.long array (:110)

    @property
    def drone_params(self):
        p = np.zeros((self.num_drones, 6))
        self._ck(self._L.dsim_get_params(self._h, p.ctypes.data_as(C.POINTER(C.c_double))))
        return [dict(zip(PARAM_KEYS, row)) for row in p]

    @drone_params.setter
    def drone_params(self, value):
        p = np.ascontiguousarray([[d[k] for k in PARAM_KEYS] for d in value], dtype=np.float64)
        assert p.shape == (self.num_drones, 6)
        self._ck(self._L.dsim_set_params(self._h, p.ctypes.data_as(C.POINTER(C.c_double)), self._stream()))

    def compiled_constants(self):
        c = np.zeros((self.num_drones, 13))
        self._ck(self._L.dsim_get_consts(self._h, c.ctypes.data_as(C.POINTER(C.c_double))))
        return c

    @property
    def data(self):
        """Snapshot with MjData's field names in the reference's drone-major flat layout (BaseDroneEnv.py:367-375)."""
        qpos, qvel, act, sens, _ = self.get_state()
        return types.SimpleNamespace(qpos=qpos.ravel(), qvel=qvel.ravel(), act=act.ravel(), sensordata=sens.ravel())

    def get_state(self):
        n, p = self.num_drones, 2 * int(bool(self.pendulum))
        qpos, qvel, act, sens = np.zeros((n, 7 + p)), np.zeros((n, 6 + p)), np.zeros((n, 4)), np.zeros((n, 3))
        ns = np.zeros(n, dtype=np.int32)
        dp = C.POINTER(C.c_double)
        self._ck(self._L.dsim_get_state(self._h, qpos.ctypes.data_as(dp), qvel.ctypes.data_as(dp), act.ctypes.data_as(dp),
                                        sens.ctypes.data_as(dp), ns.ctypes.data_as(C.POINTER(C.c_int32))))
        return qpos, qvel, act, sens, ns

    def set_state(self, qpos, qvel, act=None, num_steps=None):
        """extendedEnv.set_state (mujoco_vecenv.py:396-402): write qpos/qvel, then mj_forward."""
        n, p = self.num_drones, 2 * int(bool(self.pendulum))
        dp = C.POINTER(C.c_double)
        qpos = np.ascontiguousarray(np.asarray(qpos, dtype=np.float64).reshape(n, 7 + p))
        qvel = np.ascontiguousarray(np.asarray(qvel, dtype=np.float64).reshape(n, 6 + p))
        a = None if act is None else np.ascontiguousarray(np.asarray(act, dtype=np.float64).reshape(n, 4))
        ns = None if num_steps is None else np.ascontiguousarray(num_steps, dtype=np.int32)
        self._ck(self._L.dsim_set_state(self._h, qpos.ctypes.data_as(dp), qvel.ctypes.data_as(dp),
                                        None if a is None else a.ctypes.data_as(dp),
                                        None if ns is None else ns.ctypes.data_as(C.POINTER(C.c_int32)), self._stream()))
        self._ck(self._L.dsim_forward(self._h, 0, self._stream()))
        self._states_cache = None; self._states_frozen = False

    # ------------------------------------------------------------------ RLlib VectorEnv surface
    def _obs_to_list(self):
        self._torch.cuda.current_stream(self._device).synchronize()
        self._last_obs = self._h_obs.numpy().astype(np.float64)
        return list(self._last_obs)

    def _fetch_obs(self):
        self._h_obs.copy_(self.obs_tensor, non_blocking=True)
        return self._obs_to_list()

    def vector_step(self, actions):
        """BaseDroneEnv.vector_step (:259-294): returns (obs list, rewards list, dones, truncated, infos)."""
        a = np.asarray(actions, dtype=np.float64)
        if a.size != 4 * self.num_drones:
            raise ValueError("Action dimension mismatch")                       # mujoco_env_custom.py:200-201
        if self._np_dtype == np.float32:
            # pinned staging buffers + the zero-copy host entry point: ONE kernel launch whose bulk loads / stores move the
            # actions and the outputs over PCIe, one stream synchronise - no separate copies
            self._np_actions[...] = a.reshape(self.num_drones, 4)
            self._ck(self._L.dsim_step_host(self._h, self._p_actions, self._p_obs, self._p_reward, self._p_trunc, self._stream()))
            self.total_steps += 1
            self._sensor_stale = False
            self._states_cache = None; self._states_frozen = False
            self._last_obs = self._np_obs.astype(np.float64)
            obs = list(self._last_obs)
        else:
            self._h_actions.copy_(self._torch.from_numpy(a.reshape(self.num_drones, 4)))
            self._d_actions.copy_(self._h_actions, non_blocking=True)
            self._ck(self._L.dsim_step(self._h, C.c_void_p(self._d_actions.data_ptr()), self._stream()))
            self.total_steps += 1
            self._sensor_stale = False
            self._states_cache = None; self._states_frozen = False
            self._h_reward.copy_(self.reward_tensor, non_blocking=True)
            self._h_trunc.copy_(self.truncated_tensor, non_blocking=True)
            obs = self._fetch_obs()
        rewards = list(self._h_reward.numpy().astype(np.float64))
        truncated = [bool(t) for t in self._h_trunc.numpy()]
        dones = [False] * self.num_drones                                        # Q7
        infos = [{} for _ in range(self.num_drones)]
        if self.random_params and self.regen_env_at_steps and self.total_steps == self.regen_env_at_steps:
            self.total_steps = 0                                                 # :289-292 (Q8): rewards pre-regen, obs post-regen
            obs = self.reset_model(regen=True)
            truncated = np.ones(self.num_drones, dtype=bool)
        return obs, rewards, dones, truncated, infos

    def reset_model(self, regen=False):
        """(:296-326) re-sample every drone state; regen=True also re-draws the drone parameters, recompiles the model
        constants and starts from a fresh MjData (act = 0)."""
        if regen:
            self._regen_epoch += 1
            self._ck(self._L.dsim_regen_params(self._h, self._regen_epoch, self._stream()))
            self._ck(self._L.dsim_zero_act(self._h, self._stream()))
        self._ck(self._L.dsim_reset_all(self._h, self._stream()))
        self._sensor_stale = False
        self._states_cache = None; self._states_frozen = False
        return self._fetch_obs()

    def vector_reset(self, seeds=None, options=None):
        obs = self.reset_model()
        infos = [{}] * self.num_drones
        return obs, infos

    def reset_at(self, index, seed=None, options=None):
        """(:334-351) re-sample one drone.  Returns the STALE observation like the reference (Q1): `self.states` is
        not refreshed there, so the caller sees the terminal observation again."""
        if index is None:
            index = 0
        assert index < self.num_drones
        if self._states_cache is None and not self._states_frozen:
            # freeze the stale rows before the state changes: computed on the device now (one launch, no sync), read back
            # only if somebody looks at `env.states` before the next vector_step
            if self._sensor_stale:
                self._ck(self._L.dsim_forward(self._h, 0, self._stream()))
                self._sensor_stale = False
            self._ck(self._L.dsim_compute_states(self._h, self._stream()))
            self._states_frozen = True
        self._ck(self._L.dsim_reset_at(self._h, int(index), self._stream()))
        self._sensor_stale = True          # reference: set_state -> mj_forward refreshes sensordata of ALL drones (Q2)
        if self._last_obs is None:
            self._fetch_obs()
        return self._last_obs[index], {}

    def reset(self, *, seed=None, options=None):
        obs, _ = self.vector_reset()
        return obs, {}

    def _get_obs(self):
        if self._last_obs is None:
            return self._fetch_obs()
        return list(self._last_obs)

    def get_drone_states(self):
        """(:357-380) list of per-drone 33-vectors (29 without pendulum), float64."""
        if self._states_frozen and self._states_cache is None:
            _ = self.states                  # the buffer still holds the rows frozen by reset_at: materialise them before it is overwritten
        if self._sensor_stale:
            self._ck(self._L.dsim_forward(self._h, 0, self._stream()))
            self._sensor_stale = False
        self._ck(self._L.dsim_compute_states(self._h, self._stream()))
        return list(self.tensor(_lib.BUF_STATES33).to('cpu').numpy().astype(np.float64))

    # viewer / mocap plumbing of the reference: accepted, headless no-ops
    def move_mocap_to(self, pose, idx=0):
        assert idx < self.mocaps

    def control_reference(self):
        raise NotImplementedError("joystick polling (pygame) is out of scope; use control_reference_tensor(axes)")

    def control_reference_tensor(self, axes):
        """(:151-172) per-env setpoint update from joystick-style axes: CUDA tensor [4, N] = (x, -y, -z, -yaw).
        A contiguous [4, ld] tensor (ld = npages * 32) of the env's dtype is consumed in place, anything else is staged."""
        r = self.reference_tensor
        ld = r.shape[1] * _lib.TILE
        if not (axes.is_cuda and axes.dtype == r.dtype and axes.is_contiguous() and tuple(axes.shape) == (4, ld)):
            buf = self._torch.zeros((4, ld), dtype=r.dtype, device=self._device)
            buf[:, :self.num_drones] = axes
            axes = buf
        self._ck(self._L.dsim_control_reference(self._h, C.c_void_p(axes.data_ptr()), self._stream()))
        self._states_cache = None; self._states_frozen = False

    def render(self, *a, **k):
        return None

    def viewer_setup(self):
        return None

    def close(self):
        h = getattr(self, '_h', None)
        if h is not None and h.value:
            self._views.clear()
            self._L.dsim_destroy(h)
            self._h = C.c_void_p()

    def guard_check(self):
        """envs created with DSIM_GUARD=1 in the environment: canary bytes around the device buffers overwritten so far"""
        return int(self._L.dsim_debug_guard_check(self._h))

    def launch_count(self):
        return int(self._L.dsim_launch_count(self._h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
