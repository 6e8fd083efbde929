"""Import the reference's own Python modules under stub packages.

TEST INFRASTRUCTURE, BUILD-CONTAINER ONLY: /root/reference does not exist on the GPU box, so nothing at
test / bench / smoke run time imports this file.  It is used by tools/make_golden.py to execute the
UNMODIFIED reference code (transformation, rewards, observation_wrappers, BaseDroneEnv methods) and
freeze its outputs into tests/golden/.  mujoco / dm_control / gymnasium / ray / glfw / pygame /
matplotlib are absent here, so they are replaced by empty stand-ins; the physics call
(`do_simulation` -> mujoco.mj_step) is pluggable.
"""
import sys
import types
from typing import Generic, TypeVar

import numpy as np

REFERENCE_ROOT = "/root/reference"


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    if "gymnasium" in sys.modules and getattr(sys.modules["gymnasium"], "_dsim_stub", False):
        return
    T = TypeVar("T")

    class Space(Generic[T]):
        pass

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float64, seed=None):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    class EzPickle:
        def __init__(self, *a, **k):
            pass

    class Env(Generic[T]):
        pass

    def np_random(seed=None):
        return np.random.default_rng(seed), seed

    seeding = _mod("gymnasium.utils.seeding", np_random=np_random)
    utils = _mod("gymnasium.utils", EzPickle=EzPickle, seeding=seeding)
    spaces = _mod("gymnasium.spaces", Box=Box, Space=Space)
    error = _mod("gymnasium.error", DependencyNotInstalled=type("DependencyNotInstalled", (Exception,), {}))
    logger = _mod("gymnasium.logger", warn=lambda *a, **k: None)
    rendering = _mod("gymnasium.envs.mujoco.mujoco_rendering", Viewer=type("Viewer", (), {}),
                     RenderContextOffscreen=type("RenderContextOffscreen", (), {}))
    gmj = _mod("gymnasium.envs.mujoco", mujoco_rendering=rendering)
    genvs = _mod("gymnasium.envs", mujoco=gmj)
    core = _mod("gymnasium.core", ObsType=TypeVar("ObsType"), ActType=TypeVar("ActType"),
                RenderFrame=TypeVar("RenderFrame"))
    registration = _mod("gymnasium.envs.registration", EnvSpec=type("EnvSpec", (), {}))
    g = _mod("gymnasium", utils=utils, spaces=spaces, error=error, logger=logger, envs=genvs, Env=Env,
             core=core, Space=Space, _dsim_stub=True)
    genvs.registration = registration
    g.spaces.Space = Space

    class VectorEnv:
        def __init__(self, observation_space, action_space, num_envs):
            self.observation_space, self.action_space, self.num_envs = observation_space, action_space, num_envs

    ve = _mod("ray.rllib.env.vector_env", VectorEnv=VectorEnv)
    renv = _mod("ray.rllib.env", vector_env=ve)
    rllib = _mod("ray.rllib", env=renv)
    _mod("ray", rllib=rllib)
    _mod("mujoco")
    _mod("glfw")
    _mod("pygame")
    _mod("dm_control", mjcf=types.SimpleNamespace(), mujoco=types.SimpleNamespace())
    _mod("matplotlib.colors", hsv_to_rgb=lambda x: x)
    if "matplotlib" not in sys.modules:
        _mod("matplotlib", colors=sys.modules["matplotlib.colors"])


def load_reference():
    """Returns the reference's `environments` modules (imported unmodified from /root/reference)."""
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    mods = {}
    for name in ("transformation", "rewards", "BaseDroneEnv", "observation_wrappers"):
        mods[name] = importlib.import_module("environments." + name)
    return types.SimpleNamespace(**mods)


def make_env(cls, config, physics=None, drone_params=None):
    """Build a reference env instance WITHOUT running __init__ (which needs MuJoCo): sets exactly the
    attributes BaseDroneEnv.__init__ (BaseDroneEnv.py:60-114) would set, then plugs a physics backend.
    `physics(env, ctrl, n_frames)` replaces extendedEnv.do_simulation."""
    ref = load_reference()
    B = ref.BaseDroneEnv
    env = object.__new__(cls)
    g = config.get
    env.controlled = False
    env.render_mode = None
    env.mocaps = g('mocaps', 1)
    env.skip_steps = g('skip_steps', 1)
    env.frame_skip = env.skip_steps
    env.frequency = g('frequency', 200)
    env.reference = g('reference', [0, 0, 0, 0])
    env.num_drones = g('num_drones', 1)
    env.pendulum = g('pendulum', True)
    env.mass_interval = np.array(g('mass_interval', [1.35, 0.15]))
    env.arm_len_interval = np.array(g('arm_len_interval', [0.17, 0.02]))
    env.motor_force_interval = np.array(g('motor_force_interval', [7.5, 1.5]))
    env.motor_tau_interval = np.array(g('motor_tau_interval', [0.003, 0.002]))
    env.pendulum_length_interval = np.array(g('pendulum_length_interval', [1.2, 0.3]))
    env.weight_mass_interval = np.array(g('weight_mass_interval', [0.2, 0.1]))
    env.state_difficulty = g('state_difficulty', 0.1)
    env.param_difficulty = g('param_difficulty', 0.1)
    env.random_start_pos = g('random_start_pos', False)
    env.random_params = g('random_params', False)
    env.regen_env_at_steps = g('regen_env_at_steps', None)
    env.start_pos = g('start_pos', env.reference)
    env.max_distance = g('max_distance', 1)
    env.reward_fcn = g('reward_fcn', ref.rewards.default_reward_fcn)
    env.terminated_fcn = g('terminated_fcn', B.default_termination_fcn)
    env.max_steps = g('max_steps', 512)
    env.max_pos_offset = env.state_difficulty * g('max_random_offset', 0)
    env.angle_variance = env.state_difficulty * np.array(g('angle_variance', [0, 0]))
    env.ang_vel_variance = env.state_difficulty * np.array(g('ang_vel_variance', [0, 0, 0]))
    env.vel_variance = env.state_difficulty * np.array(g('vel_variance', [0, 0, 0]))
    env.pendulum_rp_variance = env.state_difficulty * np.array(g('pendulum_rp_variance', [0, 0]))
    env.pendulum_ang_vel_variance = env.state_difficulty * np.array(g('pendulum_ang_vel_variance', [0, 0]))
    env.total_steps = 0
    env.num_steps = np.zeros((env.num_drones,), dtype=np.int64)
    env.np_random = np.random.default_rng(g('worker_index', -1) + 1 + g('seed', 1))
    env.drone_params = drone_params if drone_params is not None else env.generate_drone_params()
    env.num_params = len(env.drone_params[0])
    n, p = env.num_drones, 2 * int(bool(env.pendulum))
    env.data = types.SimpleNamespace(
        qpos=np.tile(np.concatenate(([0, 0, 0.15, 1, 0, 0, 0], np.zeros(p))), n),
        qvel=np.zeros((6 + p) * n), sensordata=np.zeros(3 * n), act=np.zeros(4 * n), ctrl=np.zeros(4 * n),
        mocap_pos=np.zeros((env.mocaps, 3)), mocap_quat=np.zeros((env.mocaps, 4)))
    env.init_qpos = env.data.qpos.copy()
    env.init_qvel = env.data.qvel.copy()
    env.width, env.height = 640, 480
    env.observation_space = None
    env.action_space = None
    env._physics = physics
    env._forward = None

    def do_simulation(ctrl, n_frames):
        if np.array(ctrl).shape != (4 * env.num_drones,):
            raise ValueError("Action dimension mismatch")       # mujoco_env_custom.py:200-201
        env.data.ctrl[:] = ctrl
        if env._physics is not None:
            env._physics(env, ctrl, n_frames)

    def set_state(qpos, qvel):
        env.data.qpos[:] = np.copy(qpos)                           # mujoco_vecenv.py:396-402
        env.data.qvel[:] = np.copy(qvel)
        if env._forward is not None:
            env._forward(env)

    env.do_simulation = do_simulation
    env.set_state = set_state
    env.render = lambda *a, **k: None
    env.close = lambda *a, **k: None
    env.states = env.get_drone_states()
    return env
