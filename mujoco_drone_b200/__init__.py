"""mujoco_drone_b200 — B200-native batched simulator for the vectorised env-step path of TichyTech/mujoco-drone.

The package holds only what that path needs: `csrc/` (CUDA kernels + the C ABI of libdronesim_b200.so), the
ctypes binding, and the host-side mirror of the reference's RLlib VectorEnv interface
(`BaseDroneEnv`, the observation wrappers, the reward registry).  Importing the package does not need a GPU;
constructing an env does (there is no CPU fallback).
"""
from . import _lib, rewards                                     # noqa: F401
from .env import BaseDroneEnv, base_config, default_termination_fcn   # noqa: F401
from . import observation_wrappers                                # noqa: F401
from . import policy, rollout, trajectories                       # noqa: F401

__all__ = ["BaseDroneEnv", "base_config", "default_termination_fcn", "observation_wrappers", "rewards"]
