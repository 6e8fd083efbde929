"""The 14 observation variants of the reference's environments/observation_wrappers.py, same class names.

Each class only selects which `_get_obs` layout the fused CUDA step kernel emits (csrc/dsim_obs_reward.cuh
`emit_obs`) and declares the observation_space the reference class declares.  Known reference quirks are kept:
`LocalFrameFullStateZvecEnv` declares 23 values but emits 24 (observation_wrappers.py:121,149);
`LocalFramePRYaccParamsNoPendEnv._get_obs` raises NameError (`acc` is commented out at :438, used at :448).
"""
from .env import BaseDroneEnv


class GlobalFrameRPYEnv(BaseDroneEnv):                # observation_wrappers.py:7-35
    OBS_ID, DECLARED_OBS = 1, 16


class LocalFramePRYEnv(BaseDroneEnv):                 # :38-73
    OBS_ID, DECLARED_OBS = 2, 16


class LocalFrameFullStateEnv(BaseDroneEnv):           # :76-111
    OBS_ID, DECLARED_OBS = 3, 23


class LocalFrameFullStateZvecEnv(BaseDroneEnv):       # :114-151 (emits 24)
    OBS_ID, DECLARED_OBS = 4, 23


class LocalFramePRYaccEnv(BaseDroneEnv):              # :154-191
    OBS_ID, DECLARED_OBS = 5, 19


class LocalFramePRYParamsEnv(BaseDroneEnv):           # :194-230
    OBS_ID, DECLARED_OBS = 6, 22


class LocalFramePRYaccParamsEnv(BaseDroneEnv):        # :233-265
    OBS_ID, DECLARED_OBS = 7, 25


class LocalFrameRPYParamsEnv(BaseDroneEnv):           # :268-304
    OBS_ID, DECLARED_OBS = 8, 22


class LocalFrameRPYFakeParamsEnv(BaseDroneEnv):       # :307-344
    OBS_ID, DECLARED_OBS = 9, 22


class LocalFrameRPYEnv(BaseDroneEnv):                 # :347-382
    OBS_ID, DECLARED_OBS = 10, 16


class LocalFramePRYaccNoPendEnv(BaseDroneEnv):        # :385-416
    OBS_ID, DECLARED_OBS = 11, 15


class LocalFramePRYaccParamsNoPendEnv(BaseDroneEnv):  # :419-450
    OBS_ID, DECLARED_OBS = 11, 21

    def _fetch_obs(self):
        raise NameError("name 'acc' is not defined")  # what the reference raises at observation_wrappers.py:448


class LocalFrameRmParamsEnv(BaseDroneEnv):            # :453-489
    OBS_ID, DECLARED_OBS = 13, 28


class LocalFrameZvecEnv(BaseDroneEnv):                # :492-529
    OBS_ID, DECLARED_OBS = 14, 17


WRAPPERS = {c.__name__: c for c in (
    BaseDroneEnv, GlobalFrameRPYEnv, LocalFramePRYEnv, LocalFrameFullStateEnv, LocalFrameFullStateZvecEnv,
    LocalFramePRYaccEnv, LocalFramePRYParamsEnv, LocalFramePRYaccParamsEnv, LocalFrameRPYParamsEnv,
    LocalFrameRPYFakeParamsEnv, LocalFrameRPYEnv, LocalFramePRYaccNoPendEnv, LocalFramePRYaccParamsNoPendEnv,
    LocalFrameRmParamsEnv, LocalFrameZvecEnv)}
