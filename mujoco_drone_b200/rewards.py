"""Reward registry: same function names as the reference's environments/rewards.py.

A config may carry either these objects or the reference's own functions under `config['reward_fcn']`
(BaseDroneEnv.py:98); both resolve BY NAME to the reward id evaluated inside the fused CUDA step kernel
(mujoco_drone_b200/csrc/dsim_obs_reward.cuh).  They are device-side rewards: calling one on the host raises,
there is no CPU implementation in the product.
"""

REWARD_NAMES = [
    "default_reward_fcn",                       # rewards.py:5-10
    "distance_reward_fcn",                      # :13-20
    "distance_energy_reward",                   # :23-31
    "distance_energy_reward_pendulum_angle",    # :34-43
    "distance_energy_reward_pendulum_angle2",   # :46-56
    "distance_energy_reward_pendulum_angle3",   # :59-72
    "distance_energy_reward_pendulum_en",       # :75-107
    "distance_energy_reward_pendulum_en2",      # :110-146
    "distance_energy_reward_pendulum_en3",      # :149-188
    "distance_energy_reward_pendulum_en4",      # :191-230
    "distance_time_energy_reward",              # :233-242
    "reward_1",                                 # :245-257
    "reward_pendulum_dist",                     # :283-294
    "reward_pendulumDistHeading",               # :297-310
    "reward_2",                                 # :313-327
    "reward_2_penergy",                         # :330-348
    "reward_3",                                 # :351-368
]
REWARD_IDS = {n: i for i, n in enumerate(REWARD_NAMES)}


class DeviceReward:
    """Token for a reward evaluated on the GPU; `(env, state, action, num_steps)` signature kept for documentation."""

    def __init__(self, name):
        self.__name__ = name
        self.reward_id = REWARD_IDS[name]

    def __call__(self, env, state, action, num_steps):
        raise NotImplementedError(
            f"{self.__name__} is evaluated inside the fused CUDA step kernel (reward_id={self.reward_id}); "
            "there is no host implementation in mujoco_drone_b200")

    def __repr__(self):
        return f"<device reward {self.__name__} id={self.reward_id}>"


for _n in REWARD_NAMES:
    globals()[_n] = DeviceReward(_n)


def resolve_reward(fcn):
    """config['reward_fcn'] -> reward id; unknown callables raise (no CPU fallback)."""
    if fcn is None:
        return 0
    if isinstance(fcn, int):
        if 0 <= fcn < len(REWARD_NAMES):
            return fcn
        raise NotImplementedError(f"unknown reward id {fcn}")
    name = fcn if isinstance(fcn, str) else getattr(fcn, "__name__", None)
    if name in REWARD_IDS:
        return REWARD_IDS[name]
    raise NotImplementedError(
        f"reward_fcn {fcn!r} is not one of the reference's rewards.py functions; arbitrary Python rewards cannot run "
        "inside the CUDA step kernel")
