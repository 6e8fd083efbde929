"""Same function names as the reference's environments/rewards.py (:5-368); the callables are device tokens that
`config['reward_fcn']` resolves to the reward id evaluated inside the fused CUDA step kernel."""
from mujoco_drone_b200.rewards import *                       # noqa: F401,F403
from mujoco_drone_b200.rewards import REWARD_IDS, REWARD_NAMES   # noqa: F401
