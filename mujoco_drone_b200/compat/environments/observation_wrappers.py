"""Same class names as the reference's environments/observation_wrappers.py (:7-529)."""
from mujoco_drone_b200.observation_wrappers import *          # noqa: F401,F403
from mujoco_drone_b200.observation_wrappers import WRAPPERS   # noqa: F401
