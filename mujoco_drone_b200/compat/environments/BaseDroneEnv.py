"""Same import surface as the reference's environments/BaseDroneEnv.py (:12-16 default_termination_fcn, :19-50
base_config, :53 BaseDroneEnv), backed by libdronesim_b200.so."""
from mujoco_drone_b200.env import BaseDroneEnv, base_config, default_termination_fcn   # noqa: F401
from mujoco_drone_b200.rewards import default_reward_fcn                                # noqa: F401
