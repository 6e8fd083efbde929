"""Import shims with the reference's module paths.  Put this directory FIRST on sys.path (or PYTHONPATH) and the
reference's scripts (`train_RMA.py`, `train_PPO.py`, `rollout.py`, `evaluation.py`) resolve
`environments.BaseDroneEnv`, `environments.observation_wrappers` and `environments.rewards` to the CUDA-backed
classes without a source change:

    PYTHONPATH=/path/to/repo/mujoco_drone_b200/compat:/path/to/repo python train_RMA.py
"""
