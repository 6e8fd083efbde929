"""Multi-GPU sharding of the vectorised env (one process per GPU, torch.distributed).

The path shards trivially: every drone is an independent kinematic subtree (all drone geoms have conaffinity=0,
environments/env_gen.py:17-18, so drone-drone contact is impossible) and the reference already runs its rollout
workers as independent replicas (train_RMA.py:101).  Envs are split into contiguous GLOBAL id ranges; Philox streams
are keyed by the global id, so a 1-GPU and an 8-GPU run of the same config draw identical resets and parameters.
There is NO data-path collective.  The only exchange is the episode-statistics all-reduce per report interval
(NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_range(num_envs_total, rank, world_size):
    """Contiguous global env-id range [lo, hi) owned by `rank`; ragged totals put the remainder on the low ranks."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, rem = divmod(int(num_envs_total), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_config(config, rank, world_size, device=None):
    """Per-rank copy of a reference-style config: `num_drones` becomes the shard size, `env_id_offset` its first
    global id, `device` the local GPU."""
    lo, hi = shard_range(config.get('num_drones', 1), rank, world_size)
    cfg = dict(config)
    cfg['num_drones'] = hi - lo
    cfg['env_id_offset'] = int(config.get('env_id_offset', 0)) + lo
    if device is not None:
        cfg['device'] = device
    return cfg


STAT_KEYS = ("sum_return", "sum_length", "n_episodes", "n_nonfinite", "n_near_ground")


def allreduce_episode_stats(stats, device=None, group=None):
    """Sum the per-rank episode statistics (dict from BaseDroneEnv.episode_stats) over all ranks: the one collective
    of the path.  Works with backend nccl (pass the CUDA device) or gloo (CPU)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(stats[k]) for k in STAT_KEYS], dtype=torch.float64, device=device if device is not None else 'cpu')
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    out = dict(zip(STAT_KEYS, t.tolist()))
    n = max(out["n_episodes"], 1.0)
    out["mean_return"] = out["sum_return"] / n
    out["mean_length"] = out["sum_length"] / n
    return out


def broadcast_reference(reference, src=0, device=None, group=None):
    """A shared moving setpoint (control_reference, BaseDroneEnv.py:151-172) is a 4-float broadcast from rank `src`."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(np.asarray(reference, dtype=np.float64), device=device if device is not None else 'cpu')
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(t, src=src, group=group)
    return t.cpu().numpy()
