"""Setpoint / trajectory generators.

Host generators with the reference's names and return values (evaluation.py:135-152), and `TrajectoryReference`, which
evaluates the same trajectories ON THE DEVICE for every env of a batch (env i at time t + i * phase_step), writing the
env's per-env setpoint page so that a moving-reference rollout (BASELINE config 3) needs no host traffic."""
import ctypes as C

import numpy as np

from . import _lib


def gen_circle_trajectory(T=10, f=0.5, r=1, h=1):                                                       # evaluation.py:135-138
    t = np.arange(0, T, 0.01)
    return t, np.array([r * np.cos(2 * np.pi * f * t), r * np.sin(2 * np.pi * f * t), h * np.ones_like(t), np.zeros_like(t)]).T


def gen_step_trajectory(step_time=5, duration=10, start_pos=(0, 0, 0, 0), end_pos=(0, 0, 1, 0)):        # :141-144
    t = np.arange(0, duration, 0.01)
    return t, np.array([list(start_pos) if i < step_time else list(end_pos) for i in t], dtype=np.float64)


def gen_ramp_trajectory(start_time=5, duration=10, start_pos=(0, 0, 0, 0), end_pos=(0, 0, 1, 0)):       # :147-152
    t = np.arange(0, duration, 0.01)
    s, e = np.array(start_pos, dtype=np.float64), np.array(end_pos, dtype=np.float64)
    return t, np.array([s if i < start_time else (s + (i - start_time) / (duration - start_time) * (e - s)) for i in t])


class TrajectoryReference:
    """kind: 'circle' (f, r, h) | 'step' (step_time, start_pos, end_pos) | 'ramp' (start_time, duration, start_pos, end_pos).
    `advance(t)` sets every env's reference to the trajectory at time t + i * phase_step (one tiny kernel, no host data)."""

    def __init__(self, env, kind, phase_step=0.0, **kw):
        if not env.per_env_reference:
            raise ValueError("TrajectoryReference needs an env created with per_env_reference=True")
        self.env, self.phase_step = env, float(phase_step)
        z = (0.0, 0.0, 0.0, 0.0)
        if kind == "circle":
            self.kind, self.params = _lib.TRAJ_CIRCLE, (kw.get("f", 0.5), kw.get("r", 1), kw.get("h", 1))
            self.start, self.end = z, z
        elif kind == "step":
            self.kind, self.params = _lib.TRAJ_STEP, (kw.get("step_time", 5), 0.0, 0.0)
            self.start, self.end = tuple(kw.get("start_pos", z)), tuple(kw.get("end_pos", (0, 0, 1, 0)))
        elif kind == "ramp":
            self.kind, self.params = _lib.TRAJ_RAMP, (kw.get("start_time", 5), kw.get("duration", 10), 0.0)
            self.start, self.end = tuple(kw.get("start_pos", z)), tuple(kw.get("end_pos", (0, 0, 1, 0)))
        else:
            raise ValueError(f"unknown trajectory kind {kind!r}")

    def advance(self, t):
        d3, d4 = (C.c_double * 3)(*[float(x) for x in self.params]), C.c_double * 4
        e = self.env
        e._ck(e._L.dsim_trajectory_reference(e._h, self.kind, float(t), self.phase_step, d3, d4(*[float(x) for x in self.start]),
                                             d4(*[float(x) for x in self.end]), e._stream()))
        e._states_cache = None
