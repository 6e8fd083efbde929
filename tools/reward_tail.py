#!/usr/bin/env python
"""Debug: per-window reward / episode-return statistics under random actions (workload c4): are they stationary?"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c4"]
n = wl["envs_per_gpu"]
env = bench.make_env(wl, n, 0, 0)
env.reset_tensor()
g = torch.Generator(device="cuda").manual_seed(0)
acc = torch.zeros((), device="cuda", dtype=torch.float64)
mn = torch.zeros((), device="cuda")
for t in range(1, 1801):
    a = torch.rand((n, 4), device="cuda", generator=g)
    obs, rew, trunc = env.step_tensor(a)
    acc += rew.double().sum()
    mn = torch.minimum(mn, rew.min())
    if t % 200 == 0:
        st = env.episode_stats(reset=True)
        ne = max(st["n_episodes"], 1)
        print(f"steps {t - 199:5d}-{t:5d}: mean reward/step {acc.item() / (200 * n):9.4f}  min reward {mn.item():10.3f}  episodes {int(st['n_episodes']):8d} "
              f"mean return {st['sum_return'] / ne:10.3f} mean length {st['sum_length'] / ne:7.2f}")
        acc.zero_(); mn.zero_()
