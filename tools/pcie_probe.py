#!/usr/bin/env python
"""Debug: raw pinned-memory copy bandwidth of the box next to dsim_step_host's end-to-end rate (is e2e at the PCIe ceiling?)."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0)
for mb in (1.5, 12, 64):
    nbytes = int(mb * 1e6)
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name} {mb:5.1f} MB: {nbytes * 20 / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s")
wl = bench.WORKLOADS["c4"]
n = wl["envs_per_gpu"]
env = bench.make_env(wl, n, 0, 0)
env.reset_tensor()
h_act = torch.rand((n, 4)).pin_memory()
h_obs, h_rew, h_tr = torch.empty((n, env.obs_dim)).pin_memory(), torch.empty((n,)).pin_memory(), torch.empty((n,), dtype=torch.uint8).pin_memory()
for _ in range(5):
    env.step_host(h_act.numpy(), h_obs.numpy(), h_rew.numpy(), h_tr.numpy())
t = time.perf_counter()
for _ in range(50):
    env.step_host(h_act.numpy(), h_obs.numpy(), h_rew.numpy(), h_tr.numpy())
dt = (time.perf_counter() - t) / 50
byt = h_obs.numel() * 4 + h_rew.numel() * 4 + h_tr.numel()
print(f"step_host: {dt * 1e6:.1f} us/step = {n / dt:.3e} env-steps/s; D2H {byt / 1e6:.2f} MB -> {byt / dt / 1e9:.1f} GB/s effective")
# floor: the same read-back bytes in the same chunking as bare copies on one stream + one synchronize (wall clock)
fr = [0, 1, 2, 4, 8, 16]
d_obs = env.obs_tensor
d_rew, d_tr = env.reward_tensor, env.truncated_tensor
def bare():
    for k in range(5):
        a, b = n * fr[k] // 16, n * fr[k + 1] // 16
        h_obs[a:b].copy_(d_obs[a:b], non_blocking=True)
    h_rew.copy_(d_rew, non_blocking=True)
    h_tr.copy_(d_tr, non_blocking=True)
    torch.cuda.synchronize()
for _ in range(5):
    bare()
t = time.perf_counter()
for _ in range(50):
    bare()
dt2 = (time.perf_counter() - t) / 50
print(f"bare read-back (5 + 2 copies, one sync): {dt2 * 1e6:.1f} us = {byt / dt2 / 1e9:.1f} GB/s")
def one():
    h_obs.copy_(d_obs, non_blocking=True)
    torch.cuda.synchronize()
for _ in range(5):
    one()
t = time.perf_counter()
for _ in range(50):
    one()
dt3 = (time.perf_counter() - t) / 50
print(f"bare read-back (observations only, one copy): {dt3 * 1e6:.1f} us = {h_obs.numel() * 4 / dt3 / 1e9:.1f} GB/s")
