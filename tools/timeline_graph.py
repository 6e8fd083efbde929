#!/usr/bin/env python
"""Debug: %globaltimer timeline of CONSECUTIVE step launches in the bench's steady state (R replicas round-robin, CUDA graph
replay), from a library built with -DDSIM_TL_ALL (python -m mujoco_drone_b200.build --variant tl DSIM_TL_ALL) so that the
specialised kernel the bench times carries the stamps.  Every handle keeps the stamps of its LAST launch; after one
replay of an R-step graph the R buffers hold R consecutive kernels on one clock.

usage: DSIM_LIB=$PWD/mujoco_drone_b200/variants/tl.so python tools/timeline_graph.py [workload] [envs]
"""
import ctypes as C
import os
import sys
os.environ["DSIM_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench

wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c4"])
n = int(sys.argv[2]) if len(sys.argv) > 2 else wl["envs_per_gpu"]
R = 8
envs = [bench.make_env(wl, n, r * n, 0) for r in range(R)]
READY = os.environ.get("DSIM_INPUTS_READY", "0") == "1"
for e in envs:
    e.inputs_ready = READY
    e.reset_tensor()
print("inputs_ready =", READY)
bank = torch.rand((9, n, 4), device="cuda")
for i in range(300 * R):
    envs[i % R].step_tensor(bank[i % 9])
torch.cuda.synchronize()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for i in range(R):
        envs[i].step_tensor(bank[i % 9])
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(2 * R):
        envs[i % R].step_tensor(bank[i % 9])
for _ in range(200):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    g.replay()
e1.record()
torch.cuda.synchronize()
print(f"graph replay period: {e0.elapsed_time(e1) * 1e3 / (50 * 2 * R):.2f} us/step ({n} envs, {R} replicas, timeline stamps compiled in)")
npages = (n + 31) // 32
T = []
for e in envs:
    buf = np.zeros((npages, 8), dtype=np.uint64)
    e._ck(e._L.dsim_debug_timeline(e._h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), buf.size))
    t = buf.astype(np.int64)
    T.append(t[t[:, 0] > 0])
t00 = min(t[:, 0].min() for t in T)


def q(x):
    return "min %6d p10 %6d med %6d p90 %6d max %6d" % tuple(np.percentile(x, [0, 10, 50, 90, 100]).astype(int))


prev_exit = None
for k, t in enumerate(T):
    two = t[:, 4] > 0
    base = t[:, 0].min()
    print(f"--- kernel {k}: first CTA entry at {base - t00} ns; {len(t)} warps ({int(two.sum())} with a second page)"
          + (f"; starts {base - prev_exit:+d} ns relative to the previous kernel's LAST exit; period {base - prev_base} ns" if prev_exit else ""))
    print("  entry         ", q(t[:, 0] - base))
    print("  dep wait done ", q(t[:, 1] - base))
    print("  page1 landed  ", q(t[:, 2] - base))
    print("  page1 publ.   ", q(t[:, 3] - base))
    if two.any():
        print("  page2 landed  ", q(t[two, 4] - base))
        print("  page2 publ.   ", q(t[two, 5] - base))
    print("  exit          ", q(t[:, 7] - base))
    print("  per warp: entry->dep %5d  dep->landed %5d  page1 compute %5d  page2 compute %5d" % (
        np.median(t[:, 1] - t[:, 0]), np.median(t[:, 2] - t[:, 1]), np.median(t[:, 3] - t[:, 2]), np.median(t[two, 5] - t[two, 4]) if two.any() else 0))
    if prev_exit:
        print(f"  previous kernel's last exit -> this kernel's median dep-wait-done: {int(np.median(t[:, 1])) - prev_exit} ns; -> median page1 landed: {int(np.median(t[:, 2])) - prev_exit} ns")
    prev_exit, prev_base = int(t[:, 7].max()), base
for e in envs:
    e.close()
