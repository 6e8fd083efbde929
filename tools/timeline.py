#!/usr/bin/env python
"""Debug: per-warp %globaltimer timeline of one step launch (DSIM_TIMELINE=1).  usage: timeline.py [workload] [cold|hot]
Note: with the timeline enabled the library launches the GENERIC instantiation of the step kernel (the specialised ones
have the stamps compiled out), which executes ~10 % more instructions per page than the kernel the bench times."""
import ctypes as C
import os
import sys
os.environ["DSIM_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c4"]
mode = sys.argv[2] if len(sys.argv) > 2 else "cold"
n = wl["envs_per_gpu"]
env = bench.make_env(wl, n, 0, 0)
env.reset_tensor()
a = torch.rand((n, 4), device="cuda")
flush = torch.zeros(256 * 1024 * 1024 // 4, device="cuda")
pre = int(sys.argv[3]) if len(sys.argv) > 3 else 1500
for i in range(pre):
    env.step_tensor(a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(200):
    env.step_tensor(a)
e1.record()
torch.cuda.synchronize()
print(f"back-to-back period (hot L2, {pre} pre-roll steps): {e0.elapsed_time(e1) * 1e3 / 200:.2f} us/step")
res = []
for rep in range(5):
    if mode == "cold":
        flush.sum()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.step_tensor(a); e1.record()
    torch.cuda.synchronize()
    npages = (n + 31) // 32
    buf = np.zeros((npages, 8), dtype=np.uint64)
    env._ck(env._L.dsim_debug_timeline(env._h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), buf.size))
    t = buf.astype(np.int64)
    act = t[:, 0] > 0
    t = t[act]
    t0 = t[:, 0].min()
    two = t[:, 4] > 0
    three = t[:, 6] > 0
    def st(x):
        q = np.percentile(x - t0, [0, 50, 90, 99, 100]).astype(int)
        return "min %6d med %6d p90 %6d p99 %6d max %6d" % tuple(q)
    print(f"--- rep {rep} ({mode}) event {e0.elapsed_time(e1) * 1e3:.1f} us; warps {len(t)} (two pages: {two.sum()}); ns since first CTA entry:")
    print("entry        ", st(t[:, 0]))
    print("dep wait done", st(t[:, 1]))
    print("page1 landed ", st(t[:, 2]))
    print("page1 publ.  ", st(t[:, 3]))
    if two.any():
        print("page2 landed ", st(t[two, 4]))
        print("page2 publ.  ", st(t[two, 5]))
    print("exit         ", st(t[:, 7]), f" (warps with >=3 pages: {int(three.sum())})")
    print("per-warp: load wait med", int(np.median(t[:, 2] - t[:, 1])), " page1 compute med", int(np.median(t[:, 3] - t[:, 2])),
          " page2 compute med", int(np.median(t[two, 5] - t[two, 4])) if two.any() else 0, " exit wait med", int(np.median(t[:, 7] - np.where(two, t[:, 5], t[:, 3]))))
env.close()
