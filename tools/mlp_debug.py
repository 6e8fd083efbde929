import os, sys, ctypes as C
os.environ["DSIM_MLP_DEBUG"] = "1"
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import mujoco_drone_b200 as M
model = M.policy.make_rma_full().cuda()
f = M.policy.FusedRMAFull(model)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
obs, prev = torch.randn((n, 22), device="cuda"), torch.rand((n, 4), device="cuda")
for _ in range(3):
    f(obs, prev)
out = np.zeros(64, dtype=np.int64)
f._L.dsim_policy_debug.argtypes = [C.c_void_p, C.c_void_p]
print("rc", f._L.dsim_policy_debug(f._h, out.ctypes.data))
for wg in range(2):
    t = out[wg * 32: wg * 32 + 32]
    t = t[t > 0]
    print("wg", wg, "deltas(cycles):", np.diff(t).tolist(), "total first tile", t[10] - t[0] if len(t) > 10 else None)
