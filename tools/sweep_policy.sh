#!/bin/bash
# policy-kernel variant sweep on the GPU box: fused-kernel time and max error vs torch fp32 per prebuilt library variant
for lib in mujoco_drone_b200/libdronesim_b200.so mujoco_drone_b200/variants/*.so; do
  echo -n "$lib: "
  DSIM_LIB=$PWD/$lib python - <<'PY'
import torch, mujoco_drone_b200 as M
torch.manual_seed(0)
n = 524288
model = M.policy.make_rma_full().cuda()
fused = M.policy.FusedRMAFull(model)
obs, prev = torch.randn((n, 22), device="cuda"), torch.rand((n, 4), device="cuda")
lg, val = torch.empty((n, 8), device="cuda"), torch.empty((n,), device="cuda")
for _ in range(5): fused(obs, prev, logits_out=lg, value_out=val)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30): fused(obs, prev, logits_out=lg, value_out=val)
e1.record(); torch.cuda.synchronize(); fused.check()
with torch.no_grad(): rl, rv = model(obs[:65536], prev[:65536])
print(f"{e0.elapsed_time(e1) / 30 * 1e3:7.1f} us   max|dlogit| {(lg[:65536] - rl).abs().max().item():.2e}  max|dvalue| {(val[:65536] - rv).abs().max().item():.2e}  rms dlogit {(lg[:65536] - rl).pow(2).mean().sqrt().item():.2e}")
PY
done
