#!/bin/bash
python /tmp/p32.py 2>/dev/null || { cat > /tmp/p32.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch, mujoco_drone_b200 as M
pol = M.policy.make_rma_full().cuda()
n = 524288
obs, prev = torch.randn((n, 22), device="cuda"), torch.rand((n, 4), device="cuda")
net = M.policy.FP32RMAFull(pol, device=0)
for _ in range(3): net(obs, prev)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): net(obs, prev)
e1.record(); torch.cuda.synchronize()
print("fp32 policy forward, 524288 rows: %.1f us" % (e0.elapsed_time(e1) * 1e3 / 5))
with torch.no_grad():
    rl, rv = pol(obs[:4096], prev[:4096])
lg, v = net(obs[:4096], prev[:4096])
print("max err vs torch fp32:", (lg - rl).abs().max().item(), (v - rv).abs().max().item())
PY
python /tmp/p32.py; }
python -m pytest tests/test_policy_reference.py -m gpu -q 2>&1 | tail -2
