#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/p32.py <<'PY'
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
PY
python /tmp/p32.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fp32_kernel -s 1 -c 1 -o gpurun_out/prof_r2p_fp32 -f python /tmp/p32.py > gpurun_out/ncu_r2p.log 2>&1; tail -2 gpurun_out/ncu_r2p.log
