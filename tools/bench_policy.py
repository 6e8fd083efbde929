#!/usr/bin/env python
"""Micro-benchmark of the policy side of config 5: fused tcgen05 RMA_full kernel vs torch (fp32 / tf32 / bf16 autocast),
and the MyBetaDist sampling kernel.  usage: bench_policy.py [n]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mujoco_drone_b200 as M

n = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
model = M.policy.make_rma_full().cuda()
fused = M.policy.FusedRMAFull(model)
obs, prev = torch.randn((n, 22), device="cuda"), torch.rand((n, 4), device="cuda")
lg, val = torch.empty((n, 8), device="cuda"), torch.empty((n,), device="cuda")


def timeit(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


flops = 2 * n * (6 * 32 + 32 * 8 + 28 * 256 + 256 * 128 + 3 * 128 * 128 + 128 * 8 + 128)
t = timeit(lambda: fused(obs, prev, logits_out=lg, value_out=val))
fused.check()
print(f"fused tcgen05     {t:9.1f} us  {flops / t / 1e6:8.1f} TFLOP/s (useful)  {n / t:8.1f} M rows/s   (bf16 operands)")
f32 = M.policy.FP32RMAFull(model)
t = timeit(lambda: f32(obs, prev, logits_out=lg, value_out=val), 10)
print(f"fused FP32 pipe   {t:9.1f} us  {flops / t / 1e6:8.1f} TFLOP/s (useful)  {n / t:8.1f} M rows/s   (FP32-faithful)")
fused(obs, prev, logits_out=lg, value_out=val)
with torch.no_grad():
    t = timeit(lambda: model(obs, prev), 10)
    print(f"torch fp32        {t:9.1f} us  {flops / t / 1e6:8.1f} TFLOP/s")
    torch.backends.cuda.matmul.allow_tf32 = True
    t = timeit(lambda: model(obs, prev), 10)
    print(f"torch tf32        {t:9.1f} us  {flops / t / 1e6:8.1f} TFLOP/s")

    def bf():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return model(obs, prev)
    t = timeit(bf, 10)
    print(f"torch bf16        {t:9.1f} us  {flops / t / 1e6:8.1f} TFLOP/s")
act, lp = torch.empty((n, 4), device="cuda"), torch.empty((n,), device="cuda")
t = timeit(lambda: M.policy.beta_policy(lg, 1, 0, 0, actions_out=act, logp_out=lp))
print(f"beta sample+logp  {t:9.1f} us")
t = timeit(lambda: M.policy.beta_policy(lg, 1, 0, 0, actions_out=act, want_logp=False))
print(f"beta sample only  {t:9.1f} us")
