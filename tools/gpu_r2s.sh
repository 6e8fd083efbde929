#!/bin/bash
mkdir -p gpurun_out
python tools/bench_policy.py 524288 > gpurun_out/r2s_bench_policy.txt 2>&1; cat gpurun_out/r2s_bench_policy.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rma_full_forward_kernel|rma_full_forward_fp32_kernel|beta_policy_kernel' -s 6 -c 1 -o gpurun_out/prof_r02_policy_tc -f python tools/bench_policy.py 524288 > gpurun_out/ncu_r02_policy_tc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fp32_kernel' -s 2 -c 1 -o gpurun_out/prof_r02_policy_fp32 -f python tools/bench_policy.py 524288 > gpurun_out/ncu_r02_policy_fp32.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'beta_policy_kernel' -s 2 -c 1 -o gpurun_out/prof_r02_policy_beta -f python tools/bench_policy.py 524288 > gpurun_out/ncu_r02_policy_beta.log 2>&1
ls -la gpurun_out/prof_r02_policy*
