#!/bin/bash
# ncu capture of the fused tcgen05 policy kernel and the Beta sampling kernel.  usage: tools/profile_policy.sh <tag>
set -e
TAG=${1:-r01p}
CMD="python tools/bench_policy.py 524288"
$CMD > gpurun_out/plain_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'rma_full_forward_kernel|beta_policy_kernel' -s 6 -c 2 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log | cut -c1-200
