#!/bin/bash
# GPU-box profiling pass (B200_PROFILING.md recipe): plain run first, then the launch list (the first $SKIP launches are
# the one-off setup of the 8 replicas: parameter draw, model compile, reset), then one full capture of the step kernel.  usage: tools/profile_step.sh <tag> [workload]
set -e
TAG=${1:-r01}; WL=${2:-c4}
CMD="python bench.py --steps 20 --warmup 3 --workload $WL --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_$TAG.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-80} -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 5 -c 2 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/plain_$TAG.log
