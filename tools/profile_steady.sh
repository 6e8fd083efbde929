#!/bin/bash
# Full ncu capture of the step kernel in the STEADY STATE of the random-action workload (after the bench's pre-roll, so the
# in-kernel reset path is exercised).  usage: tools/profile_steady.sh <tag> [workload]
set -e
TAG=${1:-r01s}; WL=${2:-c4}
CMD="python bench.py --steps 20 --warmup 3 --preroll 300 --workload $WL --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_$TAG.log 2>&1
R=$(python -c "import json;print(json.loads([l for l in open('gpurun_out/plain_$TAG.log') if l.startswith('{')][0])['config']['replicas'])")
SKIP=$(( (300 + 3) * R + 4 ))
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s $SKIP -c 3 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log | cut -c1-300
