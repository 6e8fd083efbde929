#!/bin/bash
# kernel-variant sweep on the GPU box: one short bench line per prebuilt library variant
for lib in mujoco_drone_b200/libdronesim_b200.so mujoco_drone_b200/variants/*.so; do
  for wl in c4 c4x4; do
    echo -n "$lib $wl "
    DSIM_LIB=$PWD/$lib python bench.py --steps 400 --warmup 10 $BENCH_EXTRA --workload $wl --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); x=d.get('extras',{}); print('ms/step %.4f  value %.3e  frac %.3f  hot %.4f  flushed %.4f' % (d['ms_per_step'], d['value'], d['roofline']['frac'], x.get('hot_l2',{}).get('ms_per_step',0), x.get('flushed_per_step',{}).get('ms_per_step',0)))
"
  done
done
