#!/bin/bash
line() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); x=d.get('extras',{}); print('%-30s ms/step %.5f  frac %.3f  strict %.5f hot %.5f' % (d['config']['workload'][:30], d['ms_per_step'], d['roofline']['frac'], x.get('strict_deps',{}).get('ms_per_step',0), x.get('hot_l2',{}).get('ms_per_step',0)))
"; }
for rep in 1 2; do
for lib in libdronesim_b200.so variants/stat1.so; do
  for wl in c4 c3 c4x4; do
    echo -n "$lib $wl "; DSIM_LIB=$PWD/mujoco_drone_b200/$lib timeout 300 python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu-baseline 2>&1 | line
  done
done
done
python -m pytest tests -m gpu -q 2>&1 | tail -2
