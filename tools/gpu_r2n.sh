#!/bin/bash
mkdir -p gpurun_out
line() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); x=d.get('extras',{}); print('ms/step %.5f  frac %.3f  strict %.5f (%.3f) hot %.5f e2e %.3e' % (d['ms_per_step'], d['roofline']['frac'], x.get('strict_deps',{}).get('ms_per_step',0), x.get('strict_deps',{}).get('roofline_frac',0), x.get('hot_l2',{}).get('ms_per_step',0), d['e2e']['value']))
"; }
rm -f gpurun_out/r2n_variants.log
for lib in libdronesim_b200.so variants/blk64.so variants/deal.so variants/early2.so variants/minb3.so; do
  for wl in c4 c4x4; do
    echo -n "$lib $wl " >> gpurun_out/r2n_variants.log
    DSIM_LIB=$PWD/mujoco_drone_b200/$lib timeout 300 python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu-baseline 2>&1 | line >> gpurun_out/r2n_variants.log
  done
done
cat gpurun_out/r2n_variants.log
