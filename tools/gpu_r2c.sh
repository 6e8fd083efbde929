#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2c_pytest.log
tail -25 gpurun_out/r2c_pytest.log | cut -c1-250
line() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); x=d.get('extras',{}).get('strict_deps',{}); print('ms/step %.5f  frac %.3f  strict %.5f  e2e %.3e' % (d['ms_per_step'], d['roofline']['frac'], x.get('ms_per_step',0), d['e2e']['value']))
"; }
for lib in libdronesim_b200.so variants/dyn2.so variants/blk64.so variants/dyn2blk64.so; do
  for wl in c4 c4x4; do
    echo -n "$lib $wl " >> gpurun_out/r2c_variants.log
    DSIM_LIB=$PWD/mujoco_drone_b200/$lib timeout 300 python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu-baseline 2>&1 | line >> gpurun_out/r2c_variants.log
  done
done
for mb in 32 64 96; do
  for wl in c4 c4x4; do
    echo -n "persist${mb}MB $wl " >> gpurun_out/r2c_variants.log
    DSIM_VERBOSE=1 DSIM_L2_PERSIST_MB=$mb timeout 300 python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu-baseline --no-extras 2>gpurun_out/r2c_persist_$mb.err | line >> gpurun_out/r2c_variants.log
  done
done
cat gpurun_out/r2c_variants.log
grep -h "persisting" gpurun_out/r2c_persist_*.err | sort | uniq -c
DSIM_LIB=$PWD/mujoco_drone_b200/variants/tl.so timeout 300 python tools/timeline_graph.py c4 > gpurun_out/r2c_timeline_c4.log 2>&1
head -32 gpurun_out/r2c_timeline_c4.log
