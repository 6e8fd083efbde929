#!/bin/bash
mkdir -p gpurun_out
line() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); x=d.get('extras',{}).get('strict_deps',{}); print('ms/step %.5f  frac %.3f  strict %.5f (%.3f)  e2e %.3e' % (d['ms_per_step'], d['roofline']['frac'], x.get('ms_per_step',0), x.get('roofline_frac',0), d['e2e']['value']))
"; }
rm -f gpurun_out/r2k_variants.log
for rep in 1 2; do
for lib in libdronesim_b200.so variants/nopf.so; do
  for wl in c4 c4x4; do
    echo -n "$lib $wl " >> gpurun_out/r2k_variants.log
    DSIM_LIB=$PWD/mujoco_drone_b200/$lib timeout 300 python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu-baseline 2>&1 | line >> gpurun_out/r2k_variants.log
  done
done
done
cat gpurun_out/r2k_variants.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py -q -x > gpurun_out/r2k_pytest.log 2>&1; tail -3 gpurun_out/r2k_pytest.log
DSIM_INPUTS_READY=0 DSIM_LIB=$PWD/mujoco_drone_b200/variants/tl.so timeout 300 python tools/timeline_graph.py c4 > gpurun_out/r2k_timeline_c4_strict.log 2>&1
head -22 gpurun_out/r2k_timeline_c4_strict.log
DSIM_INPUTS_READY=1 DSIM_LIB=$PWD/mujoco_drone_b200/variants/tl.so timeout 300 python tools/timeline_graph.py c4 > gpurun_out/r2k_timeline_c4_ready.log 2>&1
head -22 gpurun_out/r2k_timeline_c4_ready.log
