#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py -q -x > gpurun_out/r2m_pytest.log 2>&1; tail -3 gpurun_out/r2m_pytest.log
line() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); x=d.get('extras',{}); print('ms/step %.5f  frac %.3f  strict %.5f (%.3f) hot %.5f e2e %.3e' % (d['ms_per_step'], d['roofline']['frac'], x.get('strict_deps',{}).get('ms_per_step',0), x.get('strict_deps',{}).get('roofline_frac',0), x.get('hot_l2',{}).get('ms_per_step',0), d['e2e']['value']))
"; }
rm -f gpurun_out/r2m_variants.log
for rep in 1 2; do
for lib in libdronesim_b200.so variants/nodefer.so; do
  for wl in c4 c4x4; do
    echo -n "$lib $wl " >> gpurun_out/r2m_variants.log
    DSIM_LIB=$PWD/mujoco_drone_b200/$lib timeout 300 python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu-baseline 2>&1 | line >> gpurun_out/r2m_variants.log
  done
done
done
cat gpurun_out/r2m_variants.log
for n in 16384 65536 262144; do echo -n "envs $n " ; timeout 300 python bench.py --steps 20 --warmup 3 --workload c4 --envs $n --no-cpu-baseline 2>&1 | line; done
DSIM_INPUTS_READY=1 DSIM_LIB=$PWD/mujoco_drone_b200/variants/tl.so timeout 300 python tools/timeline_graph.py c4 > gpurun_out/r2m_timeline_c4_ready.log 2>&1
head -22 gpurun_out/r2m_timeline_c4_ready.log
