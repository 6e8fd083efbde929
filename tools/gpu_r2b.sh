#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
for lib in libdronesim_b200.so variants/nohint.so; do
  for wl in c4 c4x4 c3; do
    echo -n "$lib $wl " >> gpurun_out/r2b_variants.log
    DSIM_LIB=$PWD/mujoco_drone_b200/$lib timeout 300 python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu-baseline --no-extras 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.5f  value %.3e  frac %.3f e2e %.3e' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value']))
" >> gpurun_out/r2b_variants.log
  done
done
cat gpurun_out/r2b_variants.log
DSIM_LIB=$PWD/mujoco_drone_b200/variants/tl.so timeout 300 python tools/timeline_graph.py c4 > gpurun_out/r2b_timeline_c4.log 2>&1
DSIM_LIB=$PWD/mujoco_drone_b200/variants/tl.so timeout 300 python tools/timeline_graph.py c4 524288 > gpurun_out/r2b_timeline_c4x4.log 2>&1
head -40 gpurun_out/r2b_timeline_c4.log
for lib in libdronesim_b200.so variants/nohint.so; do
CMD="python bench.py --steps 20 --warmup 3 --preroll 300 --no-cpu-baseline --no-extras --no-graph"
DSIM_LIB=$PWD/mujoco_drone_b200/$lib timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:step_kernel -s 2404 -c 40 --csv --log-file gpurun_out/r2b_dram_$(basename $lib .so).csv $CMD > gpurun_out/r2b_dram_$(basename $lib .so).log 2>&1
done
