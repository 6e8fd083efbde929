#!/bin/bash
# round-2 evidence pass on the final build: GPU tests, driver-style bench line, launch list, full ncu capture of the step kernel in
# the steady state of the workload (1500-step pre-roll: in-kernel resets active), size sweep incl. the 2 M-env regression + its capture
mkdir -p gpurun_out
TAG=${1:-r02a}
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc $?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${TAG}_bench.json") if l.startswith("{")][-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["roofline"]["frac"], d["clocks"])
print({k:(v.get("ms_per_step"), v.get("roofline_frac")) for k,v in d["extras"].items()})
print(d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
CMD="python bench.py --steps 20 --warmup 3 --preroll 1500 --no-cpu-baseline --no-extras --no-graph"
SKIP=$(( (1500 + 3) * 8 + 2 ))
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 20 --warmup 3 --preroll 20 --no-cpu-baseline --no-extras --no-graph > gpurun_out/ncu_launches_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s $SKIP -c 3 -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log | cut -c1-200
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:step_kernel -s $SKIP -c 24 --csv --log-file gpurun_out/dram_${TAG}.csv $CMD > gpurun_out/dram_${TAG}.log 2>&1
for n in 16384 32768 65536 131072 262144 524288 1048576 2097152; do
  timeout 600 python bench.py --workload c4 --envs $n --steps 20 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); x=d.get('extras',{}).get('strict_deps',{}); print('envs %8d  %7.2f us/step  %.3e env-steps/s  frac %.3f  strict %.2f us (%.3f)  replicas %d  e2e %.3e' % (d['config']['envs_per_gpu'], d['ms_per_step'] * 1e3, d['value'], d['roofline']['frac'], x.get('ms_per_step',0)*1e3, x.get('roofline_frac',0), d['config']['replicas'], d['e2e']['value']))
" >> gpurun_out/size_sweep_${TAG}.txt
done
cat gpurun_out/size_sweep_${TAG}.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 40 -c 2 -o gpurun_out/prof_${TAG}_2m -f python bench.py --workload c4 --envs 2097152 --steps 10 --warmup 3 --preroll 10 --no-cpu-baseline --no-extras --no-graph > gpurun_out/ncu_full_${TAG}_2m.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 40 -c 2 -o gpurun_out/prof_${TAG}_512k -f python bench.py --workload c4 --envs 524288 --steps 10 --warmup 3 --preroll 10 --no-cpu-baseline --no-extras --no-graph > gpurun_out/ncu_full_${TAG}_512k.log 2>&1
ls -la gpurun_out/prof_${TAG}*
