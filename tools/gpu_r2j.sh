#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2j_pytest.log
tail -4 gpurun_out/r2j_pytest.log | cut -c1-250
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for pd in fused_fp32; do timeout 400 python bench.py --workload c5 --steps 20 --policy-dtype $pd --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$pd', '%8.2f us/step  %.3e env-steps/s  step kernel %.2f us' % (d['ms_per_step'] * 1e3, d['value'], d['roofline']['env_step_kernel_ms']*1e3))
"; done
/usr/bin/time -v timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc $?"; grep -E "Elapsed|Maximum resident" gpurun_out/r2j_bench.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2j_bench.json") if l.startswith("{")][-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["roofline"]["frac"])
print({k:(v.get("ms_per_step"), v.get("roofline_frac"), v.get("error")) for k,v in d["extras"].items()})
print(d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
/usr/bin/time -v timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2j_bench_ref.json 2> gpurun_out/r2j_bench_ref.err; grep -E "Elapsed" gpurun_out/r2j_bench_ref.err; cut -c1-200 gpurun_out/r2j_bench_ref.json
