#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s); timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2l_bench_ref.json 2> gpurun_out/r2l_bench_ref.err; echo "reference arm rc $? in $(( $(date +%s) - S )) s"; cut -c1-160 gpurun_out/r2l_bench_ref.json
S=$(date +%s); timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "native arm rc $? in $(( $(date +%s) - S )) s"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2l_bench.json") if l.startswith("{")][-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline"]["traffic_pages"], d["e2e"]["value"], d["e2e"]["roofline"]["frac"])
print({k:(v.get("ms_per_step"), v.get("roofline_frac"), v.get("error")) for k,v in d["extras"].items()})
print(d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["clocks"]["samples"])
PY
