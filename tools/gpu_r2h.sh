#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2h_pytest.log
tail -4 gpurun_out/r2h_pytest.log | cut -c1-200
timeout 600 python tools/fp32_error_hist.py > gpurun_out/r2h_fp32_hist.txt 2> gpurun_out/r2h_fp32_hist.err; tail -3 gpurun_out/r2h_fp32_hist.err; cat gpurun_out/r2h_fp32_hist.txt
rm -f gpurun_out/r2h_sweep.txt
for n in 524288 1048576 2097152 4194304; do
  timeout 600 python bench.py --workload c4 --envs $n --steps 20 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); x=d.get('extras',{}).get('strict_deps',{}); print('envs %8d  %7.2f us/step  %.3e env-steps/s  frac %.3f  strict %.2f us (%.3f)  replicas %d  e2e %.3e' % (d['config']['envs_per_gpu'], d['ms_per_step'] * 1e3, d['value'], d['roofline']['frac'], x.get('ms_per_step',0)*1e3, x.get('roofline_frac',0), d['config']['replicas'], d['e2e']['value']))
" >> gpurun_out/r2h_sweep.txt
done
cat gpurun_out/r2h_sweep.txt
for wl in c2 c3; do timeout 300 python bench.py --workload $wl --steps 20 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); x=d.get('extras',{}).get('strict_deps',{}); print(d['config']['workload'][:30], '%7.2f us/step  %.3e  frac %.3f  strict %.2f us (%.3f) e2e %.3e' % (d['ms_per_step'] * 1e3, d['value'], d['roofline']['frac'], x.get('ms_per_step',0)*1e3, x.get('roofline_frac',0), d['e2e']['value']))
"; done
