#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2i_pytest.log
tail -30 gpurun_out/r2i_pytest.log | cut -c1-250
python -m pytest tests/test_policy_reference.py -m gpu -q -s 2>&1 | grep "vs the reference"
for pd in fused fused_fp32 fp32; do timeout 400 python bench.py --workload c5 --steps 20 --policy-dtype $pd --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$pd', '%8.2f us/step  %.3e env-steps/s  step kernel %.2f us' % (d['ms_per_step'] * 1e3, d['value'], d['roofline']['env_step_kernel_ms']*1e3))
"; done
for wl in c2 c3; do timeout 300 python bench.py --workload $wl --steps 20 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); x=d.get('extras',{}).get('strict_deps',{}); print(d['config']['workload'][:30], '%7.2f us/step  %.3e  frac %.3f  strict %.2f us (%.3f) e2e %.3e' % (d['ms_per_step'] * 1e3, d['value'], d['roofline']['frac'], x.get('ms_per_step',0)*1e3, x.get('roofline_frac',0), d['e2e']['value']), d['episode_stats'])
"; done
