#!/bin/bash
line() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); x=d.get('extras',{}); print('envs %8d ms/step %.5f  frac %.3f  strict %.5f (%.3f)' % (d['config']['envs_per_gpu'], d['ms_per_step'], d['roofline']['frac'], x.get('strict_deps',{}).get('ms_per_step',0), x.get('strict_deps',{}).get('roofline_frac',0)))
"; }
for n in 262144 524288 1048576 2097152 4194304; do timeout 400 python bench.py --steps 20 --warmup 3 --workload c4 --envs $n --no-cpu-baseline 2>&1 | line; done
