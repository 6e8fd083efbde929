#!/bin/bash
# env-steps/s of the C4 env step against the batch size (one GPU)
for n in 16384 32768 65536 131072 262144 524288 1048576 2097152; do
  python bench.py --workload ${1:-c4} --envs $n --no-cpu-baseline --no-extras 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('envs %8d  %7.2f us/step  %.3e env-steps/s  frac %.3f  replicas %d  e2e %.3e' % (d['config']['envs_per_gpu'], d['ms_per_step'] * 1e3, d['value'], d['roofline']['frac'], d['config']['replicas'], d['e2e']['value']))
"
done
