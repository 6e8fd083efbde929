#!/bin/bash
line() { python -c "
import sys, json
ok=False
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); ok=True; print('%-40s ms/step %.5f  frac %.3f  launch=%s' % (d['config']['workload'][:40], d['ms_per_step'], d['roofline']['frac'], d['config'].get('launch', d['config'].get('graph'))[:50]))
if not ok: print('NO JSON LINE')
"; }
for flags in "--workload c1" "--workload c4 --no-graph --steps 30" "--workload c4 --strict-deps" "--workload c4 --no-flush" "--workload c4x4" "--workload c2 --steps 7" "--workload c3 --steps 1000" "--workload c5 --steps 3 --fuse-sampling" "--workload c4 --envs 100000"; do
  echo -n "[$flags] "; timeout 400 python bench.py $flags --warmup 3 --no-cpu-baseline --no-extras 2>/tmp/err.txt | line; tail -2 /tmp/err.txt | cut -c1-200
done
