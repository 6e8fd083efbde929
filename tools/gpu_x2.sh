#!/bin/bash
# two-envs-per-lane step kernel (DSIM_X2=1): parity suite, then C4 / C3 timing against the one-env kernel
line() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); x=d.get('extras',{}); print('%-30s ms/step %.5f  frac %.3f  strict %.5f hot %.5f' % (d['config']['workload'][:30], d['ms_per_step'], d['roofline']['frac'], x.get('strict_deps',{}).get('ms_per_step',0), x.get('hot_l2',{}).get('ms_per_step',0)))
"; }
DSIM_X2=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py tests/test_gpu_rollout_vs_oracle.py -q -x 2>&1 | tail -5
for x in 0 1; do
  for wl in c4 c3; do
    echo -n "DSIM_X2=$x $wl "; DSIM_X2=$x timeout 300 python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu-baseline 2>&1 | line
  done
done
for n in 65536 100000 150000; do
  for x in 0 1; do echo -n "DSIM_X2=$x c4 $n "; DSIM_X2=$x timeout 300 python bench.py --steps 20 --warmup 3 --workload c4 --envs $n --no-cpu-baseline --no-extras 2>&1 | line; done
done
