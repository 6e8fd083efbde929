#!/bin/bash
line() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); x=d.get('extras',{}); print('%-34s ms/step %.5f  frac %.3f  strict %.5f hot %.5f  stats %s' % (d['config']['workload'][:34], d['ms_per_step'], d['roofline']['frac'], x.get('strict_deps',{}).get('ms_per_step',0), x.get('hot_l2',{}).get('ms_per_step',0), {k:round(v,1) for k,v in d['episode_stats'].items()}))
"; }
echo -n "c3 "; timeout 300 python bench.py --workload c3 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | line
echo -n "c3 noreset "; DSIM_BENCH_NORESET=1 timeout 300 python bench.py --workload c3 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | line
echo -n "c4@65536 "; timeout 300 python bench.py --workload c4 --envs 65536 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | line
echo -n "c4@65536 noreset "; DSIM_BENCH_NORESET=1 timeout 300 python bench.py --workload c4 --envs 65536 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | line
