#!/bin/bash
# floor-contact cost: (a) C4 in free flight with the slow path compiled in (generic instantiation) against the specialised kernel,
# (b) a batch in which every drone lies on the floor
line() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-30s ms/step %.5f  frac %.3f' % (d['config']['workload'][:30], d['ms_per_step'], d['roofline']['frac']))
"; }
echo -n "C4 specialised            "; timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | line
echo -n "C4 ground_contact=True    "; DSIM_BENCH_GROUND=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | line
echo -n "C4 generic (frame_skip via timeline env) "; DSIM_TIMELINE=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | line
PYTHONPATH=$PWD python - <<'PY'
import time, numpy as np, torch
import mujoco_drone_b200 as M
for n in (32768, 131072):
    cfg = dict(M.base_config)
    cfg.update(dict(num_drones=n, start_pos=[0, 0, 1.6, 0], reference=[0, 0, 1.6, 0], max_distance=1000, max_steps=10**6, random_params=True,
                    ground_contact=True, angle_variance=[0.5, 0.5], vel_variance=[0.5, 0.5, 0.5], ang_vel_variance=[1, 1, 1],
                    pendulum_rp_variance=[0.3, 0.3], max_random_offset=0.3, seed=9))
    env = M.observation_wrappers.LocalFrameRPYParamsEnv(cfg)
    env.reset_tensor()
    off = torch.full((n, 4), -1.0, device="cuda")
    for phase, T in (("falling / first impacts", 100), ("bouncing", 300), ("at rest on the floor", 800), ("at rest on the floor", 200)):
        torch.cuda.synchronize(); t0 = time.time()
        for _ in range(T):
            env.step_tensor(off)
        torch.cuda.synchronize(); dt = time.time() - t0
        print(f"{n} envs, {phase:26s}: {dt / T * 1e6:9.1f} us/step = {n * T / dt:.3e} env-steps/s")
    env.close()
PY
