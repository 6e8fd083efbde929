"""FP32 / FP64 deviation of the floor-contact step from the oracle over several seeds (the tolerances of
tests/test_gpu_ground_contact.py::test_contact_step_matches_oracle are 3x the maxima printed here)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import test_gpu_ground_contact as T
from oracle import oracle as O

for precision in ("fp32", "fp64"):
    for pend in (True, False):
        worst = dict(pos=0.0, vel=0.0, acc=0.0)
        touching = 0
        for seed in range(1, 9):
            rng = np.random.default_rng(seed)
            n = 256
            qpos, qvel, act, actions, params = T._near_floor(rng, n, pend, zmax=1.5 if pend else 0.3)
            env = T._mk(num_drones=n, precision=precision, pendulum=pend)
            T._set(env, qpos, qvel, act, params)
            qpos_d, qvel_d, act_d, _, _ = env.get_state()
            env.step_tensor(torch.as_tensor(actions, device="cuda"))
            qp, qv, ac, sens, ns = env.get_state()
            a_in = actions.astype(np.float32).astype(np.float64) if precision == "fp32" else actions
            prm = env.drone_params
            for i in range(n):
                m = O.compile_model(np.array(list(prm[i].values())), pend, 100, True, ground=True)
                touching += len(O.collide(m, qpos_d[i])) > 0
                oqp, oqv, oact, osens = O.step(m, qpos_d[i], qvel_d[i], act_d[i], 0.1 + 0.9 * a_in[i], 1)
                worst["pos"] = max(worst["pos"], np.abs(qp[i] - oqp).max())
                worst["vel"] = max(worst["vel"], (np.abs(qv[i] - oqv) / (1 + np.abs(oqv))).max())
                worst["acc"] = max(worst["acc"], (np.abs(sens[i] - osens) / (1 + np.abs(osens))).max())
            env.close()
        print(f"{precision} pendulum={pend}: 8 seeds x 256 states, {touching} touching: max deviation " + ", ".join(f"{k} {v:.2e}" for k, v in worst.items()))
