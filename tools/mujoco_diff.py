#!/usr/bin/env python
"""Real-MuJoCo diff harness for the env-step physics.  AUTO-ACTIVATING: it does its work the first time
`import mujoco, dm_control` succeeds AND the reference tree is reachable; otherwise it reports why it cannot run.

    python tools/mujoco_diff.py [--reference /root/reference] [--write-golden]

What it does when it can run
  1. imports the reference's own model builder, environments/env_gen.py (make_sim / mjcf_to_mjmodel, :76-133), UNMODIFIED,
     builds ONE MjModel holding all test drones (per-drone parameters like BaseDroneEnv.generate_drone_params would hand
     it), i.e. exactly the model BaseDroneEnv.__init__ (:125-133) steps;
  2. compares MuJoCo's compiled per-body constants (body_mass, body_inertia, body_ipos, body_iquat, site positions, actuator
     gear / dynprm, dof_damping, opt.timestep) with oracle.compile_model (env_gen + the "%.5g" XML stage restated);
  3. writes the seeded states / activations / controls of tests/test_gpu_parity.py::test_substep_matches_oracle into MjData
     (drone-major qpos[9i:9i+9], qvel[8i:8i+8], act[4i:4i+4]: BaseDroneEnv.py:367-375), calls mujoco.mj_forward and then
     mujoco.mj_step(model, data, nstep=frame_skip) (mujoco_vecenv.py:404-407) and compares qacc / sensordata (forward) and
     qpos / qvel / act / sensordata (after 1, 2, 3 and 100 steps) with oracle.forward / oracle.step: tolerance 1e-10;
  4. repeats 3 (1, 2, 3 steps) on near-floor states - the inputs of tests/test_gpu_ground_contact.py - against oracle models with
     ground=True: contact count, qacc / sensordata, stepped state; tolerance 1e-6 (MuJoCo's Newton solver stops at 1e-8);
  5. with --write-golden freezes MuJoCo's outputs into tests/golden/mjstep.npz, after which tests/test_mujoco_pin.py pins the
     oracle against them on every box (no mujoco needed any more).

Probe outcomes so far are recorded in DESIGN.md §3 (build container and the gpurun B200 box: ModuleNotFoundError for both
`mujoco` and `dm_control`; no wheel in /opt/wheelhouse; no network).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NOMINAL = np.array([1, 0.17, 7, 0.01, 1.2, 0.3])
PKEYS = ("mass", "arm_len", "motor_force", "motor_tau", "pendulum_len", "weight_mass")


def probe(reference_root="/root/reference"):
    """-> (ok, reason)"""
    try:
        import mujoco  # noqa: F401
    except Exception as ex:
        return False, f"import mujoco failed: {type(ex).__name__}: {ex}"
    try:
        import dm_control  # noqa: F401
        from dm_control import mjcf  # noqa: F401
    except Exception as ex:
        return False, f"import dm_control failed: {type(ex).__name__}: {ex}"
    if not os.path.exists(os.path.join(reference_root, "environments", "env_gen.py")):
        return False, f"{reference_root}/environments/env_gen.py not present (the reference tree does not travel to the GPU box)"
    return True, "mujoco + dm_control importable, reference tree present"


def seeded_cases(n=192, seed=7, pend=True):
    """the inputs of tests/test_gpu_parity.py::test_substep_matches_oracle (same generator, same seed)"""
    rng = np.random.default_rng(seed)
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    pos = np.array([0, 0, 15.0]) + rng.normal(size=(n, 3))
    qpos = np.concatenate([pos, q] + ([rng.normal(size=(n, 2)) * 0.6] if pend else []), axis=1)
    qvel = rng.normal(size=(n, 8 if pend else 6))
    act = rng.uniform(0, 1, size=(n, 4))
    actions = rng.uniform(0, 1, size=(n, 4))
    params = NOMINAL * rng.uniform(0.85, 1.15, size=(n, 6))
    if not pend:
        params[:, 4:] = 0
    return qpos, qvel, act, actions, params


def near_floor_cases(n=256, seed=11, pend=True):
    """the inputs of tests/test_gpu_ground_contact.py::test_contact_step_matches_oracle (same generator, same seed): random
    attitudes at heights where the drone's geoms reach the floor plane"""
    rng = np.random.default_rng(seed)
    zmax = 1.5 if pend else 0.3
    q = rng.normal(size=(n, 4))
    q[: n // 4] = [1, 0, 0, 0] + 0.05 * rng.normal(size=(n // 4, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    pos = np.stack([rng.normal(size=n), rng.normal(size=n), rng.uniform(0.0, zmax, size=n)], axis=1)
    qpos = np.concatenate([pos, q] + ([rng.normal(size=(n, 2)) * 0.5] if pend else []), axis=1)
    qvel = rng.normal(size=(n, 8 if pend else 6)) * 0.5
    act = rng.uniform(0, 1, size=(n, 4))
    actions = rng.uniform(0, 1, size=(n, 4))
    params = NOMINAL * rng.uniform(0.85, 1.15, size=(n, 6))
    if not pend:
        params[:, 4:] = 0
    return qpos, qvel, act, actions, params


def build_mujoco_model(reference_root, params, frequency):
    """the reference's own env_gen, imported unmodified"""
    import importlib
    import numpy
    if not hasattr(numpy, "long"):
        numpy.long = int                                  # env_gen.py:118 uses np.long (removed in numpy 2)
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    env_gen = importlib.import_module("environments.env_gen")
    dp = [dict(zip(PKEYS, [float(x) for x in p])) for p in params]
    return env_gen.mjcf_to_mjmodel(env_gen.make_sim(dp, frequency=frequency, mocaps=1))


def run(reference_root="/root/reference", write_golden=False, n=192, frequency=100):
    import mujoco
    from oracle import oracle as O
    report = {"mujoco_version": mujoco.__version__, "cases": n, "checks": {}}
    gold = {}
    for pend in (True, False):
        qpos, qvel, act, actions, params = seeded_cases(n, 7, pend)
        nq, nv = (9, 8) if pend else (7, 6)
        model = build_mujoco_model(reference_root, params, frequency)
        assert model.nq == nq * n and model.nv == nv * n and model.na == 4 * n, (model.nq, model.nv, model.na)
        tag = "pend" if pend else "nopend"
        # ---- 2. compiled constants
        worst_c = 0.0
        bodies_per_drone = (model.nbody - 1 - 1) // n     # world + mocap
        for i in range(n):
            m = O.compile_model(params[i], pend, frequency, True)
            b0 = 2 + i * bodies_per_drone                 # world, mocap0, then per drone: A (attachment frame), B, [C, D]
            for k in range(m.nbody - 1):
                mb, ob = b0 + k, 1 + k
                worst_c = max(worst_c, abs(model.body_mass[mb] - m.mass[ob]),
                              np.abs(np.sort(model.body_inertia[mb]) - np.sort(list(m.inertia[ob]))).max(),
                              np.abs(model.body_ipos[mb] - np.array(list(m.ipos[ob]))).max())
            a0 = 4 * i
            for k in range(4):
                worst_c = max(worst_c, np.abs(model.actuator_gear[a0 + k] - np.array(list(m.gear[k]))).max(),
                              abs(model.actuator_dynprm[a0 + k][0] - m.tau[k]))
        worst_c = max(worst_c, abs(model.opt.timestep - O.compile_model(params[0], pend, frequency, True).timestep))
        report["checks"][f"compiled_constants_{tag}"] = worst_c
        # ---- 3. forward + steps
        data = mujoco.MjData(model)
        ctrl = 0.1 + 0.9 * actions                        # BaseDroneEnv.py:269

        def load():
            mujoco.mj_resetData(model, data)
            data.qpos[:] = qpos.ravel(); data.qvel[:] = qvel.ravel(); data.act[:] = act.ravel(); data.ctrl[:] = ctrl.ravel()
        load()
        mujoco.mj_forward(model, data)
        mj_qacc, mj_sens = data.qacc.reshape(n, nv).copy(), data.sensordata.reshape(n, 3).copy()
        worst_f = 0.0
        for i in range(n):
            m = O.compile_model(params[i], pend, frequency, True)
            f = O.forward(m, qpos[i], qvel[i], act[i], ctrl[i])
            worst_f = max(worst_f, (np.abs(f["qacc"] - mj_qacc[i]) / (1 + np.abs(mj_qacc[i]))).max(),
                          (np.abs(f["sensordata"] - mj_sens[i]) / (1 + np.abs(mj_sens[i]))).max())
        report["checks"][f"forward_qacc_sensordata_{tag}"] = worst_f
        gold[f"{tag}_qpos"], gold[f"{tag}_qvel"], gold[f"{tag}_act"], gold[f"{tag}_ctrl"], gold[f"{tag}_params"] = qpos, qvel, act, ctrl, params
        gold[f"{tag}_fwd_qacc"], gold[f"{tag}_fwd_sens"] = mj_qacc, mj_sens
        for nstep in (1, 2, 3, 100):
            load()
            mujoco.mj_step(model, data, nstep=nstep)      # mujoco_vecenv.py:407
            out = (data.qpos.reshape(n, nq).copy(), data.qvel.reshape(n, nv).copy(), data.act.reshape(n, 4).copy(), data.sensordata.reshape(n, 3).copy())
            worst_s = 0.0
            for i in range(n):
                m = O.compile_model(params[i], pend, frequency, True)
                o = O.step(m, qpos[i], qvel[i], act[i], ctrl[i], nstep)
                for a, b in zip(o, (out[0][i], out[1][i], out[2][i], out[3][i])):
                    worst_s = max(worst_s, (np.abs(a - b) / (1 + np.abs(b))).max())
            report["checks"][f"mj_step_x{nstep}_{tag}"] = worst_s
            for name, arr in zip(("qpos", "qvel", "act", "sens"), out):
                gold[f"{tag}_step{nstep}_{name}"] = arr
    # ---- 4. floor contact: the same comparison on states whose geoms reach the floor, oracle models with
    # ground=True.  MuJoCo's Newton solver stops at tolerance 1e-8, so these legs are held to 1e-6, not 1e-10.
    for pend in (True, False):
        qpos, qvel, act, actions, params = near_floor_cases(256, 11, pend)
        n2 = len(qpos)
        nq, nv = (9, 8) if pend else (7, 6)
        model = build_mujoco_model(reference_root, params, frequency)
        data = mujoco.MjData(model)
        ctrl = 0.1 + 0.9 * actions
        tag = "ground_pend" if pend else "ground_nopend"

        def load2():
            mujoco.mj_resetData(model, data)
            data.qpos[:] = qpos.ravel(); data.qvel[:] = qvel.ravel(); data.act[:] = act.ravel(); data.ctrl[:] = ctrl.ravel()
        load2()
        mujoco.mj_forward(model, data)
        report[f"{tag}_mujoco_ncon"] = int(data.ncon)
        mj_qacc, mj_sens = data.qacc.reshape(n2, nv).copy(), data.sensordata.reshape(n2, 3).copy()
        worst_f, ncon_oracle = 0.0, 0
        for i in range(n2):
            m = O.compile_model(params[i], pend, frequency, True, ground=True)
            f = O.forward_contact(m, qpos[i], qvel[i], act[i], ctrl[i])
            ncon_oracle += f["ncon"]
            worst_f = max(worst_f, (np.abs(f["qacc"] - mj_qacc[i]) / (1 + np.abs(mj_qacc[i]))).max(),
                          (np.abs(f["sensordata"] - mj_sens[i]) / (1 + np.abs(mj_sens[i]))).max())
        report[f"{tag}_oracle_ncon"] = ncon_oracle
        report["checks"][f"contact_count_{tag}"] = float(abs(ncon_oracle - int(data.ncon)))
        report["checks"][f"contact_forward_qacc_sensordata_{tag}"] = worst_f
        gold[f"{tag}_qpos"], gold[f"{tag}_qvel"], gold[f"{tag}_act"], gold[f"{tag}_ctrl"], gold[f"{tag}_params"] = qpos, qvel, act, ctrl, params
        gold[f"{tag}_fwd_qacc"], gold[f"{tag}_fwd_sens"] = mj_qacc, mj_sens
        for nstep in (1, 2, 3):
            load2()
            mujoco.mj_step(model, data, nstep=nstep)
            out = (data.qpos.reshape(n2, nq).copy(), data.qvel.reshape(n2, nv).copy(), data.act.reshape(n2, 4).copy(), data.sensordata.reshape(n2, 3).copy())
            worst_s = 0.0
            for i in range(n2):
                m = O.compile_model(params[i], pend, frequency, True, ground=True)
                o = O.step(m, qpos[i], qvel[i], act[i], ctrl[i], nstep)
                for a, b in zip(o, (out[0][i], out[1][i], out[2][i], out[3][i])):
                    worst_s = max(worst_s, (np.abs(a - b) / (1 + np.abs(b))).max())
            report["checks"][f"contact_mj_step_x{nstep}_{tag}"] = worst_s
            for name, arr in zip(("qpos", "qvel", "act", "sens"), out):
                gold[f"{tag}_step{nstep}_{name}"] = arr
    report["pass_1e-10"] = {k: bool(v <= (1e-6 if ("x100" in k or "contact_" in k) else 1e-10)) if "contact_count" not in k else bool(v == 0)
                            for k, v in report["checks"].items()}
    if write_golden:
        path = os.path.join(ROOT, "tests", "golden", "mjstep.npz")
        np.savez_compressed(path, mujoco_version=np.array(mujoco.__version__), frequency=np.array(float(frequency)), **gold)
        report["golden"] = path
    return report


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("DSIM_REFERENCE_ROOT", "/root/reference"))
    ap.add_argument("--write-golden", action="store_true")
    a = ap.parse_args()
    ok, why = probe(a.reference)
    if not ok:
        print(json.dumps({"mujoco_pin": "unavailable", "reason": why}))
        sys.exit(0)
    print(json.dumps(run(a.reference, a.write_golden), indent=1))
