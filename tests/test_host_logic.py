"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol include/dronesim_b200.h declares, the host
mirror keeps the reference's names, the env-range sharding is exact, and the N>1 path (episode-statistics all-reduce,
setpoint broadcast) works over gloo with world_size 2.  No compute call is made: there is no CPU path to call."""
import os
import re
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol():
    import mujoco_drone_b200 as M
    hdr = open(os.path.join(ROOT, "include", "dronesim_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(dsim_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 28
    L = M._lib.load()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(M._lib.EXPORTS) == declared                      # the ctypes binding binds exactly the header
    assert L.dsim_abi_version() == M._lib.ABI_VERSION == int(re.search(r"#define DSIM_ABI_VERSION (\d+)", hdr).group(1))
    assert L.dsim_obs_dim(8, 1) == 22 and L.dsim_obs_dim(0, 0) == 29 and L.dsim_obs_dim(99, 1) < 0


def test_config_struct_matches_header_layout():
    """ctypes DsimConfig field order == the header's struct (ABI guard: struct_size is also checked by dsim_create)"""
    import mujoco_drone_b200 as M
    hdr = open(os.path.join(ROOT, "include", "dronesim_b200.h")).read()
    body = hdr[hdr.index("typedef struct DsimConfig {"):hdr.index("} DsimConfig;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.split("{")[-1].strip()
        if not decl:
            continue
        for part in decl.split(",")[0:]:
            nm = re.sub(r"\[.*?\]", "", part).split()[-1].lstrip("*")
            names.append(nm)
    assert names == [f[0] for f in M._lib.DsimConfig._fields_]


def test_product_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import mujoco_drone_b200 as M
    with pytest.raises(M._lib.DsimError, match="no CPU fallback"):
        M.BaseDroneEnv(dict(M.base_config, num_drones=4))


def test_host_mirror_keeps_reference_names():
    import mujoco_drone_b200 as M
    for name in ("vector_reset", "vector_step", "reset_at", "reset_model", "get_drone_states", "control_reference",
                 "move_mocap_to", "render", "close", "reset", "_get_obs"):
        assert callable(getattr(M.BaseDroneEnv, name))
    assert set(M.base_config) >= {"seed", "frequency", "skip_steps", "reference", "start_pos", "max_distance", "random_start_pos",
                                  "random_params", "pendulum", "state_difficulty", "param_difficulty", "max_random_offset",
                                  "rp_variance", "vel_variance", "ang_vel_variance", "mass_interval", "arm_len_interval",
                                  "motor_force_interval", "motor_tau_interval", "pendulum_length_interval", "weight_mass_interval",
                                  "pendulum_rp_variance", "pendulum_ang_vel_variance", "reward_fcn", "terminated_fcn", "max_steps",
                                  "regen_env_at_steps", "train_vis", "window_title", "controlled", "mocaps"}
    assert len(M.observation_wrappers.WRAPPERS) == 15 and len(M.rewards.REWARD_NAMES) == 17
    with pytest.raises(NotImplementedError):
        M.rewards.resolve_reward(lambda env, s, a, n: 0.0)          # arbitrary Python rewards cannot run in the kernel


def test_compat_import_shims_resolve_reference_module_paths():
    compat = os.path.join(ROOT, "mujoco_drone_b200", "compat")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "environments" or k.startswith("environments.")}
    sys.path.insert(0, compat)
    try:
        from environments.BaseDroneEnv import BaseDroneEnv, base_config, default_termination_fcn   # noqa: F401
        from environments.observation_wrappers import LocalFrameRPYParamsEnv, LocalFramePRYaccEnv   # noqa: F401
        from environments.rewards import distance_energy_reward, reward_3                           # noqa: F401
        import mujoco_drone_b200 as M
        assert BaseDroneEnv is M.BaseDroneEnv and LocalFrameRPYParamsEnv.OBS_ID == 8
        assert M.rewards.resolve_reward(distance_energy_reward) == 2 and M.rewards.resolve_reward(reward_3) == 16
    finally:
        sys.path.remove(compat)
        for k in [k for k in sys.modules if k == "environments" or k.startswith("environments.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_shard_ranges_partition_exactly():
    from mujoco_drone_b200 import dist
    for total in (1, 7, 64, 4096, 1048576, 1048577):
        for world in (1, 2, 3, 8):
            r = [dist.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    cfg = dist.shard_config({"num_drones": 1048576, "env_id_offset": 5}, 3, 8, device=3)
    assert cfg["num_drones"] == 131072 and cfg["env_id_offset"] == 5 + 3 * 131072 and cfg["device"] == 3
    with pytest.raises(ValueError):
        dist.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, out):
    import torch.distributed as tdist
    from mujoco_drone_b200 import dist
    tdist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        lo, hi = dist.shard_range(1000, rank, world)
        # each rank reports the statistics of its own env range (synthetic: return = global env id)
        st = {"sum_return": float(sum(range(lo, hi))), "sum_length": float(10 * (hi - lo)), "n_episodes": float(hi - lo),
              "n_nonfinite": float(rank), "n_near_ground": 0.0}
        red = dist.allreduce_episode_stats(st)
        ref = dist.broadcast_reference([1.0 + rank, 2.0, 15.0, 0.5], src=0)
        out.put((rank, red, ref.tolist()))
    finally:
        tdist.destroy_process_group()


def test_two_rank_gloo_statistics_allreduce_and_reference_broadcast():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, red, ref in res:
        assert red["sum_return"] == sum(range(1000)) and red["n_episodes"] == 1000 and red["n_nonfinite"] == 1.0
        assert red["mean_return"] == pytest.approx(499.5) and red["mean_length"] == pytest.approx(10.0)
        assert ref == [1.0, 2.0, 15.0, 0.5]                          # rank 0's setpoint everywhere


def test_single_process_allreduce_is_identity():
    from mujoco_drone_b200 import dist
    st = {"sum_return": 6.0, "sum_length": 30.0, "n_episodes": 3.0, "n_nonfinite": 0.0, "n_near_ground": 2.0}
    out = dist.allreduce_episode_stats(st)
    assert out["mean_return"] == 2.0 and out["mean_length"] == 10.0 and out["n_near_ground"] == 2.0


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the oracle port timed on the host cores) prints ONE JSON line with the contract's keys"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env-steps/sec" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["value"] > 0 and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
