"""bench.py prints ONE JSON line with the driver's contract keys (GPU arm; small batch so that it runs in seconds)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import has_cuda

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
def test_bench_native_arm_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "40", "--warmup", "3", "--envs", "8192", "--preroll", "50",
                          "--no-cpu-baseline", "--no-extras"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "e2e", "gpu_launches", "roofline", "clocks"):
        assert k in d, k
    assert d["metric"] == "env-steps/sec" and d["n_gpus"] == 1 and d["steps"] == 40 and d["scaling"] == "weak" and d["vs_baseline"] is None
    # --steps is a lower bound of the timed region: the graph of steps is replayed until the event window is >= 30 ms
    ts = d["config"]["timed_steps"]
    assert d["value"] > 0 and ts >= 40 and d["gpu_launches"] == ts and "workload" in d["config"] and "CUDA graph" in d["config"]["launch"]
    assert d["config"]["timed_window_ms"] >= 25.0 and abs(d["ms_per_step"] - d["config"]["timed_window_ms"] / ts) < 1e-9
    assert d["clocks"]["samples"] >= 10
    er = d["e2e"]["roofline"]
    assert er["bound"] == "pcie" and er["peak"] > 0 and 0 < er["frac"] < 1.5
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 8192 * 16 and d["e2e"]["d2h_bytes_per_step"] == 8192 * (4 * 22 + 5)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert "sm_mhz" in d["clocks"] and "reasons" in d["clocks"]
