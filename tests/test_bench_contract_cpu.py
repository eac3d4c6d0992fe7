"""CPU: the bench lines committed under profiles/ (written by bench.py on the GPU box) carry every key of the driver's
contract; bench.py's argument handling and its reference arm's early exit for non-zero ranks work without a GPU."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "clocks", "e2e", "gpu_launches"]


def last_line(name):
    path = os.path.join(ROOT, "profiles", name)
    return json.loads(open(path).read().strip().splitlines()[-1])


@pytest.mark.parametrize("name,n", [("r01_bench_n1.json", 1), ("r01_bench_n2.json", 2), ("r01_bench_n4.json", 4), ("r01_bench_n8.json", 8)])
def test_committed_bench_lines_follow_the_contract(name, n):
    d = last_line(name)
    for k in BASE_KEYS:
        assert k in d, k
    assert d["n_gpus"] == n and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "bf16" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert abs(d["value"] - n * 1e3 / d["ms_per_step"]) <= 1e-6 * d["value"]            # whole-job steps/s
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
    c = d["clocks"]
    assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if n == 1:
        r = d["roofline"]
        for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
            assert k in r, k
        assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
        b = d["cpu_baseline"]
        for k in ("value", "unit", "cores", "kind", "sample"):
            assert k in b, k
        assert b["kind"] in ("reference", "port") and b["cores"] >= 1
    else:
        assert d["replicas_consistent"] is True


def test_committed_reference_arm_line():
    d = last_line("r01_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["metric"] == last_line("r01_bench_n1.json")["metric"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("reference", "port")


def test_reference_arm_nonzero_rank_exits_quietly():
    """Under torchrun only rank 0 times the CPU reference; the other ranks exit 0 without output or work."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "3"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_multi_gpu_request_without_torchrun_is_refused():
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "2"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 2 and "torch.distributed.run" in r.stderr
