"""bench.py contract checks that need no GPU: the reference arm's JSON line and the loud failure of our arm."""
import json
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def test_reference_arm_prints_one_contract_line():
    """`--impl reference` times the CPU port of the reference call pattern (bounded sample) and prints ONE
    JSON line with the contract keys, impl = reference, a cpu_baseline describing the run and a zero-copy e2e."""
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "c1s",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "retrieval queries/sec" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"] and d["vs_baseline"] is None
    # every host core, whatever OMP_NUM_THREADS the launcher exported; same config object as the product arm
    import os
    assert d["cpu_baseline"]["cores"] == os.cpu_count() == d["cpu_baseline"]["threads"]
    sys.path.insert(0, str(ROOT))
    import bench
    assert d["config"] == bench.config_of("c1s", 1)
    assert "clustered" in d["cpu_baseline"]["sample"] and d["config"]["data_kind"] == "clustered"


def test_reference_arm_uses_all_cores_under_torchrun_environment():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to every rank: the reference arm must not inherit it."""
    import os
    env = {**os.environ, "OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"}
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "c1s", "--gpus", "2",
                        "--steps", "2", "--warmup", "3"], capture_output=True, text=True, timeout=600, env=env)
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    assert d["cpu_baseline"]["threads"] == os.cpu_count() and d["n_gpus"] == 2
    env["RANK"] = "1"
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "c1s", "--gpus", "2",
                        "--steps", "2", "--warmup", "3"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and not r.stdout.strip()          # other ranks exit 0 without work


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_our_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: without a B200 the product arm must fail loudly instead of timing something else."""
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1", "--no-extras",
                        "--no-cpu-baseline"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
