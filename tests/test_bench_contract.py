"""bench.py's reference (CPU) arm prints ONE JSON line with the contract's keys (runs on CPU)."""
import json
import os
import subprocess
import sys

from tests.conftest import ROOT

REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
            "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def run_bench(extra_env=None, *args):
    env = dict(os.environ, **(extra_env or {}))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                        "--steps", "2", "--warmup", "1", "--cpu-rows", "20000", *args],
                       capture_output=True, text=True, env=env, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    return lines


def test_reference_arm_json_line():
    lines = run_bench()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d)
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["rows"] == 10_000_000 and d["config"]["dim"] == 768 and d["config"]["k"] == 10
    assert d["vs_baseline"] is None            # BASELINE.md publishes no number for this metric


def test_reference_arm_other_ranks_print_nothing():
    # under torchrun (N > 1) rank 0 alone runs and prints; the other ranks exit 0 without work
    lines = run_bench({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert lines == []
