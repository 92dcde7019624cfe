"""bench.py's reference arm (the CPU oracle port on the host cores) prints ONE JSON line with the keys the driver
reads; under a multi-rank launch only rank 0 prints.  CPU tier: small sample so that it finishes in seconds."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"}


def run_bench(extra_env, *flags):
    env = dict(os.environ, **extra_env)
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
           "--games", "256", "--lock-steps", "4", *flags]
    return subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)


def test_reference_arm_prints_one_contract_line():
    r = run_bench({"RANK": "0", "WORLD_SIZE": "1"})
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "td_updates_per_sec" and d["unit"] == "updates/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    r = run_bench({"RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == ""
