"""GPU tier: bench.py's own arm prints ONE JSON line with the contract keys (headline + a `configs` entry), on a small
shape so that it takes seconds.  The values are not asserted beyond being positive and consistent."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

HEAD_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "roofline", "cpu_baseline", "clocks", "gpu_launches", "configs"}


def test_bench_line_contract():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--games", "1024",
           "--lock-steps", "64", "--no-extras", "--only-config", "0", "--config-steps", "2", "--cpu-seconds", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=600,
                       env=dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert HEAD_KEYS <= set(d), HEAD_KEYS - set(d)
    assert d["metric"] == "td_updates_per_sec" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["gpu_launches"] == 2                                   # one persistent launch per bench step
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 4 * 1114112 and e["d2h_bytes_per_step"] > e["h2d_bytes_per_step"]
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12 and rf["unit"] == "GB/s"
    assert 0 < rf["l2"]["frac"] < 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and "thread_scaling" in cb
    assert "workload" in d["config"] and "model" not in d["config"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    c0 = d["configs"]["configs[0]"]
    assert c0["metric"] == "greedy_moves_per_sec" and c0["value"] > 0 and c0["scaling"] == "strong"
    assert c0["oracle_match"] is True                               # the CPU-baseline leg doubles as a checker
    assert c0["e2e"]["value"] > 0 and c0["roofline"]["l2_gather"]["achieved_gathers_per_sec"] > 0
    assert c0["cpu_baseline"]["kind"] == "port" and c0["config"]["longest_game_moves"] > 0
