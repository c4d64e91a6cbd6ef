"""The reference arm of bench.py runs on the host alone: its JSON line must carry the keys the driver reads
(metric / config of the engine arm, `impl`, `cpu_baseline`, an `e2e` block that repeats the line's own value)."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), cwd=ROOT, env=env,
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    lines = _run("--impl", "reference", "--workload", "c1", "--steps", "2", "--warmup", "1")
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["metric"] == "grid-point RK-steps/sec" and d["unit"] == "point-steps/s"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["unit"] == d["unit"] and cb["value"] == d["value"] and cb["sample"]
    assert cb["all_cores_independent_grids"]["cores"] == (os.cpu_count() or 1) and cb["all_cores_independent_grids"]["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_under_torchrun_runs_on_rank_zero_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    assert _run("--impl", "reference", "--workload", "c1", "--gpus", "2", "--steps", "1", "--warmup", "0", env=env) == []
