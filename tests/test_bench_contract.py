"""CPU: the bench contract that can be checked without a GPU — the reference arm's JSON line and the workload table."""

from __future__ import annotations

import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def test_reference_arm_prints_one_contract_line() -> None:
    proc = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                          capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [line for line in proc.stdout.splitlines() if line.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert BASE_KEYS <= set(line) and line["impl"] == "reference"
    assert line["metric"].startswith("agent-steps/sec") and line["unit"] == "agent-steps/s" and line["higher_is_better"] is True
    assert line["vs_baseline"] is None and line["value"] > 0 and line["steps"] == 2
    assert line["config"]["workload"].startswith("configs[1]")
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_leaves_other_ranks_idle() -> None:
    """Under torchrun only rank 0 runs the CPU arm; the other ranks exit 0 without a line."""
    import os

    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    proc = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                          capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert proc.returncode == 0 and not [line for line in proc.stdout.splitlines() if line.startswith("{")]


def test_workload_table_matches_baseline_json() -> None:
    sys.path.insert(0, str(ROOT))
    import bench

    baseline = json.loads((ROOT / "BASELINE.json").read_text())
    assert "agent-steps" in baseline["metric"]
    c2 = bench.WORKLOADS["c2"]
    assert (c2["towns"], c2["grid"], c2["agents"], c2["variant"]) == (1024, 48, 8, "hybrid")
    assert bench.WORKLOADS["c3-full"]["bytes_per_agent_step"] == 3680 and bench.WORKLOADS["c3-compact"]["bytes_per_agent_step"] == 1756
    assert c2["bytes_per_agent_step"] == 4 * 4 * 11 * 11 + 4 * 81 + 36 + 416
