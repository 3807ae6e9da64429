"""bench.py's output contract, checked on the CPU through the reference arm (the GPU arm needs a B200; its line carries the same keys
plus roofline / clocks / gpu_launches and is recorded under profiles/r1_bench_*.json)."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(*args, env=None):
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=600, env={**os.environ, **(env or {})})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line_on_cpu():
    line = _run("--impl", "reference", "--workload", "tiny", "--steps", "2", "--warmup", "1")
    assert BASE_KEYS <= set(line) and line["impl"] == "reference"
    assert line["metric"] == "frames_per_sec" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["steps"] == 2 and line["value"] > 0 and abs(line["ms_per_step"] * line["value"] - 1000.0) < 1.0
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_print_nothing():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "0", "--gpus", "2"],
                         capture_output=True, text=True, timeout=300, env={**os.environ, "RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""
