"""bench.py's reference arm on a box without a GPU: the JSON line carries every key of the contract
(metric / unit / value / config.workload / cpu_baseline / e2e with zero copies), non-zero ranks of a
torchrun launch print nothing, and the main arm refuses to run without a device instead of
falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*argv, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], capture_output=True, text=True,
                          timeout=600, cwd=ROOT, env={**os.environ, **(env or {})})


def test_reference_arm_line(built_lib):
    res = run_bench("--impl", "reference", "--workload", "region", "--steps", "2", "--warmup", "1", "--cpu-seconds", "1")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    out = json.loads(lines[0])
    assert out["impl"] == "reference" and out["metric"] == "force_evals_per_s" and out["unit"] == "force_evals/s"
    assert out["steps"] == 2 and out["warmup"] == 1 and out["n_gpus"] == 1
    assert out["higher_is_better"] is True and out["scaling"] == "weak" and out["vs_baseline"] is None
    assert out["dtype"] == "f64" and out["data"] == "synthetic" and out["gpu_launches"] == 0
    assert out["value"] > 0 and out["ms_per_step"] == pytest.approx(1e3 / out["value"])
    assert "workload" in out["config"] and "model" not in out["config"]
    cb = out["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == out["value"] and "beads" in cb["sample"]
    assert "openmm" in cb  # the probe for the real reference backend is recorded
    assert out["e2e"] == dict(value=out["value"], unit=out["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_reference_arm_other_ranks_exit_quietly(built_lib):
    res = run_bench("--impl", "reference", "--workload", "region", "--steps", "1", "--warmup", "0", "--gpus", "2",
                    env=dict(RANK="1", LOCAL_RANK="1", WORLD_SIZE="2"))
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_main_arm_needs_a_gpu(built_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    res = run_bench("--workload", "region", "--steps", "1", "--warmup", "0", "--no-cpu")
    assert res.returncode != 0
    assert not [l for l in res.stdout.splitlines() if l.startswith("{")]  # no number without a device
