"""Embedding API (mirror of the reference's bridge.SimulationEngine, src/multimm/bridge.py:16-164) on a
box without a GPU: schema / validation, and the retry-and-log plumbing around a run that can only
fail here — loudly, naming the device, never on a CPU platform."""
import os

import pytest

from multimm_b200._lib import Error
from multimm_b200.bridge import ATTEMPTS, SimulationEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BEDPE = os.path.join(ROOT, "tests", "golden", "synthetic_loops.bedpe")


def test_schema_and_validation():
    schema = SimulationEngine.get_schema()
    assert {"N_BEADS", "LOOPS_PATH", "PLATFORM", "EV_POWER"} <= set(schema["properties"])
    out = SimulationEngine.validate_params(dict(LOOPS_PATH=BEDPE, OUT_PATH="/tmp/x", N_BEADS="1234", SHUFFLE_CHROMS="yes"))
    assert out["N_BEADS"] == 1234 and out["SHUFFLE_CHROMS"] is True
    with pytest.raises(Exception):
        SimulationEngine.validate_params(dict(LOOPS_PATH="", OUT_PATH="/tmp/x"))


def test_cpu_fallback_is_refused(tmp_path):
    with pytest.raises(Error, match="no CPU platform"):
        SimulationEngine.run_in_process(dict(LOOPS_PATH=BEDPE, OUT_PATH=str(tmp_path / "o")), fallback_to_cpu=True)


def test_device_failure_is_retried_logged_and_raised(built_lib, tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    params = dict(PLATFORM="B200", N_BEADS=20000, LOOPS_PATH=BEDPE, OUT_PATH=str(tmp_path / "o"), SAVE_PLOTS=False)
    with pytest.raises(Error, match="CUDA error"):
        SimulationEngine.run_in_process(params)
    log = (tmp_path / "o" / "metadata" / "output.log").read_text()
    assert log.count("Simulation failed (platform error)") == ATTEMPTS
    assert log.count("retrying") == ATTEMPTS - 1
    assert os.path.exists(tmp_path / "o" / "metadata" / "config_auto.ini")
