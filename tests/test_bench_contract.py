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


def test_reference_arm_threads_are_set_explicitly(built_lib):
    """torchrun exports OMP_NUM_THREADS=1 (round 1's SCALE reference arm ran on one core while
    labelled 32): the oracle's thread count is passed explicitly and `cores` is what was measured."""
    res = run_bench("--impl", "reference", "--workload", "region", "--steps", "1", "--warmup", "0", "--cpu-seconds", "1",
                    env=dict(OMP_NUM_THREADS="1"))
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][0])
    host = len(os.sched_getaffinity(0))
    assert out["cpu_baseline"]["cores"] == host and out["cpu_baseline"]["omp_num_threads_env"] == "1"
    # one REAL evaluation of the full system rides beside the extrapolated steps
    assert out["full_evaluation"]["beads"] == 10000 and out["full_evaluation"]["seconds"] > 0
    assert out["full_evaluation"]["cores"] == host
    assert "EXTRAPOLATED" in out["config"]["note"]


def test_openmm_arm_control_flow_with_stand_in_modules(monkeypatch):
    """OpenMM cannot be installed in this image, so the arm that times the unmodified reference on it
    (bench_openmm.py) is exercised against stand-in `openmm` / `multimm` modules: every call it makes
    on the reference objects exists in model.py:722-886, and the numbers it reports are wired."""
    import types

    sys.path.insert(0, ROOT)
    import bench_openmm

    calls = []

    class Q:
        def __init__(self, v): self.v = v
        def value_in_unit(self, _u): return self.v

    class State:
        def getPotentialEnergy(self): return Q(-12.5)

    class Platform:
        def __init__(self, name): self.name = name
        def getName(self): return self.name
        def setPropertyDefaultValue(self, k, v): calls.append(("prop", k, v))
        @staticmethod
        def getPlatformByName(name): return Platform(name)
        @staticmethod
        def getNumPlatforms(): return 2
        @staticmethod
        def getPlatform(i): return Platform(("Reference", "CPU")[i])

    class Context:
        def __init__(self, platform): self.platform = platform
        def setPositions(self, p): calls.append("setPositions")
        def getState(self, **kw): calls.append(("getState", tuple(sorted(kw)))); return State()
        def getPlatform(self): return self.platform

    class Simulation:
        def __init__(self, topology, system, integrator, platform): self.context = Context(platform)
        def minimizeEnergy(self, reporter=None):
            calls.append("minimizeEnergy")
            if reporter is not None:
                assert reporter.report(3, None, None, {}) is False

    class MinimizationReporter:
        pass

    mm = types.ModuleType("openmm")
    mm.Platform, mm.MinimizationReporter = Platform, MinimizationReporter
    mm.unit = types.SimpleNamespace(kilojoule_per_mole="kJ/mol")
    app = types.ModuleType("openmm.app")
    app.Simulation = Simulation
    pkg = types.ModuleType("multimm")
    cfg = types.ModuleType("multimm.config")
    cfg.SimulationConfig = lambda **kw: types.SimpleNamespace(**kw)
    mdl = types.ModuleType("multimm.model")

    class RefMultiMM:
        def __init__(self, args): self.args = args
        def set_radiuses(self): calls.append("set_radiuses")
        def initialize_simulation(self):
            calls.append("initialize_simulation")
            self.pdb = types.SimpleNamespace(topology="top", positions="pos")
            self.system, self.integrator = "system", "integrator"
        def add_forcefield(self): calls.append("add_forcefield")

    mdl.MultiMM = RefMultiMM
    for name, mod in (("openmm", mm), ("openmm.app", app), ("multimm", pkg), ("multimm.config", cfg), ("multimm.model", mdl)):
        monkeypatch.setitem(sys.modules, name, mod)
    assert bench_openmm.available()[0] is True
    assert bench_openmm.platforms() == ["Reference", "CPU"]
    out = bench_openmm.time_arm(dict(N_BEADS=10, LOOPS_PATH="x.bedpe"), "CPU", steps=3, warmup=1, threads=4, minimize_cap_s=5.0)
    assert calls[:3] == ["set_radiuses", "initialize_simulation", "add_forcefield"]
    assert ("prop", "Threads", "4") in calls and "minimizeEnergy" in calls
    assert calls.count(("getState", ("getEnergy", "getForces"))) == 4  # 1 warm-up + 3 timed
    assert out["platform"] == "CPU" and out["threads"] == 4 and out["force_evals_per_s"] > 0
    assert out["minimize"]["energy_final_kj_mol"] == -12.5 and out["minimize"]["capped"] is False
    assert out["minimize"]["iterations_seen"] == 3
