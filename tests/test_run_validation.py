"""Config validation and CLI / ensemble host logic, modelled on the reference's
tests/test_run_validation.py (12 tests, 4 classes) plus the ensemble plumbing the reference never tests."""
import os
import sys
import tarfile
import textwrap

import pytest
from pydantic import ValidationError

from multimm_b200 import run
from multimm_b200.config import SimulationConfig
from multimm_b200.run import args_tests

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BEDPE = os.path.join(os.path.dirname(__file__), "golden", "synthetic_loops.bedpe")
BED = os.path.join(os.path.dirname(__file__), "golden", "synthetic_subcompartments.bed")


def make_config(**kwargs):
    defaults = dict(LOOPS_PATH=BEDPE, OUT_PATH="/tmp/mmm_b200_output")
    defaults.update(kwargs)
    return SimulationConfig(**defaults)


class TestRequiredPaths:
    def test_missing_loops_raises(self):
        with pytest.raises(ValidationError):
            SimulationConfig(LOOPS_PATH=None, OUT_PATH="/tmp/output")

    def test_empty_loops_raises(self):
        with pytest.raises(ValidationError):
            SimulationConfig(LOOPS_PATH="", OUT_PATH="/tmp/output")

    def test_args_tests_with_valid_path_passes(self):
        args_tests(make_config())

    def test_nonexistent_file_raises(self):
        with pytest.raises(ValueError, match="not found"):
            args_tests(make_config(LOOPS_PATH="/some/file.bedpe"))


class TestCompartmentConflicts:
    def test_compartment_blocks_without_bed_raises(self):
        with pytest.raises(ValueError, match="COB_USE_COMPARTMENT_BLOCKS"):
            args_tests(make_config(COB_USE_COMPARTMENT_BLOCKS=True))

    def test_subcompartment_blocks_without_bed_raises(self):
        with pytest.raises(ValueError, match="SCB_USE_SUBCOMPARTMENT_BLOCKS"):
            args_tests(make_config(SCB_USE_SUBCOMPARTMENT_BLOCKS=True))

    def test_lamina_without_compartments_raises(self):
        with pytest.raises(ValueError):
            args_tests(make_config(IBL_USE_B_LAMINA_INTERACTION=True))

    def test_lamina_without_compartment_force_raises(self):
        with pytest.raises(ValueError):
            args_tests(make_config(IBL_USE_B_LAMINA_INTERACTION=True, COMPARTMENT_PATH=BED))

    def test_lamina_with_compartment_force_passes(self):
        args_tests(make_config(IBL_USE_B_LAMINA_INTERACTION=True, COMPARTMENT_PATH=BED, COB_USE_COMPARTMENT_BLOCKS=True))


class TestNucleosomeConflicts:
    def test_nuc_interpolation_without_atacseq_raises(self):
        with pytest.raises(ValueError, match="NUC_DO_INTERPOLATION"):
            args_tests(make_config(NUC_DO_INTERPOLATION=True))


class TestChromConflicts:
    def test_central_force_with_single_chrom_raises(self):
        with pytest.raises(ValueError, match="CF_USE_CENTRAL_FORCE"):
            args_tests(make_config(CF_USE_CENTRAL_FORCE=True, CHROM="chr1"))

    def test_chromosomal_blocks_with_single_chrom_only_warns(self):
        """The reference's own test expects a raise here, its code only warns (run.py:296-300): follow the code."""
        args_tests(make_config(CHB_USE_CHROMOSOMAL_BLOCKS=True, CHROM="chr1"))

    def test_central_force_genome_wide_passes(self):
        args_tests(make_config(CF_USE_CENTRAL_FORCE=True))


class TestPresets:
    def test_gw_preset(self):
        cfg = make_config(MODELLING_LEVEL="GW", COMPARTMENT_PATH=BED, CHB_USE_CHROMOSOMAL_BLOCKS=True, N_BEADS=777)
        run.ArgumentChanger(cfg).convenient_argument_changer()
        assert cfg.N_BEADS == 200000 and cfg.SC_USE_SPHERICAL_CONTAINER is True
        assert cfg.COB_USE_COMPARTMENT_BLOCKS is True and cfg.IBL_USE_B_LAMINA_INTERACTION is True
        assert cfg.CHB_USE_CHROMOSOMAL_BLOCKS is False and cfg.SIM_RUN_MD is False

    def test_gw_preset_without_compartments(self):
        cfg = make_config(MODELLING_LEVEL="genome")
        run.ArgumentChanger(cfg).convenient_argument_changer()
        assert cfg.COB_USE_COMPARTMENT_BLOCKS is False and cfg.IBL_USE_B_LAMINA_INTERACTION is False

    def test_chromosome_preset_sets_region(self):
        cfg = make_config(MODELLING_LEVEL="chrom", CHROM="chr2")
        run.ArgumentChanger(cfg).convenient_argument_changer()
        assert cfg.N_BEADS == 20000 and cfg.LOC_START == 1 and cfg.LOC_END == 242696752

    def test_interpolation_always_off(self):
        cfg = make_config(NUC_DO_INTERPOLATION=True, ATACSEQ_PATH="/x.bw")
        run.ArgumentChanger(cfg).convenient_argument_changer()
        assert cfg.NUC_DO_INTERPOLATION is False and cfg.ATACSEQ_PATH is None


class TestCli:
    def test_precedence_defaults_ini_cli(self, tmp_path):
        ini = tmp_path / "c.ini"
        ini.write_text(textwrap.dedent(f"""\
            [Main]
            platform = B200
            n_beads = 1234
            loops_path = {BEDPE}
            out_path = {tmp_path}/out
            ev_power = 3.0
        """))
        args, gpus = run.get_config(["-c", str(ini), "--n_beads", "4321", "--gpus", "0,1"])
        assert args.N_BEADS == 4321 and args.EV_POWER == 3.0 and gpus == "0,1"
        auto = tmp_path / "out" / "metadata" / "config_auto.ini"
        assert auto.exists()
        # the dump round-trips through the same parser
        again = SimulationConfig(**run.read_ini(str(auto)))
        assert again.N_BEADS == 4321 and again.LOOPS_PATH == BEDPE
        assert again.POL_HARMONIC_BOND_R0.md == args.POL_HARMONIC_BOND_R0.md


class TestEnsemblePlumbing:
    def test_replicas_are_dealt_round_robin(self):
        plan = run.assign_replicas(64, list(range(8)))
        assert all(len(v) == 8 for v in plan.values())
        assert plan[3] == [3, 11, 19, 27, 35, 43, 51, 59]
        assert sorted(i for v in plan.values() for i in v) == list(range(64))
        assert run.assign_replicas(3, [5]) == {5: [0, 1, 2]}

    def test_replica_paths_zero_padded_like_the_reference(self):
        assert run.replica_paths("/o", 12)[3] == "/o/run_03" and run.replica_paths("/o", 5)[4] == "/o/run_4"

    def test_archive_run(self, tmp_path):
        d = tmp_path / "run_0"
        (d / "model").mkdir(parents=True)
        (d / "model" / "x.cif").write_text("data_\n")
        tar = run.archive_run(str(d))
        assert not d.exists() and tarfile.is_tarfile(tar)
        with tarfile.open(tar) as t:
            assert "run_0/model/x.cif" in t.getnames()

    def test_ensemble_needs_a_count(self):
        with pytest.raises(ValueError):
            run.run_ensemble(make_config(GENERATE_ENSEMBLE=True), devices=[0])


class _FakeProc:
    def __init__(self, alive_polls, exitcode):
        self.alive_polls, self.exitcode = alive_polls, exitcode

    def is_alive(self):
        self.alive_polls -= 1
        return self.alive_polls >= 0


class TestWorkerLiveness:
    def test_dead_worker_does_not_hang_the_parent(self):
        """One worker reported, the other was killed (e.g. inside native code) without a word."""
        import queue as q

        box = q.Queue()
        box.put(("ok", dict(replica=0, device=0)))
        procs = [_FakeProc(0, 0), _FakeProc(1, -9)]
        results, errors = run.collect_reports(procs, box, 3, poll_seconds=0.05)
        assert [r["replica"] for r in results] == [0]
        assert sorted(e["replica"] for e in errors) == [1, 2]
        assert "exited without a report" in errors[0]["error"] and "-9" in errors[0]["error"]

    def test_all_reports_arrive(self):
        import queue as q

        box = q.Queue()
        for i in range(4):
            box.put(("ok" if i != 2 else "error", dict(replica=i, device=i % 2, error="boom")))
        results, errors = run.collect_reports([_FakeProc(5, 0)], box, 4, poll_seconds=0.05)
        assert len(results) == 3 and [e["replica"] for e in errors] == [2]


def test_coarse_stage_is_bounded(tmp_path, monkeypatch):
    """MIN_COARSE_CUTOFF: the cut-off stage never runs unbounded; the exact stage is a bounded probe that hands
    back to a tighter coarse stage when it does not converge, and runs with the user's bound (unbounded by
    default) in the last round only."""
    from multimm_b200 import model

    calls = []
    probe_converges = [True]

    class Eng:
        def __init__(self, n, device=0):
            self.n = n

        def __getattr__(self, name):
            def rec(*a, **k):
                calls.append((name, a, k))
                if name == "minimize":
                    exact_probe = k["max_iter"] == 30
                    return dict(iterations=1, evaluations=2, e_initial=1.0, e_final=0.0, rms_force=1.0, wall_seconds=0.0,
                                converged=int(probe_converges[0] or not exact_probe), ls_status=0)
                if name == "get_positions":
                    import numpy as np
                    return np.zeros((self.n, 3))
                if name == "hilbert_points":
                    from multimm_b200 import structures
                    return structures.hilbert_points_host(self.n, 8)
            return rec

    monkeypatch.setattr(model, "Engine", Eng)

    def run():
        calls.clear()
        m = model.MultiMM(make_config(PLATFORM="B200", N_BEADS=6000, OUT_PATH=str(tmp_path / "o"), SAVE_PLOTS=False,
                                      MIN_COARSE_CUTOFF=0.5))
        m.set_radiuses(); m.initialize_simulation(); m.add_forcefield(); m.min_energy()
        return m, [(n, a, k) for n, a, k in calls if n in ("set_cutoff", "minimize")]

    m, seq = run()  # the exact probe converges: one round
    assert [n for n, _, _ in seq] == ["set_cutoff", "minimize", "set_cutoff", "minimize"]
    assert seq[0][1] == (0.5,) and seq[2][1] == (0.0,)
    assert seq[1][2]["max_iter"] == 20000 and seq[1][2]["tol"] == 10.0
    assert seq[3][2]["max_iter"] == 30 and seq[3][2]["tol"] == 10.0
    assert m.timings["coarse_rounds"] == 1
    probe_converges[0] = False  # it never does: four rounds, tighter coarse tolerance each, last exact stage unbounded
    m, seq = run()
    mins = [k for n, _, k in seq if n == "minimize"]
    assert len(mins) == 8 and m.timings["coarse_rounds"] == 4
    assert [k["max_iter"] for k in mins] == [20000, 30, 20000, 30, 20000, 30, 20000, 0]
    tols = [k["tol"] for k in mins[0::2]]
    assert tols[0] == 10.0 and all(abs(b / a - 0.7) < 1e-12 for a, b in zip(tols, tols[1:]))
    assert all(k["tol"] == 10.0 for k in mins[1::2])  # the exact stage always at the user's tolerance
    assert [a for n, a, _ in seq if n == "set_cutoff"] == [(0.5,), (0.0,)] * 4


def test_ensemble_script_is_spawn_safe(monkeypatch):
    """run_ensemble starts workers with `spawn`, which re-imports the main script in every worker: the
    measuring script must do nothing at import (its unguarded first version left an 8-GPU run waiting
    for workers that had died while bootstrapping, profiles/r01_ensemble.md)."""
    import runpy

    monkeypatch.setattr(sys, "argv", ["gpu_ensemble.py"])  # an unguarded body would fail on argv[1]
    monkeypatch.chdir(ROOT)
    runpy.run_path(os.path.join(ROOT, "scripts", "gpu_ensemble.py"), run_name="__mp_main__")


def test_two_worker_ensemble_reports_instead_of_hanging(tmp_path):
    """The whole spawn plumbing on a box without a GPU: two workers start, each replica fails inside
    mmm_create (there is no CPU fallback), the failures travel back through the queue and the parent
    raises — within seconds, and naming the device error."""
    import subprocess

    script = tmp_path / "ens.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r})
        from multimm_b200 import run
        from multimm_b200.config import SimulationConfig

        if __name__ == "__main__":
            args = SimulationConfig(PLATFORM="B200", N_BEADS=20000, LOOPS_PATH={os.path.join(ROOT, 'tests', 'golden', 'synthetic_loops.bedpe')!r},
                                    OUT_PATH={str(tmp_path / 'out')!r}, SAVE_PLOTS=False, GENERATE_ENSEMBLE=True, N_ENSEMBLE=3)
            try:
                run.run_ensemble(args, devices=[0, 1])
            except RuntimeError as e:
                print("PARENT:", e)
                sys.exit(7)
        """))
    res = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=300,
                         env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
    assert res.returncode == 7, res.stdout + res.stderr
    assert "3 ensemble member(s) failed" in res.stdout
    assert "CUDA" in res.stdout


def test_unavailable_integrator_and_platform_fail_before_any_work():
    """A config asking for an integrator this engine does not have (the two variable-step ones)
    used to run the whole minimisation first and fail in run_md; a PLATFORM that is no platform at all
    used to fail after the loaders had run.  Both are cross-field checks now."""
    from multimm_b200.run import args_tests

    args_tests(make_config(SIM_RUN_MD=True, SIM_INTEGRATOR_TYPE="amd"))
    for kind in ("variable_verlet", "variable_langevin"):
        with pytest.raises(ValueError, match="SIM_INTEGRATOR_TYPE"):
            args_tests(make_config(SIM_RUN_MD=True, SIM_INTEGRATOR_TYPE=kind))
        args_tests(make_config(SIM_RUN_MD=False, SIM_INTEGRATOR_TYPE=kind))  # irrelevant without MD
    args_tests(make_config(SIM_RUN_MD=True, SIM_INTEGRATOR_TYPE="langevin", PLATFORM="OpenCL"))
    with pytest.raises(ValueError, match="PLATFORM"):
        args_tests(make_config(PLATFORM="TPU"))


def test_archiver_runs_in_the_background_and_reports_failures(tmp_path):
    """Replica k is archived while replica k + 1 runs: same tar.gz as run.py:423-445 writes inline;
    a failed archive surfaces at wait() and leaves the run directory in place."""
    import tarfile

    from multimm_b200 import run

    ok = tmp_path / "run_0"
    (ok / "model").mkdir(parents=True)
    (ok / "model" / "MultiMM_minimized.cif").write_text("data\n" * 1000)
    arch = run.Archiver()
    arch.submit(str(ok))
    assert arch.wait() >= 0.0
    assert tarfile.is_tarfile(str(ok) + ".tar.gz") and not ok.exists()
    with tarfile.open(str(ok) + ".tar.gz") as t:
        assert "run_0/model/MultiMM_minimized.cif" in t.getnames()
    run.check_archives([dict(archive=str(ok) + ".tar.gz")])
    with pytest.raises(RuntimeError, match="missing"):
        run.check_archives([dict(archive=str(tmp_path / "nope.tar.gz"))])
    arch = run.Archiver()
    arch.submit(str(tmp_path / "does_not_exist"))
    with pytest.raises(Exception):
        arch.wait()


def test_worker_environment_and_early_cuda_start(monkeypatch):
    """One visible device per ensemble worker, exported by the parent before the worker starts (the worker
    begins to initialise CUDA while it is still importing); the parent's own environment is untouched."""
    import threading

    from multimm_b200 import run

    env, local = run.worker_environment(3, base={})
    assert (env["CUDA_VISIBLE_DEVICES"], local, env["MMM_WORKER_WARMUP"], env["MMM_WORKER_LOCAL_DEVICE"]) == ("3", 0, "1", "0")
    env, local = run.worker_environment(1, base={"CUDA_VISIBLE_DEVICES": "4,6"})
    assert (env["CUDA_VISIBLE_DEVICES"], local) == ("6", 0)
    env, local = run.worker_environment(5, base={"CUDA_VISIBLE_DEVICES": "4,6"})  # nothing to narrow to
    assert (env["CUDA_VISIBLE_DEVICES"], local, env["MMM_WORKER_WARMUP"]) == ("4,6", 5, "0")
    # outside a worker nothing is started at import
    monkeypatch.delenv("MMM_WORKER_WARMUP", raising=False)
    assert run._early_cuda_start() is None
    # inside one a daemon thread is (here it fails quietly: no device, or no library)
    monkeypatch.setenv("MMM_WORKER_WARMUP", "1")
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "")
    t = run._early_cuda_start()
    assert isinstance(t, threading.Thread) and t.daemon
    t.join(timeout=60)
    assert not t.is_alive()


def test_run_ensemble_leaves_the_parents_environment_alone(tmp_path, monkeypatch):
    from multimm_b200 import run

    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "")
    monkeypatch.delenv("MMM_WORKER_WARMUP", raising=False)
    monkeypatch.delenv("MMM_WORKER_LOCAL_DEVICE", raising=False)
    args = make_config(GENERATE_ENSEMBLE=True, N_ENSEMBLE=2, OUT_PATH=str(tmp_path / "out"))
    with pytest.raises(RuntimeError, match="ensemble member"):
        run.run_ensemble(args, devices=[0, 1])  # no device: both workers report a failure
    assert os.environ.get("CUDA_VISIBLE_DEVICES") == ""
    assert "MMM_WORKER_WARMUP" not in os.environ and "MMM_WORKER_LOCAL_DEVICE" not in os.environ
