"""End-to-end driver runs on the GPU, modelled on the reference's tests/test_simulations.py (which
asserts that the output files exist) — plus numeric checks the reference never makes."""
import os
import tarfile
import textwrap

import numpy as np
import pytest

from common import O

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
BEDPE = os.path.join(GOLD, "synthetic_loops.bedpe")
BED = os.path.join(GOLD, "synthetic_subcompartments.bed")


def _ini(tmp_path, **fields):
    base = dict(PLATFORM="B200", N_BEADS=3000, LOOPS_PATH=BEDPE, OUT_PATH=str(tmp_path / "out"), SAVE_PLOTS=False,
                SIM_RUN_MD=False, MIN_MAX_ITERATIONS=300)
    base.update(fields)
    path = tmp_path / "config.ini"
    path.write_text("[Main]\n" + "".join(f"{k} = {v}\n" for k, v in base.items()))
    return str(path), base["OUT_PATH"]


def test_cli_genome_wide_run(built_lib, tmp_path):
    from multimm_b200 import cif, run

    ini, out = _ini(tmp_path, COMPARTMENT_PATH=BED, SCB_USE_SUBCOMPARTMENT_BLOCKS=True, SC_USE_SPHERICAL_CONTAINER=True,
                    CHB_USE_CHROMOSOMAL_BLOCKS=True, IBL_USE_B_LAMINA_INTERACTION=True, CF_USE_CENTRAL_FORCE=True,
                    SHUFFLE_CHROMS=True)
    assert run.main(["-c", ini]) == 0
    for rel in ("model/MultiMM_minimized.cif", "metadata/MultiMM_init.cif", "metadata/MultiMM.psf",
                "metadata/config_auto.ini", "metadata/output.log", "metadata/parameters.txt"):
        assert os.path.exists(os.path.join(out, rel)), rel
    chroms = os.listdir(os.path.join(out, "model", "chromosomes"))
    assert len(chroms) == 22 and all(c.startswith("MultiMM_minimized_chr") for c in chroms)
    x0 = cif.read_cif_coordinates(os.path.join(out, "metadata/MultiMM_init.cif"), include_hetatm=True)
    x1 = cif.read_cif_coordinates(os.path.join(out, "model/MultiMM_minimized.cif"), include_hetatm=True)
    assert x0.shape == x1.shape == (3000, 3)
    # the start is the Hilbert lattice in Angstrom (0.1 nm spacing), and the structure moved
    assert np.array_equal(x0, O.hilbert_points(3000, 8).astype(float))
    assert np.abs(x1 - x0).max() > 0.1


def test_cli_region_run_other_start(built_lib, tmp_path):
    from multimm_b200 import run

    ini, out = _ini(tmp_path, N_BEADS=500, CHROM="chr1", LOC_START=10000000, LOC_END=60000000,
                    INITIAL_STRUCTURE_TYPE="helix", SAVE_PLOTS=True)
    assert run.main(["-c", ini]) == 0
    assert os.path.exists(os.path.join(out, "model/MultiMM_minimized.cif"))
    # SAVE_PLOTS: the structure reports of make_plots (numbers, no figures)
    for name in ("initial_structure", "minimized_structure"):
        rep = open(os.path.join(out, "analysis", f"{name}_report.txt")).read()
        assert rep.startswith("===== STRUCTURE ANALYSIS =====") and "Mean distance:" in rep
        assert os.path.exists(os.path.join(out, "analysis", f"{name}_curves.npz"))


def test_cli_platform_names(built_lib, tmp_path):
    """An ini written for the reference (PLATFORM = OpenCL / CPU) runs here, on the GPU; a name that is
    no platform at all is refused before any work is done."""
    from multimm_b200 import run

    ini, out = _ini(tmp_path, PLATFORM="TPU")
    assert run.main(["-c", ini]) == 1
    assert not os.path.exists(os.path.join(out, "model", "MultiMM_minimized.cif"))
    ini, out = _ini(tmp_path, PLATFORM="OpenCL", MIN_MAX_ITERATIONS=20)
    assert run.main(["-c", ini]) == 0
    assert os.path.exists(os.path.join(out, "model", "MultiMM_minimized.cif"))


def test_ensemble_members_differ_and_are_archived(built_lib, tmp_path):
    from multimm_b200 import run

    ini, out = _ini(tmp_path, N_BEADS=4000, GENERATE_ENSEMBLE=True, N_ENSEMBLE=3, SHUFFLE_CHROMS=True,
                    COMPARTMENT_PATH=BED, SCB_USE_SUBCOMPARTMENT_BLOCKS=True, MIN_MAX_ITERATIONS=100)
    args, _ = run.get_config(["-c", ini])
    reports = run.run_ensemble(args, devices=[0])
    assert [r["replica"] for r in reports] == [0, 1, 2]
    for i in range(3):
        tar = os.path.join(out, f"run_{i}.tar.gz")
        assert tarfile.is_tarfile(tar) and not os.path.exists(os.path.join(out, f"run_{i}"))
        with tarfile.open(tar) as t:
            assert f"run_{i}/model/MultiMM_minimized.cif" in t.getnames()
    # different SHUFFLING_SEED -> different chromosome order -> different energies
    assert len({round(r["e_final"], 3) for r in reports}) == 3


def test_bridge_in_process(built_lib, tmp_path):
    from multimm_b200.bridge import SimulationEngine

    assert "N_BEADS" in SimulationEngine.get_schema()["properties"]
    params = dict(PLATFORM="B200", N_BEADS=3000, LOOPS_PATH=BEDPE, OUT_PATH=str(tmp_path / "b"), SAVE_PLOTS=False,
                  SIM_RUN_MD=False, MIN_MAX_ITERATIONS=50)
    auto = SimulationEngine.run_in_process(params)
    assert os.path.exists(auto) and os.path.exists(str(tmp_path / "b" / "model" / "MultiMM_minimized.cif"))
