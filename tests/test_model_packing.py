"""Host mirror of the force-field builder (model.py:164-857) without a GPU: a recording stand-in
for the engine captures what MultiMM.add_* would hand to the C-ABI."""
import os

import numpy as np
import pytest

from multimm_b200 import loaders, model
from multimm_b200.config import SimulationConfig

GOLD = os.path.join(os.path.dirname(__file__), "golden")
BEDPE = os.path.join(GOLD, "synthetic_loops.bedpe")
BED = os.path.join(GOLD, "synthetic_subcompartments.bed")


class Recorder:
    def __init__(self, n_beads, device=0):
        self.n, self.device, self.calls = n_beads, device, []

    def __getattr__(self, name):
        def rec(*a, **k):
            self.calls.append((name, a, k))
            if name == "minimize":
                return dict(iterations=0, evaluations=1, e_initial=0.0, e_final=0.0, rms_force=0.0, wall_seconds=0.0,
                            converged=1, ls_status=0)
            if name == "get_positions":
                return np.zeros((self.n, 3))
            if name == "hilbert_points":
                from multimm_b200 import structures
                return structures.hilbert_points_host(self.n, 8)
        return rec

    def named(self, name):
        return [(a, k) for n, a, k in self.calls if n == name]


@pytest.fixture()
def gw(tmp_path, monkeypatch):
    monkeypatch.setattr(model, "Engine", Recorder)
    args = SimulationConfig(PLATFORM="B200", N_BEADS=6000, LOOPS_PATH=BEDPE, COMPARTMENT_PATH=BED,
                            OUT_PATH=str(tmp_path / "o"), SAVE_PLOTS=False, SHUFFLE_CHROMS=True, SHUFFLING_SEED=3,
                            SC_USE_SPHERICAL_CONTAINER=True, CHB_USE_CHROMOSOMAL_BLOCKS=True,
                            SCB_USE_SUBCOMPARTMENT_BLOCKS=True, IBL_USE_B_LAMINA_INTERACTION=True,
                            CF_USE_CENTRAL_FORCE=True)
    m = model.MultiMM(args)
    m.set_radiuses()
    m.initialize_simulation()
    m.add_forcefield()
    return m


def test_radii(gw):
    """model.py:1016-1067; SURVEY 8(a) row R: N = 1e4 -> R2 = 2.154, R1 = 1.260."""
    assert gw.radius2 == pytest.approx(0.1 * 6000 ** (1 / 3)) and gw.r_comp == pytest.approx(0.15)
    assert gw.radius1 == pytest.approx(gw.radius2 * 0.2 ** (1 / 3))
    assert 0.1 * 1e4 ** (1 / 3) == pytest.approx(2.154, abs=1e-3)


def test_terms_are_added_in_the_reference_order_with_the_reference_parameters(gw):
    eng = gw.engine
    order = [n for n, _, _ in eng.calls if n.startswith("set_") and n not in ("set_positions",)]
    assert order == ["set_bead_params", "set_pair_term", "set_pair_term", "set_pair_term", "set_external_term",
                     "set_external_term", "set_external_term", "set_bonds", "set_loops", "set_angles"]
    pair = eng.named("set_pair_term")
    assert [a[0] for a, _ in pair] == ["EV", "SCB", "CHB"]  # COB is off in config_gw.ini
    # EV: sigma is the LOOP bond length (model.py:175), not the polymer bond length
    assert pair[0][0][1] == 0 and list(pair[0][0][2]) == [100.0, 0.05, 0.1, 6.0]
    assert list(pair[1][0][2]) == [pytest.approx(0.15), 1.0, 1.33, 1.66, 2.0]
    assert list(pair[2][0][2]) == [0.3, 1e-4]
    ext = eng.named("set_external_term")
    assert [a[0] for a, _ in ext] == ["SC", "LAM", "CF"]
    assert list(ext[0][0][2][:3]) == [1000.0, gw.radius1, gw.radius2] and np.allclose(ext[0][0][2][3:], gw.mass_center)
    assert ext[1][0][2][0] == 400.0 and ext[2][0][2][:2] == [20.0, gw.radius1]


def test_topology_quirks(gw):
    """model.py:628-635 / 711-719: bond i skipped for i in chr_ends (bead 0 is unbonded, (e_k - 1, e_k)
    joins consecutive chromosomes); angle i skipped for i in chr_ends or chr_ends - 1."""
    n, ce = 6000, np.asarray(gw.chr_ends)
    (bi, bj, r0, k), _ = gw.engine.named("set_bonds")[0]
    assert len(bi) == n - 1 - (len(ce) - 1) and 0 not in bi and np.array_equal(bj, bi + 1)
    assert all((e - 1) in bi and e not in bi for e in ce[1:-1])
    assert r0 == pytest.approx(0.1) and k == pytest.approx(3.0e5)
    (ai, aj, ak, t0, kt), _ = gw.engine.named("set_angles")[0]
    assert not np.isin(ai, ce).any() and not np.isin(ai, ce - 1).any()
    assert len(ai) == (n - 2) - (len(ce) - 1) - (len(ce) - 2)
    assert t0 == pytest.approx(np.pi) and kt == pytest.approx(100.0)


def test_bead_params_follow_the_shuffled_order(gw):
    """chrom_spin is the chromosome id per bead; chrom_strength is indexed by POSITION in the shuffled
    order (model.py:158-162, appendix A Q8)."""
    (s, chrom, cstr), _ = gw.engine.named("set_bead_params")[0]
    ce = gw.chr_ends
    assert len(set(gw.chrom_idxs)) == 22 and list(gw.chrom_idxs) != sorted(gw.chrom_idxs)  # shuffled
    for kk in range(len(ce) - 1):
        assert (chrom[ce[kk]:ce[kk + 1]] == gw.chrom_idxs[kk]).all()
        assert (cstr[ce[kk]:ce[kk + 1]] == loaders.CHROM_STRENGTH[kk]).all()
    assert set(np.unique(s)) <= {-2, -1, 0, 1, 2}


def test_loop_distances(gw, tmp_path, monkeypatch):
    (lm, ln, r0, k), kw = gw.engine.named("set_loops")[0]
    assert np.array_equal(lm, gw.ms) and np.array_equal(ln, gw.ns) and kw["form"] == 0
    assert np.array_equal(r0, np.asarray(gw.ds, float)) and k == pytest.approx(3.0e4)
    assert (ln > lm + 2).all() and r0.min() >= 0.1 - 1e-12 and r0.max() <= 0.2 + 1e-12
    args = SimulationConfig(PLATFORM="B200", N_BEADS=6000, LOOPS_PATH=BEDPE, OUT_PATH=str(tmp_path / "p"),
                            SAVE_PLOTS=False, LE_FIXED_DISTANCES=True)
    m = model.MultiMM(args)
    m.set_radiuses(); m.initialize_simulation(); m.add_forcefield()
    (_, _, r0f, _), _ = m.engine.named("set_loops")[0]
    assert np.all(r0f == 0.1)
    assert [a[0] for a, _ in m.engine.named("set_pair_term")] == ["EV"]


def test_unknown_forms_raise_value_error_like_the_reference(tmp_path, monkeypatch):
    monkeypatch.setattr(model, "Engine", Recorder)
    for field, val, msg in (("EV_FORCE_TYPE", "soft_lj", "Unknown EV_FORCE_TYPE"),
                            ("CHB_FORCE_TYPE", "cubic", "Unknown CHB_FORCE_TYPE"),
                            ("LE_LOOP_FORCE_TYPE", "spring", "Unknown loop force type")):
        args = SimulationConfig(PLATFORM="B200", N_BEADS=6000, LOOPS_PATH=BEDPE, OUT_PATH=str(tmp_path / field),
                                SAVE_PLOTS=False, CHB_USE_CHROMOSOMAL_BLOCKS=True, **{field: val})
        m = model.MultiMM(args)
        m.set_radiuses(); m.initialize_simulation()
        with pytest.raises(ValueError, match=msg):
            m.add_forcefield()


def test_compartment_forces_need_a_compartment_file(tmp_path, monkeypatch):
    monkeypatch.setattr(model, "Engine", Recorder)
    args = SimulationConfig(PLATFORM="B200", N_BEADS=6000, LOOPS_PATH=BEDPE, OUT_PATH=str(tmp_path / "q"),
                            SAVE_PLOTS=False, SCB_USE_SUBCOMPARTMENT_BLOCKS=True)
    m = model.MultiMM(args)
    m.set_radiuses(); m.initialize_simulation()
    with pytest.raises(ValueError, match="COMPARTMENT_PATH"):
        m.add_forcefield()


def test_openmm_platform_names_run_on_the_engine(tmp_path, monkeypatch, caplog):
    """Every ini the reference ships says PLATFORM = OpenCL and its default is CPU: those names are a
    preference (model.py:862-871), not a request for a CPU path; unknown names are refused."""
    monkeypatch.setattr(model, "Engine", Recorder)
    for name in ("OpenCL", "CPU", "CUDA", "Reference", "B200"):
        args = SimulationConfig(PLATFORM=name, N_BEADS=6000, LOOPS_PATH=BEDPE, OUT_PATH=str(tmp_path / name), SAVE_PLOTS=False)
        m = model.MultiMM(args)
        m.set_radiuses()
        with caplog.at_level("WARNING"):
            caplog.clear()
            m.initialize_simulation()
        assert isinstance(m.engine, Recorder)
        assert (name == "B200") == (not any("OpenMM platform name" in r.message for r in caplog.records))
    args = SimulationConfig(PLATFORM="TPU", N_BEADS=6000, LOOPS_PATH=BEDPE, OUT_PATH=str(tmp_path / "t"), SAVE_PLOTS=False)
    m = model.MultiMM(args)
    m.set_radiuses()
    with pytest.raises(ValueError, match="B200 only"):
        m.initialize_simulation()


def test_gene_level_region(tmp_path, monkeypatch):
    """model.py:65-97: the gene table gives (chromosome, window, gene); gene_start / gene_end are the
    gene's bead range inside the window.  The table itself (data/hg38_gtf_annotations.tsv, 4 MB) is
    the reference's data file and is passed by path (GENE_TSV)."""
    monkeypatch.setattr(model, "Engine", Recorder)
    tsv = tmp_path / "genes.tsv"
    tsv.write_text("gene_id\tgene_name\tchromosome\tstart\tend\n"
                   "ENSG0001\tAAA\tchr1\t30000000\t30600000\n"
                   "ENSG0002\tBBB\tchr2\t5000\t90000\n")
    assert loaders.get_gene_region(str(tsv), gene_name="BBB", window_size=100000) == ("chr2", [0, 190000], [5000, 90000])
    with pytest.raises(ValueError, match="not found"):
        loaders.get_gene_region(str(tsv), gene_id="ENSG9999")
    with pytest.raises(ValueError, match="must be provided"):
        loaders.get_gene_region(str(tsv))
    args = SimulationConfig(PLATFORM="B200", N_BEADS=1000, LOOPS_PATH=BEDPE, OUT_PATH=str(tmp_path / "g"), SAVE_PLOTS=False,
                            MODELLING_LEVEL="gene", GENE_TSV=str(tsv), GENE_NAME="AAA", GENE_WINDOW=20_000_000)
    m = model.MultiMM(args)
    span = 40_600_000
    assert m.gene_start == (20_000_000 * 1000) // span and m.gene_end == (20_600_000 * 1000) // span
    assert list(m.chr_ends) == [0, 1000] and len(m.ms) > 0


def test_gene_level_without_the_table_fails_early(tmp_path, monkeypatch):
    """The default GENE_TSV points at package data this repo does not ship: a gene-level run says so
    before it creates the output tree or parses anything (it used to die inside pandas)."""
    monkeypatch.setattr(model, "Engine", Recorder)
    out = tmp_path / "g2"
    args = SimulationConfig(PLATFORM="B200", N_BEADS=1000, LOOPS_PATH=BEDPE, OUT_PATH=str(out), SAVE_PLOTS=False,
                            MODELLING_LEVEL="gene", GENE_NAME="AAA")
    with pytest.raises(ValueError, match="GENE_TSV"):
        model.MultiMM(args)
    assert not out.exists()


def test_run_is_prepare_compute_finish_and_writes_the_minimised_structure_files(tmp_path, monkeypatch):
    """run() = prepare() -> compute() -> finish() (model.py:1216-1248): the engine is driven in that order,
    model/MultiMM_minimized.cif and the per-chromosome files hold the MINIMISED structure even when an MD
    relaxation moved the beads afterwards (the reference saves the chromosomes before MD, model.py:1222-1226),
    and parameters.txt is written last."""
    from multimm_b200 import cif

    class Moving(Recorder):
        """positions: 1 nm after the minimisation, 2 nm after MD"""
        def __getattr__(self, name):
            base = super().__getattr__(name)
            if name == "get_positions":
                def pos(*a, **k):
                    base(*a, **k)
                    md_done = any(n == "md_run" for n, _, _ in self.calls)
                    return np.full((self.n, 3), 2.0 if md_done else 1.0)
                return pos
            if name == "md_run":
                def md_run(*a, **k):
                    base(*a, **k)
                    return dict(step=0, potential=0.0, kinetic=0.0, temperature=0.0)
                return md_run
            return base

    monkeypatch.setattr(model, "Engine", Moving)
    out = tmp_path / "o"
    args = SimulationConfig(PLATFORM="B200", N_BEADS=3000, LOOPS_PATH=BEDPE, OUT_PATH=str(out), SAVE_PLOTS=False,
                            SIM_RUN_MD=True, SIM_N_STEPS=20, SIM_SAMPLING_STEP=10, TRJ_FRAMES=2)
    m = model.MultiMM(args)
    rep = m.run()
    assert rep["converged"] == 1
    names = [n for n, _, _ in m.engine.calls]
    assert names.index("hilbert_init") < names.index("set_pair_term") < names.index("minimize") < names.index("md_run")
    mini = cif.read_cif_coordinates(str(out / "model/MultiMM_minimized.cif"), include_hetatm=True)
    after = cif.read_cif_coordinates(str(out / "model/MultiMM_afterMD.cif"), include_hetatm=True)
    assert np.allclose(mini, 10.0) and np.allclose(after, 20.0)  # Angstrom
    chrom_files = sorted(os.listdir(out / "model" / "chromosomes"))
    assert chrom_files, "per-chromosome files of the minimised structure"
    one = cif.read_cif_coordinates(str(out / "model" / "chromosomes" / chrom_files[0]), include_hetatm=True)
    assert np.allclose(one, 10.0)
    assert (out / "metadata" / "parameters.txt").exists()
    # the same three stages called one by one (what the ensemble driver does) leave the same files
    out2 = tmp_path / "o2"
    m2 = model.MultiMM(SimulationConfig(**{**args.model_dump(), "OUT_PATH": str(out2)}))
    m2.prepare()
    assert not (out2 / "model/MultiMM_minimized.cif").exists()
    m2.compute()
    assert not (out2 / "model/MultiMM_minimized.cif").exists()  # left to finish(): written while the next member minimises
    m2.finish()
    for rel in ("model/MultiMM_minimized.cif", "model/MultiMM_afterMD.cif", "model/chromosomes/" + chrom_files[0]):
        assert (out / rel).read_bytes() == (out2 / rel).read_bytes(), rel
    assert (out2 / "metadata" / "parameters.txt").exists()
