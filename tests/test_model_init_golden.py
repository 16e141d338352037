"""MultiMM.__init__ (input ingestion, chromosome spins / strengths, gene window) against the
REFERENCE's own __init__ run on the same inputs (tests/golden/make_golden_model_init.py)."""
import os

import numpy as np
import pytest

from multimm_b200 import model
from multimm_b200.config import SimulationConfig

GOLD = os.path.join(os.path.dirname(__file__), "golden")
G = np.load(os.path.join(GOLD, "model_init_golden.npz"))
BEDPE = os.path.join(GOLD, "synthetic_loops.bedpe")
BED = os.path.join(GOLD, "synthetic_subcompartments.bed")
GENES = "gene_id\tgene_name\tchromosome\tstart\tend\nENSG01\tAAA\tchr1\t30000000\t30600000\nENSG02\tBBB\tchr2\t50000000\t50090000\n"

CASES = {
    "gw_shuffle": dict(N_BEADS=20000, COMPARTMENT_PATH=BED, SHUFFLE_CHROMS=True, SHUFFLING_SEED=3),
    "gw_plain_downsampled": dict(N_BEADS=8000, COMPARTMENT_PATH=BED, SHUFFLING_SEED=1, DOWNSAMPLING_PROB=0.8,
                                 COMPARTMENT_FLIP_PROB=0.2),
    "chr1_region": dict(N_BEADS=2000, CHROM="chr1", LOC_START=10_000_000, LOC_END=110_000_000, COMPARTMENT_PATH=BED),
    "chr6_no_comps": dict(N_BEADS=3000, CHROM="chr6", LOC_START=0, LOC_END=172126628),
    "gene_by_name": dict(N_BEADS=1000, MODELLING_LEVEL="gene", GENE_NAME="AAA", GENE_WINDOW=20_000_000),
    "gene_by_id": dict(N_BEADS=1000, MODELLING_LEVEL="gene", GENE_ID="ENSG02", GENE_WINDOW=30_000_000),
}


@pytest.mark.parametrize("name", list(CASES))
def test_init_matches_the_reference(name, tmp_path):
    tsv = tmp_path / "genes.tsv"
    tsv.write_text(GENES)
    args = SimulationConfig(PLATFORM="B200", LOOPS_PATH=BEDPE, OUT_PATH=str(tmp_path / name), GENE_TSV=str(tsv),
                            **CASES[name])
    m = model.MultiMM(args)
    for attr in ("ms", "ns", "chr_ends", "chrom_idxs", "Cs", "chrom_spin", "gene_start", "gene_end"):
        key = f"{name}.{attr}"
        if key in G:
            assert np.array_equal(np.asarray(getattr(m, attr)), G[key]), attr
        else:
            assert getattr(m, attr, None) is None, attr
    assert np.allclose(m.ds, G[f"{name}.ds"], rtol=1e-13, atol=0)
    assert np.allclose(m.chrom_strength, G[f"{name}.chrom_strength"], rtol=0, atol=0)
    # the output tree of model.py:46-55
    for sub in ("md_frames", "plots", "metadata", "model"):
        assert os.path.isdir(os.path.join(args.OUT_PATH, sub))
    whole = name.startswith("gw")
    assert os.path.isdir(os.path.join(args.OUT_PATH, "model", "chromosomes")) == whole
