"""Structure report against the text the REFERENCE's analyze_structure wrote for the same inputs
(tests/golden/make_golden_analysis.py)."""
import os

import numpy as np
import pytest

from multimm_b200 import analysis

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["walk400", "helix250"])
def test_report_is_identical_to_the_reference(tmp_path, name):
    V = np.load(os.path.join(GOLD, f"analysis_{name}_input.npy"))
    analysis.analyze_structure(V, str(tmp_path), name=name)
    ours = (tmp_path / "analysis" / f"{name}_report.txt").read_text()
    with open(os.path.join(GOLD, f"analysis_{name}_report.txt")) as f:
        assert ours == f.read()


def test_local_rg_prefix_sums_match_the_reference_loop():
    rng = np.random.default_rng(3)
    V = np.cumsum(rng.normal(size=(700, 3)), axis=0)
    window = 10
    ref = []
    for i in range(len(V) - window):  # plots.py:715-718
        chunk = V[i:i + window]
        cm = np.mean(chunk, axis=0)
        ref.append(np.sqrt(np.mean(np.sum((chunk - cm) ** 2, axis=1))))
    assert np.allclose(analysis.local_rg(V, window), ref, rtol=1e-9, atol=1e-9)


@pytest.mark.gpu
def test_device_mean_distance_matches_cdist(built_lib, tmp_path):
    from scipy.spatial import distance

    from multimm_b200.engine import Engine

    rng = np.random.default_rng(5)
    for n in (3, 257, 5000):
        V = np.cumsum(rng.normal(0, 0.1, size=(n, 3)), axis=0)
        eng = Engine(n)
        eng.set_positions(V)
        got = eng.mean_pair_distance()
        assert got == pytest.approx(np.mean(distance.cdist(V, V)), rel=2e-6)
        eng.close()
    # the report through the engine is the same text
    V = np.load(os.path.join(GOLD, "analysis_walk400_input.npy"))
    eng = Engine(len(V))
    analysis.analyze_structure(V, str(tmp_path), name="walk400", engine=eng)
    eng.close()
    with open(os.path.join(GOLD, "analysis_walk400_report.txt")) as f:
        assert (tmp_path / "analysis" / "walk400_report.txt").read_text() == f.read()


def test_contact_map_block_means_never_need_the_full_matrix():
    """contact_map(bins=B) equals the block means of the N x N matrix get_heatmap draws (plots.py:540-561)."""
    rng = np.random.default_rng(1)
    V = np.cumsum(rng.normal(size=(301, 3)), axis=0)
    for kw in (dict(), dict(log_scale=False), dict(reorder_by_diagonal=True)):
        full = analysis.contact_map(V, **kw)
        assert full.shape == (301, 301) and np.allclose(full, full.T)
        want_diag = np.log1p(1.0) if kw.get("log_scale", True) else 1.0
        assert np.allclose(np.diag(full), want_diag)
        for bins in (1, 9, 64, 301, 5000):
            cg = analysis.contact_map(V, bins=bins, **kw)
            b = min(bins, 301)
            edges = np.linspace(0, 301, b + 1).astype(int)
            ref = np.array([[full[edges[p]:edges[p + 1], edges[q]:edges[q + 1]].mean() for q in range(b)] for p in range(b)])
            assert cg.shape == (b, b) and np.allclose(cg, ref, rtol=1e-12, atol=1e-14)


@pytest.mark.gpu
def test_device_contact_map_matches_host(built_lib):
    """Block-mean contact map + mean pair distance in one device pass (mmm_contact_map) against the
    host restatement of get_heatmap (plots.py:540-561), for an array whose length is no multiple of
    anything (the report reads back CIFs without their HETATM rows)."""
    from multimm_b200 import analysis

    rng = np.random.default_rng(5)
    V = np.cumsum(rng.normal(0.0, 0.6, size=(3001, 3)), axis=0)
    for bins, log_scale in ((37, True), (3001, True), (64, False)):
        got, mean = analysis.device_pair_stats(V, bins, device=0, log_scale=log_scale)
        want = analysis.contact_map(V, log_scale=log_scale, bins=None if bins == len(V) else bins)
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max() + 1e-13
        assert abs(mean - analysis.mean_pair_distance_host(V)) <= 1e-12 * mean
