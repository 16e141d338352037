"""GPU parity: the CUDA path, called through the C-ABI, against the CPU oracle on identical
coordinates and parameters.  Bars (BASELINE.json north_star): per-term energies 1e-5 relative,
per-bead forces 1e-4 relative, topology / integer outputs bit-exact."""
import numpy as np
import pytest

from common import O, force_rel_err, make_case, to_engine, to_oracle

pytestmark = pytest.mark.gpu

E_TOL = 1e-5
F_TOL = 1e-4


def _fp32_positions(case):
    """Both sides see the same coordinates: the engine's FP32 copy is exact for these values."""
    c = case["x"].mean(axis=0)
    case["x"] = (case["x"] - c).astype(np.float32).astype(np.float64) + c
    return case


def _check(case, e_tol=E_TOL, f_tol=F_TOL):
    eng = to_engine(case)
    e, f = eng.energy_forces()
    e_ref, f_ref = O.energy_forces(to_oracle(case), case["x"])
    for t in range(10):
        scale = max(abs(e_ref[t]), 1e-12)
        assert abs(e[t] - e_ref[t]) <= e_tol * scale + 1e-9, (O.TERM_NAMES[t], e[t], e_ref[t])
    err = force_rel_err(f, f_ref)
    assert err <= f_tol, err
    eng.close()
    return e, f


@pytest.mark.parametrize("n,n_chrom", [(1000, 1), (2500, 3), (10000, 5)])
def test_all_default_terms(built_lib, n, n_chrom):
    _check(make_case(n, n_chrom=n_chrom, seed=n))


def test_non_finite_coordinates_are_refused_and_host_forces_are_minus_the_gradient(built_lib):
    """mmm_set_positions checks finiteness in the same host pass that forms the centre; the forces
    mmm_energy_forces copies to the host are negated on the device (no host pass over them)."""
    from multimm_b200 import Error

    case = make_case(700, n_chrom=2, seed=12)
    eng = to_engine(case)
    e, f = eng.energy_forces()
    for bad in (np.nan, np.inf, -np.inf):
        x = case["x"].copy()
        x[311, 1] = bad
        with pytest.raises(Error, match="non-finite"):
            eng.set_positions(x)
        with pytest.raises(Error):  # the rejected coordinates are not evaluated
            eng.energy_forces()
    eng.set_positions(case["x"])
    e2, f2 = eng.energy_forces()
    assert np.array_equal(e, e2) and np.array_equal(f, f2)
    f_ref = O.energy_forces(to_oracle(case), case["x"])[1]
    assert force_rel_err(f, f_ref) <= F_TOL  # the sign in particular
    eng.close()


def test_s1_region_terms(built_lib):
    """configs[0]: specific-region model, {bonds, angles, loops, EV}."""
    _check(make_case(10000, terms=("EV", "BOND", "LOOP", "ANGLE"), seed=3))


def test_cob_and_scb_together(built_lib):
    _check(make_case(3000, n_chrom=2, terms=("EV", "COB", "SCB", "BOND", "ANGLE"), seed=4))


def test_cob_only(built_lib):
    _check(make_case(3000, n_chrom=2, terms=("EV", "COB", "CHB"), seed=5))


def test_ragged_sizes(built_lib):
    for n in (2, 3, 31, 33, 255, 257, 1023):
        terms = ("EV", "SCB", "CHB", "SC", "LAM", "CF", "BOND", "ANGLE") if n >= 40 else ("EV", "SCB", "CHB", "SC", "LAM", "CF")
        case = make_case(n, n_chrom=1 if n < 40 else 2, seed=n, terms=terms)
        _check(case)


def test_ev_power_3(built_lib):
    _check(make_case(2000, terms=("EV", "BOND"), ev_power=3.0, seed=6))


def test_generic_forms(built_lib):
    """Non-default functional forms go through the generic pair path (model.py alternates)."""
    for forms in ({"EV": 1}, {"COB": 1, "SCB": 1}, {"COB": 2, "SCB": 2}, {"CHB": 1}, {"CHB": 2},
                  {"LAM": 1, "CF": 1, "LOOP": 1}, {"LAM": 2, "CF": 2, "LOOP": 2}, {"LAM": 3}):
        case = make_case(1500, n_chrom=3, seed=11, forms=forms, chb_de=1.0,
                         terms=("EV", "COB", "SCB", "CHB", "SC", "LAM", "CF", "BOND", "LOOP", "ANGLE"))
        _check(case)  # the north star's bars: the generic body is FP64 (mmm_pairmath.cuh)


def test_non_integer_ev_power(built_lib):
    _check(make_case(1500, terms=("EV",), ev_power=4.5, seed=7))


@pytest.mark.parametrize("forms,ev_power", [({"EV": 1, "COB": 1, "SCB": 2, "CHB": 1}, 6.0), ({"COB": 2, "CHB": 2}, 4.5),
                                             ({"COB": 1, "SCB": 1}, 6.0)])
def test_generic_forms_on_the_newton3_machinery(built_lib, forms, ev_power):
    """Every non-default form runs on the Newton-3 work items with the generic FP64 body (each unordered
    pair once) and agrees with the gather kernel (each pair from both sides) and with the oracle — across
    several i-blocks and a ragged last one, so that diagonal stages (ordered pairs, where the s1-only Yukawa
    COB of model.py:262-266 needs the lower index) and off-diagonal ones are both exercised."""
    case = make_case(2300, n_chrom=3, seed=21, forms=forms, chb_de=1.0, ev_power=ev_power,
                     terms=("EV", "COB", "SCB", "CHB", "BOND"))
    eng = to_engine(case)
    e, f = eng.energy_forces()
    assert eng.pair_kernel_in_use == 2
    e_again, f_again = eng.energy_forces()
    assert np.array_equal(e, e_again) and np.array_equal(f, f_again)  # fixed-point accumulation: reproducible
    eng.set_pair_kernel(1)
    eg, fg = eng.energy_forces()
    assert eng.pair_kernel_in_use == 1
    eng.close()
    e_ref, f_ref = O.energy_forces(to_oracle(case), case["x"])
    for t in range(10):
        scale = max(abs(e_ref[t]), 1e-12)
        assert abs(e[t] - e_ref[t]) <= E_TOL * scale + 1e-9, (O.TERM_NAMES[t], e[t], e_ref[t])
        assert abs(e[t] - eg[t]) <= E_TOL * scale + 1e-9, (O.TERM_NAMES[t], e[t], eg[t])
    assert force_rel_err(f, f_ref) <= F_TOL
    assert force_rel_err(f, fg) <= F_TOL


def test_hilbert_bit_exact(built_lib):
    from multimm_b200.engine import Engine

    for n in (8, 4096, 100000):
        eng = Engine(n)
        pts = eng.hilbert_points(8)
        assert np.array_equal(pts, O.hilbert_points(n, 8))
        eng.hilbert_init(8, 0.1)
        assert np.array_equal(eng.get_positions(), pts.astype(np.float64) * 0.1)
        eng.close()


def test_determinism_and_newton(built_lib):
    case = make_case(5000, n_chrom=4, seed=9, terms=("EV", "SCB", "CHB"))
    eng = to_engine(case)
    e1, f1 = eng.energy_forces()
    e2, f2 = eng.energy_forces()
    assert np.array_equal(e1, e2) and np.array_equal(f1, f2)
    # pair forces sum to zero (Newton's third law) to FP32 accumulation accuracy
    assert np.abs(f1.sum(axis=0)).max() < 1e-3 * np.abs(f1).max()
    eng.close()


@pytest.mark.parametrize("n,n_chrom,terms", [
    (700, 1, ("EV",)),
    (5000, 4, ("EV", "SCB", "CHB")),
    (12345, 7, ("EV", "COB", "SCB", "CHB", "SC", "LAM", "CF", "BOND", "LOOP", "ANGLE")),
])
def test_newton3_kernel_matches_gather_kernel(built_lib, n, n_chrom, terms):
    """The two exact kernels (Newton-3 with fixed-point accumulation, and gather) are independent
    implementations of the same sum: both must meet the oracle, and agree with each other."""
    case = make_case(n, n_chrom=n_chrom, seed=n, terms=terms)
    eng = to_engine(case)
    e_n3, f_n3 = eng.energy_forces()
    assert eng.pair_kernel_in_use == 2
    e_n3b, f_n3b = eng.energy_forces()
    # bit-reproducible although work items are handed out dynamically (integer accumulation)
    assert np.array_equal(e_n3, e_n3b) and np.array_equal(f_n3, f_n3b)
    eng.set_pair_kernel(1)
    e_g, f_g = eng.energy_forces()
    assert eng.pair_kernel_in_use == 1
    eng.close()
    e_ref, f_ref = O.energy_forces(to_oracle(case), case["x"])
    for e in (e_n3, e_g):
        for t in range(10):
            assert abs(e[t] - e_ref[t]) <= E_TOL * max(abs(e_ref[t]), 1e-12) + 1e-9, (O.TERM_NAMES[t], e[t], e_ref[t])
    assert force_rel_err(f_n3, f_ref) <= F_TOL
    assert force_rel_err(f_g, f_ref) <= F_TOL
    assert force_rel_err(f_n3, f_g) <= F_TOL


@pytest.mark.parametrize("n,n_chrom,rc,terms", [
    (3000, 2, 0.45, ("EV", "SCB", "BOND", "ANGLE")),
    (8000, 3, 0.30, ("EV", "COB", "SCB", "SC", "LAM", "CF", "BOND", "LOOP", "ANGLE")),
    (4000, 4, 0.60, ("EV", "SCB", "CHB", "BOND")),          # CHB is never truncated: exact pass beside the cells
    (500, 1, 5.00, ("EV", "SCB")),                            # cut-off larger than the system: one cell
])
def test_cutoff_mode_cell_list(built_lib, n, n_chrom, rc, terms):
    """Cut-off mode: Morton cell list bit-exact against the oracle's, same number of pairs inside the
    cut-off, energies/forces at the usual bars against the oracle with the same truncation."""
    case = make_case(n, n_chrom=n_chrom, seed=n + 1, terms=terms)
    eng = to_engine(case, cutoff=rc)
    e, f = eng.energy_forces()
    assert eng.pair_kernel_in_use == 3
    sysd = to_oracle(case, cutoff=rc)
    e_ref, f_ref = O.energy_forces(sysd, case["x"])
    for t in range(10):
        assert abs(e[t] - e_ref[t]) <= E_TOL * max(abs(e_ref[t]), 1e-12) + 1e-9, (O.TERM_NAMES[t], e[t], e_ref[t])
    assert force_rel_err(f, f_ref) <= F_TOL
    # integer outputs: bit-exact
    grid = eng.cell_grid()
    assert grid["cell"] == np.float32(rc) / 4 and 1 <= grid["dim"] <= 1024  # key cells: a fixed fraction of the cut-off
    order, keys = eng.cell_list()
    xc = (case["x"] - case["x"].mean(axis=0)).astype(np.float32)
    keys_ref, order_ref = O.cell_list(xc, grid["cell"], grid["dim"], grid["origin"])
    assert np.array_equal(order, order_ref)
    assert np.array_equal(keys, keys_ref)
    assert grid["pairs"] == O.count_pairs(sysd, case["x"])
    # truncation really happened (fewer pairs than all-pairs) unless the cut-off spans the system
    if rc < 1.0:
        assert grid["pairs"] < n * (n - 1) // 2
    e2, f2 = eng.energy_forces()
    assert np.array_equal(e, e2) and np.array_equal(f, f2)
    eng.close()


def test_cutoff_mode_two_kernels_agree(built_lib):
    """Cut-off mode has two independent implementations: the Newton-3 kernel over Morton-sorted tiles
    (default forms) and the gather kernel over a cell list (any form; forced with set_pair_kernel(1)).
    Same truncation, same pair count, energies and forces within the bars of each other and of the oracle."""
    case = make_case(9000, n_chrom=3, seed=31, terms=("EV", "COB", "SCB", "CHB", "SC", "BOND", "LOOP", "ANGLE"))
    rc = 0.4
    eng = to_engine(case, cutoff=rc)
    e1, f1 = eng.energy_forces()
    p1 = eng.cell_grid()["pairs"]
    eng.set_pair_kernel(1)
    e2, f2 = eng.energy_forces()
    p2 = eng.cell_grid()["pairs"]
    eng.set_pair_kernel(2)  # the CTA-level variant of the Newton-3 cut-off pass (default: one warp per item)
    e3, f3 = eng.energy_forces()
    p3 = eng.cell_grid()["pairs"]
    eng.close()
    sysd = to_oracle(case, cutoff=rc)
    e_ref, f_ref = O.energy_forces(sysd, case["x"])
    assert p1 == p2 == p3 == O.count_pairs(sysd, case["x"])
    assert force_rel_err(f1, f3) <= 1e-5  # same pairs, same FP32 pair arithmetic, different FP32 partial sums
    for e in (e1, e2, e3):
        for t in range(10):
            assert abs(e[t] - e_ref[t]) <= E_TOL * max(abs(e_ref[t]), 1e-12) + 1e-9, (O.TERM_NAMES[t], e[t], e_ref[t])
    assert force_rel_err(f1, f_ref) <= F_TOL and force_rel_err(f2, f_ref) <= F_TOL and force_rel_err(f1, f2) <= F_TOL


def test_cutoff_mode_outlier_bead(built_lib):
    """One bead thrown 100 nm out (an L-BFGS trial step can do that): the grid keeps cells of exactly
    the cut-off (hashed Morton keys, up to 1024 cells per axis), so nothing coarsens; keys, order, pair
    count and forces still match the oracle, and the evaluation does not degrade towards O(N^2)."""
    case = make_case(20000, n_chrom=2, seed=41, terms=("EV", "SCB", "BOND", "ANGLE"))
    case["x"] = case["x"].copy()
    case["x"][1234] += np.array([100.0, -60.0, 30.0])
    rc = 0.5
    eng = to_engine(case, cutoff=rc)
    e, f = eng.energy_forces()
    grid = eng.cell_grid()
    assert grid["cell"] == np.float32(rc) / 4 and grid["dim"] > 64
    sysd = to_oracle(case, cutoff=rc)
    e_ref, f_ref = O.energy_forces(sysd, case["x"])
    for t in range(10):
        assert abs(e[t] - e_ref[t]) <= E_TOL * max(abs(e_ref[t]), 1e-12) + 1e-9, (O.TERM_NAMES[t], e[t], e_ref[t])
    assert force_rel_err(f, f_ref) <= F_TOL
    order, keys = eng.cell_list()
    xc = (case["x"] - case["x"].mean(axis=0)).astype(np.float32)
    keys_ref, order_ref = O.cell_list(xc, grid["cell"], grid["dim"], grid["origin"])
    assert np.array_equal(order, order_ref) and np.array_equal(keys, keys_ref)
    assert grid["pairs"] == O.count_pairs(sysd, case["x"])
    # cost: the same system without the outlier takes (almost) the same time per evaluation
    eng.evaluate_timed(3, flush_l2=False)
    t_out = eng.evaluate_timed(10, flush_l2=False)[0]
    case["x"][1234] -= np.array([100.0, -60.0, 30.0])
    eng.set_positions(case["x"])
    eng.evaluate_timed(3, flush_l2=False)
    t_in = eng.evaluate_timed(10, flush_l2=False)[0]
    eng.close()
    assert t_out <= 1.5 * t_in + 0.5, (t_out, t_in)


def test_cutoff_mode_minimizes(built_lib):
    case = make_case(2000, n_chrom=2, seed=21, terms=("EV", "SCB", "SC", "BOND", "LOOP", "ANGLE"))
    eng = to_engine(case, cutoff=0.5)
    rep = eng.minimize(tol=10.0, max_iter=40)
    assert rep["e_final"] < rep["e_initial"] and np.isfinite(rep["e_final"])
    e_chk = O.energy_forces(to_oracle(case, cutoff=0.5), eng.get_positions(), want_forces=False)[0].sum()
    assert abs(e_chk - rep["e_final"]) <= 1e-5 * abs(e_chk)
    eng.close()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_pair_work_is_bit_identical(built_lib, world):
    """Several GPUs, one system: the Newton-3 items are dealt round-robin to the ranks and the
    fixed-point force planes / per-item energies are summed.  Emulated on one GPU (every rank's
    share in turn, same accumulators): forces, energies and a short minimisation must be the SAME
    BITS as the unsharded run — integer sums are exact, each energy slot has one writer."""
    case = make_case(6000, n_chrom=4, seed=77)
    eng = to_engine(case)
    e1, f1 = eng.energy_forces()
    rep1 = eng.minimize(tol=10.0, max_iter=12)
    x1 = eng.get_positions()
    eng.close()
    eng = to_engine(case)
    eng.dist_emulate(world)
    e2, f2 = eng.energy_forces()
    assert eng.pair_kernel_in_use == 2
    rep2 = eng.minimize(tol=10.0, max_iter=12)
    x2 = eng.get_positions()
    eng.close()
    assert np.array_equal(e1, e2) and np.array_equal(f1, f2)
    assert rep1["e_final"] == rep2["e_final"] and rep1["evaluations"] == rep2["evaluations"]
    assert np.array_equal(x1, x2)


@pytest.mark.parametrize("world", [2, 5])
def test_sharded_cutoff_mode_is_bit_identical(built_lib, world):
    """Cut-off mode on several GPUs: the Morton-sorted order is cut into contiguous slabs of i-blocks
    (spatial slabs along the curve), rank r evaluates its slab against itself and the stages above it
    within the cut-off, CHB's exact pass is dealt round-robin; one all-reduce of the fixed-point planes
    combines them.  Emulated on one GPU: same bits as the unsharded cut-off run."""
    case = make_case(7000, n_chrom=3, seed=78)
    eng = to_engine(case, cutoff=0.45)
    e1, f1 = eng.energy_forces()
    p1 = eng.cell_grid()["pairs"]
    rep1 = eng.minimize(tol=10.0, max_iter=10)
    x1 = eng.get_positions()
    eng.close()
    eng = to_engine(case, cutoff=0.45)
    eng.dist_emulate(world)
    e2, f2 = eng.energy_forces()
    assert eng.pair_kernel_in_use == 3
    p2 = eng.cell_grid()["pairs"]
    rep2 = eng.minimize(tol=10.0, max_iter=10)
    x2 = eng.get_positions()
    eng.close()
    assert p1 == p2 and np.array_equal(e1, e2) and np.array_equal(f1, f2)
    assert rep1["e_final"] == rep2["e_final"] and np.array_equal(x1, x2)


def test_full_size_properties(built_lib):
    """BASELINE.json's genome-wide size (N = 2e5, 22 chromosomes, full term set), where the O(N^2) FP64
    oracle takes minutes: size-independent properties instead.  (1) The two independent exact kernels
    agree (Newton-3 / fixed-point vs gather / FP64 partials); (2) pair forces sum to zero (Newton's
    third law) to FP32 accumulation accuracy; (3) energies do not
    change under a rigid translation; (4) the oracle agrees on a 12 000-bead PREFIX of the same
    system (bonded and pair terms restricted to the prefix)."""
    n = 200000
    case = make_case(n, n_chrom=22, seed=2024)
    eng = to_engine(case)
    e_n3, f_n3 = eng.energy_forces()
    assert eng.pair_kernel_in_use == 2
    eng.set_pair_kernel(1)
    e_g, f_g = eng.energy_forces()
    assert np.allclose(e_n3, e_g, rtol=2e-6, atol=1e-6)
    assert force_rel_err(f_n3, f_g) <= F_TOL
    eng.close()
    # (2) pair terms only
    pair_case = make_case(n, n_chrom=22, seed=2024, terms=("EV", "SCB", "CHB"))
    eng = to_engine(pair_case)
    e1, f = eng.energy_forces()
    # the two sides of a pair are accumulated in different FP32 partial sums before the exact
    # fixed-point reduction: the total cancels to FP32 accumulation accuracy
    assert np.abs(f.sum(axis=0)).max() <= 1e-6 * np.abs(f).sum()
    # (3)
    eng.set_positions(pair_case["x"] + np.array([0.7, -1.3, 0.4]))
    e2, _ = eng.energy_forces()
    assert np.allclose(e1, e2, rtol=3e-6)
    eng.close()
    # (4) prefix against the oracle
    m = 12000
    sub = make_case(m, n_chrom=1, seed=1, terms=("EV", "SCB", "CHB"))
    sub["x"], sub["s"], sub["chrom"] = pair_case["x"][:m].copy(), pair_case["s"][:m].copy(), pair_case["chrom"][:m].copy()
    _check(sub)


def test_translation_invariance(built_lib):
    case = make_case(4000, n_chrom=2, seed=10, terms=("EV", "SCB", "CHB", "BOND", "ANGLE", "LOOP"))
    eng = to_engine(case)
    e1, _ = eng.energy_forces()
    eng.set_positions(case["x"] + np.array([3.0, -2.0, 1.0]))
    e2, _ = eng.energy_forces()
    assert np.allclose(e1, e2, rtol=2e-6, atol=1e-6)
    eng.close()


def test_minimize_matches_oracle_energy(built_lib):
    """Final minimised energy vs the CPU L-BFGS at OpenMM's default tolerance.  The north star's
    bar is 1e-3 relative, "trajectories are not compared, because minimization is chaotic" — and
    the chaos reaches the final energy too: the FP64 oracle started from coordinates that differ by
    1e-9 nm ends in basins whose energies differ by several 1e-3 (measured here, every run).  So
    the bar applied is max(1e-3, 2 x the oracle's own spread); what IS held to 1e-5 is that the
    engine's reported final energy equals the oracle's energy at the engine's final positions, and
    that the engine's stopping rule is met."""
    case = make_case(600, n_chrom=2, seed=12, noise=0.0)
    sysd = to_oracle(case)
    tol = 10.0
    eng = to_engine(case)
    rep = eng.minimize(tol=tol, max_iter=0)
    _, rep_ref = O.minimize(sysd, case["x"], tol=tol, max_iter=0)
    assert rep["converged"] == 1 and rep_ref["converged"] == 1, (rep, rep_ref)
    assert rep["e_final"] < rep["e_initial"]
    assert rep["rms_force"] <= tol * 1.0001
    e_chk = O.energy_forces(sysd, eng.get_positions(), want_forces=False)[0].sum()
    assert abs(e_chk - rep["e_final"]) <= 1e-5 * abs(e_chk)
    rng = np.random.default_rng(0)
    spread = 0.0
    for _ in range(3):
        _, rep_p = O.minimize(sysd, case["x"] + rng.normal(0.0, 1e-9, size=case["x"].shape), tol=tol, max_iter=0)
        spread = max(spread, abs(rep_p["e_final"] - rep_ref["e_final"]))
    bar = max(1e-3 * abs(rep_ref["e_final"]), 2.0 * spread)
    assert abs(rep["e_final"] - rep_ref["e_final"]) <= bar, (rep, rep_ref, spread)
    eng.close()


def test_minimize_end_point_is_converged_for_the_oracle(built_lib):
    """Basin-independent form of the final-energy bar: start the FP64 oracle's L-BFGS (same stopping
    rule) FROM the engine's final positions.  If the engine really converged, the oracle has
    (almost) nothing left to do: a handful of iterations at most, and an energy change far below
    1e-3 relative."""
    for n, terms in ((60, ("EV", "SC", "BOND", "ANGLE")), (900, ("EV", "SCB", "CHB", "SC", "LAM", "CF", "BOND", "LOOP", "ANGLE"))):
        case = make_case(n, n_chrom=1 if n < 100 else 3, seed=5, noise=0.0, terms=terms)
        eng = to_engine(case)
        rep = eng.minimize(tol=10.0, max_iter=0)
        assert rep["converged"] == 1, rep
        x1 = eng.get_positions()
        eng.close()
        _, rep2 = O.minimize(to_oracle(case), x1, tol=10.0, max_iter=0)
        assert rep2["converged"] == 1 and rep2["iterations"] <= 5, rep2
        assert abs(rep2["e_initial"] - rep["e_final"]) <= 1e-5 * abs(rep["e_final"])
        assert abs(rep2["e_final"] - rep["e_final"]) <= 1e-3 * abs(rep["e_final"]), (rep, rep2)


def test_errors(built_lib):
    from multimm_b200.engine import Engine, Error

    eng = Engine(100)
    with pytest.raises(Error):
        eng.energy_forces()  # positions never set
    with pytest.raises(ValueError):
        eng.set_pair_term("EV", 7, [1, 2, 3, 4])  # unknown form -> ValueError like model.py:213-215
    with pytest.raises(ValueError):
        eng.set_loops([0], [5], [0.1], [1.0], form=9)
    eng.close()


@pytest.mark.parametrize("cutoff", [0.0, 0.45])
def test_graph_replay_is_bit_identical_to_plain_launches(built_lib, cutoff):
    """mmm_minimize replays a captured CUDA graph per evaluation (per Morton-order period in cut-off
    mode); with the graph switched off it launches kernel by kernel.  Same kernels, same order, same
    arguments: the trajectories are the same bits, and so is the launch count."""
    case = make_case(3000, n_chrom=2, seed=55)
    out = []
    for graph in (True, False):
        eng = to_engine(case, cutoff=cutoff)
        eng.set_graph(graph)
        n0 = eng.launch_count
        rep = eng.minimize(tol=10.0, max_iter=60)
        out.append((rep["e_final"], rep["evaluations"], rep["iterations"], eng.get_positions()))
        eng.close()
    assert out[0][0] == out[1][0] and out[0][1] == out[1][1] and out[0][2] == out[1][2]
    assert np.array_equal(out[0][3], out[1][3])


def test_far_field_on_clusters(built_lib):
    """Coarse-stage far field (cut-off mode, opt-in): CHB and the EV tail beyond the cut-off between
    cluster centroids.  (1) CHB close to the exact same-chromosome sum (a few per cent at most on a compact
    structure), the EV tail positive and small against the truncated sum; (2) a proper potential: its force
    is the gradient of its energy (central differences along a random direction); (3) switching it off
    restores the plain cut-off evaluation bit for bit; (4) exact mode ignores it."""
    case = make_case(6000, n_chrom=3, seed=91, terms=("EV", "CHB"), chb_de=5.0)
    eng = to_engine(case, cutoff=0.4)

    def both(x):
        """(plain cut-off evaluation, evaluation with the far field on clusters) at x"""
        eng.set_positions(x)
        eng.set_chb_surrogate(False)
        plain = eng.energy_forces()
        eng.set_chb_surrogate(True)
        far = eng.energy_forces()
        return plain, far

    (e_exact, f_exact), (e_s, f_s) = both(case["x"])
    assert abs(e_s[3] - e_exact[3]) <= 0.05 * abs(e_exact[3]) and e_s[3] != e_exact[3]
    tail = e_s[0] - e_exact[0]
    e_full = O.energy_forces(to_oracle(case), case["x"], want_forces=False)[0]
    # the tail the truncation drops, recovered to a few tens of per cent by the cluster sum
    assert tail > 0 and abs(tail - (e_full[0] - e_exact[0])) <= 0.5 * (e_full[0] - e_exact[0]), (tail, e_full[0] - e_exact[0])
    rng = np.random.default_rng(3)
    d = rng.normal(size=case["x"].shape)
    d /= np.linalg.norm(d)
    h = 1e-4

    def far_energy(x):
        (ep, _), (es, _) = both(x)
        return (es[0] - ep[0]) + es[3]

    slope = -(far_energy(case["x"] + h * d) - far_energy(case["x"] - h * d)) / (2 * h)
    f_far = f_s - (f_exact - _chb_only_forces(case))  # far-field force: drop the (identical) truncated EV part
    assert abs(slope - float((f_far * d).sum())) <= 2e-4 * max(abs(slope), 1.0), (slope, float((f_far * d).sum()))
    eng.set_positions(case["x"])
    eng.set_chb_surrogate(False)
    e_back, f_back = eng.energy_forces()
    assert np.array_equal(e_back, e_exact) and np.array_equal(f_back, f_exact)
    eng.set_cutoff(0.0)
    eng.set_chb_surrogate(True)
    e_nocut, _ = eng.energy_forces()
    eng.close()
    assert abs(e_nocut[3] - e_full[3]) <= E_TOL * abs(e_full[3])  # exact mode: the reference's potential, always
    assert abs(e_nocut[0] - e_full[0]) <= E_TOL * abs(e_full[0])


def _chb_only_forces(case):
    """Exact CHB forces of the case from the oracle (FP64)."""
    only = dict(case)
    only["ev"] = None
    _, f = O.energy_forces(to_oracle(only), case["x"])
    return f
