"""GPU parity at BASELINE.json's own sizes, against the FULL oracle (no prefix, every term):

  S1  configs[0]  specific region,   N = 1e4, {EV, bonds, loops, angles}
  S2  configs[1]  single chromosome, N = 5e4, + SCB
  S3  configs[2]  genome-wide,       N = 2e5, 22 chromosomes, full config_gw.ini term set

built exactly as bench.py builds them (synthetic .bedpe/.bed -> loaders -> MultiMM.add_* -> C-ABI).
Bars (north star): all ten per-term energies 1e-5 relative, per-bead forces 1e-4 relative.  The
force bar is applied in two forms: the suite's usual one (error relative to max(|F_i|, RMS |F|)) and
the STRICT per-bead one (error relative to |F_i| itself) over the beads with |F_i| > 1e-3 RMS.
The oracle is O(N^2) FP64 on the host: ~2 s at S2 and ~30 s at S3 on the GPU box's cores (its
thread count is set explicitly — torchrun-style environments export OMP_NUM_THREADS=1).
"""
import os
import sys
import tempfile

import numpy as np
import pytest

from common import O, force_rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

E_TOL = 1e-5
F_TOL = 1e-4
NTHREADS = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def strict_force_err(f, f_ref, floor=1e-3):
    """max, 99.9th and 99th percentile of |F_i - F_ref_i| / |F_ref_i| over beads with |F_ref_i| > floor * RMS."""
    mag = np.linalg.norm(f_ref, axis=1)
    rms = float(np.sqrt((mag ** 2).mean()))
    sel = mag > floor * rms
    err = np.linalg.norm(f - f_ref, axis=1)[sel] / mag[sel]
    return float(err.max()), float(np.quantile(err, 0.999)), float(np.quantile(err, 0.99)), int(sel.sum())


def _compare(m, x, label, record):
    import bench

    eng = m.engine
    eng.set_positions(x)
    e, f = eng.energy_forces()
    assert eng.pair_kernel_in_use == 2  # the default Newton-3 kernel is what is being checked
    sysd, _ = bench.oracle_system(m, m.args.N_BEADS)
    e_ref, f_ref = O.energy_forces(sysd, x, nthreads=NTHREADS)
    worst = 0.0
    for t in range(10):
        scale = max(abs(e_ref[t]), 1e-12)
        worst = max(worst, abs(e[t] - e_ref[t]) / scale if abs(e_ref[t]) > 1e-9 else 0.0)
        assert abs(e[t] - e_ref[t]) <= E_TOL * scale + 1e-9, (label, O.TERM_NAMES[t], e[t], e_ref[t])
    soft = force_rel_err(f, f_ref)
    smax, s999, s99, nsel = strict_force_err(f, f_ref)
    record(f"{label}: worst term energy rel err {worst:.2e}; force err vs max(|F_i|, RMS) {soft:.2e}; "
           f"strict per-bead over {nsel} beads with |F| > 1e-3 RMS: 99 % {s99:.2e}, 99.9 % {s999:.2e}, max {smax:.2e}")
    assert soft <= F_TOL, (label, soft)
    # strict per-bead (error relative to the bead's own |F|): 99 % of the beads inside the bar, 99.9 %
    # within 3x, the worst bead within 10x.  Measured on B200 (profiles/r02_parity_full_sizes.md): the
    # tail grows with the coordinate magnitude (FP32 deltas of centred coordinates: 2 nm at S1, 6 nm at
    # S3) on beads whose net force is the small difference of large near-neighbour pair forces.
    assert s99 <= F_TOL, (label, s99)
    assert s999 <= 3 * F_TOL, (label, s999)
    assert smax <= 10 * F_TOL, (label, smax)
    return e, f


@pytest.fixture(scope="module")
def record():
    lines = []
    yield lines.append
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_full_sizes.log"), "a") as fh:
            fh.write("\n".join(lines) + "\n")
    except OSError:
        pass
    print("\n" + "\n".join(lines))


@pytest.mark.parametrize("workload", ["region", "chrom", "gw"])
def test_baseline_config_matches_full_oracle(built_lib, workload, record):
    """Hilbert start (lattice: many exactly equal distances, forces cancelling by symmetry) and the
    structure after 40 L-BFGS iterations (generic positions, coordinates spread over the nucleus)."""
    import bench

    with tempfile.TemporaryDirectory(prefix="mmm_parity_") as tmp:
        m = bench.build_model(workload, seed=0, device=0, tmp=tmp)
        x0 = m.positions.copy()
        _compare(m, x0, f"{workload} N={m.args.N_BEADS} Hilbert start", record)
        m.engine.set_positions(x0)
        m.engine.minimize(tol=10.0, max_iter=40)
        x1 = m.engine.get_positions()
        _compare(m, x1, f"{workload} N={m.args.N_BEADS} after 40 L-BFGS iterations", record)
        m.close()


def test_s1_minimised_energy_matches_oracle_lbfgs(built_lib, record):
    """configs[0] (N = 1e4): final minimised energy of the on-device L-BFGS against the oracle's
    liblbfgs restatement started from the same structure, OpenMM's default tolerance and unlimited
    iterations.  North-star bar: 1e-3 relative.  Also the basin-independent form: the oracle's
    L-BFGS started FROM the engine's end point has nothing left to do."""
    import bench

    with tempfile.TemporaryDirectory(prefix="mmm_parity_") as tmp:
        m = bench.build_model("region", seed=0, device=0, tmp=tmp)
        x0 = m.positions.copy()
        rep = m.engine.minimize(tol=10.0, max_iter=0)
        assert rep["converged"] == 1, rep
        x1 = m.engine.get_positions()
        sysd, _ = bench.oracle_system(m, m.args.N_BEADS)
        m.close()
    e_chk = O.energy_forces(sysd, x1, want_forces=False, nthreads=NTHREADS)[0].sum()
    assert abs(e_chk - rep["e_final"]) <= 1e-5 * abs(e_chk)
    _, rep_end = O.minimize(sysd, x1, tol=10.0, max_iter=0, nthreads=NTHREADS)
    # the two sides disagree by ~1e-3 on |g| at the end point (FP32 pair forces), so the oracle may sit
    # just above the threshold the engine just crossed: a few more iterations, a tiny energy change
    record(f"S1 end point: oracle L-BFGS started from the engine's final positions: {rep_end['iterations']} iterations, "
           f"{rep_end['e_initial']:.4f} -> {rep_end['e_final']:.4f} kJ/mol")
    assert rep_end["converged"] == 1 and rep_end["iterations"] <= 50, rep_end
    assert abs(rep_end["e_final"] - rep["e_final"]) <= 1e-3 * abs(rep["e_final"])
    # The final energy itself: minimisation from the Hilbert lattice is chaotic (the lattice is full of
    # exactly cancelling forces, rounding decides how the symmetry breaks), and the oracle is not even
    # reproducible against itself — its OpenMP partial sums depend on the dynamic schedule.  So the
    # oracle's L-BFGS is run three times from the SAME start; the engine must land within the oracle's
    # own range widened by twice its width (at least 1e-3 relative, the north star's bar).
    finals = []
    for _ in range(3):
        _, rep_ref = O.minimize(sysd, x0, tol=10.0, max_iter=0, nthreads=NTHREADS)
        assert rep_ref["converged"] == 1, rep_ref
        finals.append(rep_ref["e_final"])
    lo, hi = min(finals), max(finals)
    width = max(hi - lo, 1e-3 * abs(hi))
    rel = abs(rep["e_final"] - finals[0]) / abs(finals[0])
    record(f"S1 minimisation: engine {rep['e_final']:.3f} kJ/mol in {rep['iterations']} iterations; oracle L-BFGS, three runs "
           f"from the same start: {', '.join(f'{v:.3f}' for v in finals)} (own spread {(hi - lo) / abs(hi):.2e} relative); "
           f"engine vs first oracle run {rel:.2e}")
    assert lo - 2.0 * width <= rep["e_final"] <= hi + 2.0 * width, (rep, finals)
