"""Small end-to-end run for compute-sanitizer: every kernel family once (exact N3, gather, cut-off
cells + CHB-only, bonded, L-BFGS, MD, Hilbert) on a 3 000-bead system with ragged padding."""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from common import make_case, to_engine  # noqa: E402

case = make_case(2999, n_chrom=3, seed=42)
eng = to_engine(case)
e, f = eng.energy_forces()
eng.set_pair_kernel(1)
eng.energy_forces()
eng.set_pair_kernel(0)
eng.minimize(10.0, 4)
eng.md_configure("langevin", 0.001, 310.0, 0.5, 16427.889, 1)
eng.set_velocities_to_temperature(310.0, 1)
eng.md_run(2)
eng.set_cutoff(0.4)
eng.energy_forces()
eng.minimize(10.0, 2)
eng.set_cutoff(0.0)
eng.dist_emulate(3)
eng.energy_forces()
eng.hilbert_points(8)
eng.close()
print("ok", float(e.sum()))
