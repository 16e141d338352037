"""A/B of cut-off pass variants (library chosen with MMM_LIB_NAME): ms per cut-off pass on the genome-wide
model at the Hilbert start and on a disordered structure, and bits of the result for comparison.
usage: MMM_LIB_NAME=libmultimm_b200_x.so python scripts/gpu_cutoff_ab.py [rc_nm]"""
import hashlib
import json
import os
import sys
import tempfile

sys.path.insert(0, ".")
import bench  # noqa: E402

rc = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
out = {"lib": os.environ.get("MMM_LIB_NAME", "libmultimm_b200.so"), "rc_nm": rc}
with tempfile.TemporaryDirectory() as tmp:
    m = bench.build_model("gw", seed=0, device=0, tmp=tmp)
    eng = m.engine
    x0 = m.positions.copy()
    eng.set_cutoff(rc)
    eng.set_pair_kernel(0)
    eng.set_chb_surrogate(True)
    eng.set_positions(x0)
    eng.minimize(tol=10.0, max_iter=200)
    x1 = eng.get_positions()
    for name, x in (("start", x0), ("relaxed", x1)):
        eng.set_positions(x)
        e, f = eng.energy_forces()
        eng.evaluate_timed(5, flush_l2=False)
        best = None
        for _ in range(3):
            tot, pair = eng.evaluate_timed(40, flush_l2=False)
            best = (tot / 40, pair / 40) if best is None or pair / 40 < best[1] else best
        out[name] = dict(ms_per_eval=best[0], cut_pass_ms=best[1], pairs_in_cutoff=eng.cell_grid()["pairs"],
                         energy=float(e.sum()), forces_sha=hashlib.sha1(f.tobytes()).hexdigest()[:12])
    m.close()
print(json.dumps(out))
