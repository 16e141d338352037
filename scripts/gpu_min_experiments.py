"""GPU experiments on minimisation wall time: (1) L-BFGS history length (set MMM_LIB_NAME to a
variant library), (2) two-stage minimisation: cut-off forces first, exact forces to finish."""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, ".")
import bench  # noqa: E402

which = sys.argv[1].split(",")
stage_rc = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
tag = os.environ.get("MMM_LIB_NAME", "default")
out = {}
for wl in which:
    with tempfile.TemporaryDirectory() as tmp:
        m = bench.build_model(wl, seed=0, device=0, tmp=tmp)
        eng = m.engine
        t0 = time.time()
        reps = []
        if stage_rc > 0:
            eng.set_cutoff(stage_rc)
            reps.append(eng.minimize(10.0, 0))
            eng.set_cutoff(0.0)
        reps.append(eng.minimize(10.0, 0))
        wall = time.time() - t0
        out[wl] = dict(lib=tag, stage_rc=stage_rc, wall=wall, stages=reps)
        print(wl, tag, stage_rc, "wall %.2f" % wall, [(r["iterations"], r["evaluations"], round(r["e_final"], 1), r["converged"], round(r["wall_seconds"], 2)) for r in reps], flush=True)
        m.close()
json.dump(out, open(f"gpurun_out/minexp_{tag}_{stage_rc}.json", "w"), indent=1)
