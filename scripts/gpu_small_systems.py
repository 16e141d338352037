"""Small systems (configs[0], S1 = 1e4 beads; S2 = 5e4): minimisation wall time and per-evaluation time with
the CUDA-graph replay on and off.  usage: python scripts/gpu_small_systems.py"""
import json
import sys
import tempfile

sys.path.insert(0, ".")
import bench  # noqa: E402

out = {}
with tempfile.TemporaryDirectory() as tmp:
    for wl in ("region", "chrom"):
        m = bench.build_model(wl, seed=0, device=0, tmp=tmp)
        x0 = m.positions.copy()
        for graph in (True, False, True):
            m.engine.set_graph(graph)
            m.engine.set_positions(x0)
            rep = m.engine.minimize(tol=10.0, max_iter=0)
            out[f"{wl}_graph_{int(graph)}"] = dict(wall_s=rep["wall_seconds"], evaluations=rep["evaluations"],
                                                   iterations=rep["iterations"], us_per_evaluation=1e6 * rep["wall_seconds"] / rep["evaluations"],
                                                   e_final=rep["e_final"], converged=rep["converged"])
        m.close()
print(json.dumps(out))
