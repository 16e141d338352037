"""Short GPU command for ncu: the coarse stage of the two-stage minimisation on the bench's GW model (pair terms
truncated at rc, far field on cluster centroids), a few L-BFGS iterations with plain launches."""
import sys
import tempfile

sys.path.insert(0, ".")
import bench  # noqa: E402

rc = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
with tempfile.TemporaryDirectory() as tmp:
    m = bench.build_model("gw", seed=0, device=0, tmp=tmp)
    eng = m.engine
    eng.set_cutoff(rc)
    eng.set_chb_surrogate(True)
    eng.set_graph(False)
    rep = eng.minimize(tol=10.0, max_iter=int(sys.argv[2]) if len(sys.argv) > 2 else 20)
    print(rep)
    m.close()
