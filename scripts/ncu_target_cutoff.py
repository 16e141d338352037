"""Short GPU command for ncu: the bench's GW model in cut-off mode (rc = 0.5 nm), a few evaluations."""
import sys
import tempfile

sys.path.insert(0, ".")
import bench  # noqa: E402

rc = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
with tempfile.TemporaryDirectory() as tmp:
    m = bench.build_model("gw", seed=0, device=0, tmp=tmp)
    m.engine.set_cutoff(rc)
    m.engine.evaluate_n(10)
    print("cut pass ms", m.engine.last_pair_kernel_ms, m.engine.cell_grid())
    m.close()
