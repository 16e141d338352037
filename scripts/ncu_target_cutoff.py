"""Short GPU command for ncu: the genome-wide model in cut-off mode (rc = 0.5 nm), a few evaluations."""
import sys
import tempfile

sys.path.insert(0, ".")
import bench  # noqa: E402

with tempfile.TemporaryDirectory() as tmp:
    m = bench.build_model("gw", seed=0, device=0, tmp=tmp)
    m.engine.set_cutoff(0.5)
    m.engine.evaluate_n(4)
    print("cells pass ms", m.engine.last_pair_kernel_ms, m.engine.cell_grid())
    m.close()
