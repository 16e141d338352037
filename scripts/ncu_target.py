"""Short GPU command for ncu: builds the bench's GW model and runs a few evaluations (+ a few
L-BFGS iterations with --min)."""
import sys
import tempfile

sys.path.insert(0, ".")
import bench  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "gw"
with tempfile.TemporaryDirectory() as tmp:
    m = bench.build_model(wl, seed=0, device=0, tmp=tmp)
    m.engine.evaluate_n(3)
    if "--min" in sys.argv:
        print(m.engine.minimize(10.0, 5))
    print("pair ms", m.engine.last_pair_kernel_ms)
    m.close()
