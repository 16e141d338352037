"""GPU run: minimise the BASELINE workloads to OpenMM's default tolerance (10 kJ/mol/nm, unlimited
iterations unless --cap) and record iterations / evaluations / wall time; also time cut-off mode."""
import json
import sys
import tempfile
import time

sys.path.insert(0, ".")
import bench  # noqa: E402

cap = int(sys.argv[1]) if len(sys.argv) > 1 else 0
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["region", "chrom", "gw"]
out = {}
for wl in which:
    with tempfile.TemporaryDirectory() as tmp:
        m = bench.build_model(wl, seed=0, device=0, tmp=tmp)
        eng = m.engine
        eng.evaluate_n(2)
        tot, pair = eng.evaluate_timed(5, flush_l2=True)
        t0 = time.time()
        rep = eng.minimize(10.0, cap)
        rep["host_wall_seconds"] = time.time() - t0
        rep["ms_per_eval_standalone"] = tot / 5
        rep["pair_kernel"] = eng.pair_kernel_in_use
        e, _ = eng.energy_forces()
        rep["e_terms_final"] = [float(v) for v in e]
        out[wl] = rep
        print(wl, rep, flush=True)
        if wl == "gw":
            for rc in (0.5, 1.0):
                m.engine.set_positions(m.positions)
                m.engine.set_cutoff(rc)
                m.engine.evaluate_n(2)
                tot, pair = m.engine.evaluate_timed(10, flush_l2=True)
                g = m.engine.cell_grid()
                out[f"gw_cutoff_{rc}"] = dict(ms_per_eval=tot / 10, pair_ms=pair / 10, **g)
                print("cutoff", rc, out[f"gw_cutoff_{rc}"], flush=True)
        m.close()
json.dump(out, open("gpurun_out/minimize_r1.json", "w"), indent=1)
