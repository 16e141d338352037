"""Cut-off mode on the genome-wide model: ms per evaluation of the two implementations (Newton-3 kernel
over Morton-sorted tiles; cell-list gather kernel), pairs inside the cut-off, and the two-stage
minimisation wall time.  usage: python scripts/gpu_cutoff_timing.py [rc_nm] [workload]"""
import json
import sys
import tempfile
import time

sys.path.insert(0, ".")
import bench  # noqa: E402

rc = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
wl = sys.argv[2] if len(sys.argv) > 2 else "gw"
out = {"rc_nm": rc, "workload": wl}
with tempfile.TemporaryDirectory() as tmp:
    m = bench.build_model(wl, seed=0, device=0, tmp=tmp)
    eng = m.engine
    x0 = m.positions.copy()
    eng.set_cutoff(rc)
    for name, pref in (("n3_warp_per_item", 0), ("n3_cta_level", 2), ("cells_gather", 1)):
        eng.set_pair_kernel(pref)
        eng.set_positions(x0)
        e, _ = eng.energy_forces()
        eng.evaluate_timed(5, flush_l2=False)
        tot, pair = eng.evaluate_timed(40, flush_l2=False)
        out[name] = dict(ms_per_eval=tot / 40, cut_pass_ms=pair / 40, pairs_in_cutoff=eng.cell_grid()["pairs"],
                         grid=eng.cell_grid(), energy=float(e.sum()))
    print(json.dumps(out), flush=True)
    # after 200 L-BFGS iterations on the truncated potential the structure is disordered: time again
    eng.set_pair_kernel(0)
    eng.set_positions(x0)
    eng.minimize(tol=10.0, max_iter=200)
    x1 = eng.get_positions()
    for name, pref in (("n3_warp_per_item_relaxed", 0), ("n3_cta_level_relaxed", 2), ("cells_gather_relaxed", 1)):
        eng.set_pair_kernel(pref)
        eng.set_positions(x1)
        eng.evaluate_timed(5, flush_l2=False)
        tot, pair = eng.evaluate_timed(40, flush_l2=False)
        out[name] = dict(ms_per_eval=tot / 40, cut_pass_ms=pair / 40, pairs_in_cutoff=eng.cell_grid()["pairs"])
    # the same with CHB on cluster centroids (coarse-stage surrogate)
    eng.set_pair_kernel(0)
    eng.set_chb_surrogate(True)
    eng.set_positions(x1)
    eng.evaluate_timed(5, flush_l2=False)
    tot, pair = eng.evaluate_timed(40, flush_l2=False)
    out["n3_warp_per_item_relaxed_chb_clusters"] = dict(ms_per_eval=tot / 40, cut_pass_ms=pair / 40)
    # two-stage minimisation (MIN_COARSE_CUTOFF): coarse stage on the truncated potential, exact stage after
    for name, surrogate in (("two_stage", False), ("two_stage_far_field_clusters", True)):
        eng.set_cutoff(rc)
        eng.set_chb_surrogate(surrogate)
        eng.set_positions(x0)
        t0 = time.perf_counter()
        rep_c = eng.minimize(tol=10.0, max_iter=20000)
        eng.set_cutoff(0.0)
        eng.set_chb_surrogate(False)
        rep_e = eng.minimize(tol=10.0, max_iter=0)
        out[name] = dict(total_s=time.perf_counter() - t0, coarse=rep_c, exact=rep_e)
        print(json.dumps({name: out[name]}), flush=True)
    m.close()
print(json.dumps(out))
