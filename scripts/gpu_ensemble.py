"""GPU run: genome-wide ensemble through the driver (run.run_ensemble): structures/hour.
usage: gpu_ensemble.py <n_ensemble> <gpus comma list> [coarse_cutoff_nm]

Everything below the imports sits under the __main__ guard: run_ensemble starts its workers with the
"spawn" method, which re-imports this file in every worker; an unguarded body would run there too
and the workers would die while bootstrapping (that is what left the first 8-GPU attempt waiting
for reports that never came, see profiles/r01_ensemble.md)."""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, ".")
from multimm_b200 import run, synthetic  # noqa: E402
from multimm_b200.config import SimulationConfig  # noqa: E402


def main():
    n_ens = int(sys.argv[1])
    devices = [int(t) for t in sys.argv[2].split(",")]
    coarse = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    tmp = tempfile.mkdtemp(prefix="mmm_ens_")
    bedpe, bed = os.path.join(tmp, "loops.bedpe"), os.path.join(tmp, "comps.bed")
    synthetic.write_bedpe(bedpe, n_loops=10000, seed=100)
    synthetic.write_bed(bed, seed=100)
    args = SimulationConfig(PLATFORM="B200", N_BEADS=200000, LOOPS_PATH=bedpe, COMPARTMENT_PATH=bed,
                            OUT_PATH=os.path.join(tmp, "out"), SAVE_PLOTS=False, SHUFFLE_CHROMS=True,
                            SC_USE_SPHERICAL_CONTAINER=True, CHB_USE_CHROMOSOMAL_BLOCKS=True,
                            SCB_USE_SUBCOMPARTMENT_BLOCKS=True, IBL_USE_B_LAMINA_INTERACTION=True,
                            CF_USE_CENTRAL_FORCE=True, GENERATE_ENSEMBLE=True, N_ENSEMBLE=n_ens, MIN_COARSE_CUTOFF=coarse)
    t0 = time.time()
    reports = run.run_ensemble(args, devices=devices)
    dt = time.time() - t0
    out = dict(n_ensemble=n_ens, gpus=len(devices), coarse_cutoff=coarse, wall_seconds=dt,
               structures_per_hour=3600.0 * n_ens / dt,
               per_replica=[{k: r.get(k) for k in ("replica", "device", "seconds", "iterations", "evaluations", "e_final",
                                                   "converged", "minimize_s", "initialize_s", "forcefield_s", "write_cif_s",
                                                   "coarse_iterations", "coarse_rounds", "exact_iterations")}
                            for r in reports])
    print(json.dumps(out))
    json.dump(out, open(f"gpurun_out/ensemble_{n_ens}x{len(devices)}gpu_{coarse}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
