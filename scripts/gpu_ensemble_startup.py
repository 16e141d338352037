"""Start-up cost of the ensemble workers: two workers on GPU 0, four small members, with and without the
early CUDA start (MMM_NO_EARLY_CUDA=1).  usage: python scripts/gpu_ensemble_startup.py"""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, ".")
from multimm_b200 import run, synthetic  # noqa: E402
from multimm_b200.config import SimulationConfig  # noqa: E402


def main():
    tmp = tempfile.mkdtemp(prefix="mmm_ens_")
    bedpe, bed = os.path.join(tmp, "loops.bedpe"), os.path.join(tmp, "comps.bed")
    synthetic.write_bedpe(bedpe, n_loops=2000, seed=100)
    synthetic.write_bed(bed, seed=100)
    out = {}
    for label, off in (("early_cuda_start", "0"), ("no_early_start", "1"), ("early_cuda_start_again", "0")):
        os.environ["MMM_NO_EARLY_CUDA"] = off
        args = SimulationConfig(PLATFORM="B200", N_BEADS=20000, LOOPS_PATH=bedpe, COMPARTMENT_PATH=bed,
                                OUT_PATH=os.path.join(tmp, "out_" + label), SAVE_PLOTS=False, SHUFFLE_CHROMS=True,
                                SCB_USE_SUBCOMPARTMENT_BLOCKS=True, GENERATE_ENSEMBLE=True, N_ENSEMBLE=4,
                                MIN_MAX_ITERATIONS=300)
        t0 = time.time()
        reports = run.run_ensemble(args, devices=[0, 0])
        out[label] = dict(wall_seconds=time.time() - t0, first_initialize_s=sorted(r["initialize_s"] for r in reports)[-2:],
                          e_final=[r["e_final"] for r in reports], devices=[r["device"] for r in reports])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
