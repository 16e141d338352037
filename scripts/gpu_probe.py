"""Quick GPU probe: FP32/MUFU peaks and force-evaluation rates at the BASELINE sizes."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from common import make_case, to_engine  # noqa: E402
from multimm_b200.engine import measure_fp32_peak  # noqa: E402

out = {"fp32_tflops,mufu_tops": measure_fp32_peak(0)}
print(out, flush=True)
for n, n_chrom, terms in [
    (10000, 1, ("EV", "BOND", "LOOP", "ANGLE")),
    (50000, 1, ("EV", "SCB", "BOND", "LOOP", "ANGLE")),
    (200000, 22, ("EV", "BOND", "LOOP", "ANGLE")),
    (200000, 22, ("EV", "SCB", "CHB", "SC", "LAM", "CF", "BOND", "LOOP", "ANGLE")),
]:
    case = make_case(n, n_chrom=n_chrom, seed=1, terms=terms)
    eng = to_engine(case)
    eng.evaluate_n(3)
    t0 = time.time()
    k = 10
    eng.evaluate_n(k)
    dt = (time.time() - t0) / k
    key = f"n={n} terms={'+'.join(terms)}"
    out[key] = {"ms_per_eval": dt * 1e3, "pair_ms": eng.last_pair_kernel_ms,
                "gpairs_per_s": n * (n - 1) / 2 / dt / 1e9}
    print(key, out[key], flush=True)
    if n == 200000 and "SCB" in terms:
        t0 = time.time()
        rep = eng.minimize(10.0, 200)
        out["minimize_200_iters"] = rep
        print(rep, time.time() - t0, flush=True)
    eng.close()
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1, default=str)
