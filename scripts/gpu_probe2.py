"""GPU probe: pair-kernel time of the Newton-3 kernel vs the gather kernel at the BASELINE sizes."""
import json
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from common import make_case, to_engine  # noqa: E402

out = {}
for n, n_chrom, terms in [
    (10000, 1, ("EV", "BOND", "LOOP", "ANGLE")),
    (50000, 1, ("EV", "SCB", "BOND", "LOOP", "ANGLE")),
    (200000, 22, ("EV", "BOND", "LOOP", "ANGLE")),
    (200000, 22, ("EV", "SCB", "CHB", "SC", "LAM", "CF", "BOND", "LOOP", "ANGLE")),
]:
    case = make_case(n, n_chrom=n_chrom, seed=1, terms=terms)
    eng = to_engine(case)
    key = f"n={n} terms={'+'.join(terms)}"
    out[key] = {}
    for which, name in ((0, "n3"), (1, "gather")):
        eng.set_pair_kernel(which)
        eng.evaluate_n(2)
        tot, pair = eng.evaluate_timed(5, flush_l2=False)
        e, f = eng.energy_forces()
        out[key][name] = {"pair_ms": pair / 5, "total_ms": tot / 5, "kernel": eng.pair_kernel_in_use,
                          "gpairs_per_s": n * (n - 1) / 2 / (pair / 5 * 1e-3) / 1e9,
                          "E": [float(v) for v in e[:4]], "fsum": [float(v) for v in f.sum(axis=0)]}
        print(key, name, out[key][name], flush=True)
    eng.close()
json.dump(out, open("gpurun_out/probe2.json", "w"), indent=1)
