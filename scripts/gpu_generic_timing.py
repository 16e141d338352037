"""Non-default functional forms: ms per evaluation on the gather kernel (every pair from both sides) and on the
Newton-3 work items with the generic FP64 body.  usage: python scripts/gpu_generic_timing.py [n_beads]"""
import json
import sys

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
from common import make_case, to_engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
out = {"n_beads": n}
for name, forms, p in (("yukawa_blocks+gaussian_chb", {"COB": 1, "SCB": 1, "CHB": 1}, 6.0), ("ev_power_4.5", {}, 4.5)):
    case = make_case(n, n_chrom=5, seed=3, forms=forms, chb_de=1.0, ev_power=p, terms=("EV", "COB", "SCB", "CHB"))
    eng = to_engine(case)
    res = {}
    for label, pref in (("newton3_generic", 0), ("gather_generic", 1)):
        eng.set_pair_kernel(pref)
        eng.energy_forces()
        tot, pair = eng.evaluate_timed(5, flush_l2=False)
        res[label] = dict(kernel=eng.pair_kernel_in_use, ms_per_eval=tot / 5, pair_ms=pair / 5)
    eng.close()
    out[name] = res
print(json.dumps(out))
