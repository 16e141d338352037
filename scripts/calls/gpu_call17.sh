#!/bin/bash
# round 2, call 17 (1 GPU): the default bench line with the final code (what the driver runs), its reference arm,
# and 16 ensemble members on one GPU (the 1-GPU leg of the structures/hour scaling)
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2c17_bench.json 2> gpurun_out/r2c17_bench.err; echo "bench rc=$?"
timeout 600 python scripts/gpu_ensemble.py 16 0 0.5 > gpurun_out/r2c17_ensemble_16x1.log 2>&1; echo "ens rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2c17_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches","clocks")}, d["roofline"]["frac_nominal"], d["roofline"]["traffic"], d["minimize_full"]["exact"]["wall_seconds"], d["minimize_full"].get("two_stage",{}).get("minimize_s"), d["ensemble"]["structures_per_hour"], d["cpu_baseline"]["value"])
d=json.load(open("gpurun_out/ensemble_16x1gpu_0.5.json")); print(d["structures_per_hour"], d["wall_seconds"], max(r["iterations"] for r in d["per_replica"]))
PY
