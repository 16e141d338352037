#!/bin/bash
# round 2, call 19 (1 GPU): 16 ensemble members with the FP32 cluster far field (statistics of the coarse / exact-probe rounds)
mkdir -p gpurun_out
timeout 600 python scripts/gpu_ensemble.py 16 0 0.5 > gpurun_out/r2c19_ensemble_16x1.log 2>&1; echo "ens rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/ensemble_16x1gpu_0.5.json")); print(d["structures_per_hour"], d["wall_seconds"], [(r["iterations"], round(r["minimize_s"],2), r.get("coarse_iterations"), r.get("coarse_rounds")) for r in d["per_replica"]])
PY
