#!/bin/bash
# round 2, call 26 (8 GPUs): configs[3] with the final code — the genome-wide ensemble of 64 structures over 8 GPUs
mkdir -p gpurun_out
echo "start $(date +%s)" > gpurun_out/r2c26_times.txt
timeout 60 python -c "
from multimm_b200.engine import Engine
e = Engine(1000); e.close(); print('warm')" >> gpurun_out/r2c26_times.txt 2>&1
echo "warm $(date +%s)" >> gpurun_out/r2c26_times.txt
timeout 300 python scripts/gpu_ensemble.py 64 0,1,2,3,4,5,6,7 0.5 > gpurun_out/r2c26_ensemble_64x8.log 2>&1; echo "ens64 rc=$? $(date +%s)" >> gpurun_out/r2c26_times.txt
cat gpurun_out/r2c26_times.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/ensemble_64x8gpu_0.5.json")); pr=d["per_replica"]
print(d["structures_per_hour"], d["wall_seconds"], "rounds>1:", sum(1 for r in pr if (r.get("coarse_rounds") or 1)>1), "max exact it", max(r["iterations"] for r in pr), "init first", [round(r["initialize_s"],1) for r in pr[:8]], "converged", sum(r["converged"] for r in pr))
PY
