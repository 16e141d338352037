#!/bin/bash
# round 2, call 28 (1 GPU): the three-stage ensemble pipeline on real members: 8 genome-wide members on one GPU; driver tests
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_driver.py -m gpu -q -x > gpurun_out/r2c28_pytest.log 2>&1; echo "pytest rc=$?"
timeout 110 python scripts/gpu_ensemble.py 8 0 0.5 > gpurun_out/r2c28_ensemble_8x1.log 2>&1; echo "ens rc=$?"
tail -n 3 gpurun_out/r2c28_pytest.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/ensemble_8x1gpu_0.5.json")); pr=d["per_replica"]
print(d["structures_per_hour"], d["wall_seconds"], [(r["replica"], r["iterations"], round(r["minimize_s"],2), round(r["seconds"],2), r["converged"]) for r in pr])
PY
tail -n 5 gpurun_out/r2c28_ensemble_8x1.log | cut -c1-300
