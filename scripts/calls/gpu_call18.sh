#!/bin/bash
# round 2, call 18 (1 GPU): far field on clusters with FP32 pair arithmetic — its tests, the coarse evaluation time,
# the two-stage minimisation, 8 ensemble members
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_driver.py -m gpu -q -x -k "far_field or cutoff or two_stage or coarse or ensemble" > gpurun_out/r2c18_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c18_pytest.log
timeout 240 python scripts/gpu_cutoff_ab.py 0.5 > gpurun_out/r2c18_cutoff_ab.jsonl 2> gpurun_out/r2c18_cutoff_ab.err
timeout 600 python scripts/gpu_ensemble.py 8 0 0.5 > gpurun_out/r2c18_ensemble_8x1.log 2>&1; echo "ens rc=$?"
tail -n 6 gpurun_out/r2c18_pytest.log
cat gpurun_out/r2c18_cutoff_ab.jsonl
python - <<'PY'
import json
d=json.load(open("gpurun_out/ensemble_8x1gpu_0.5.json")); print(d["structures_per_hour"], d["wall_seconds"], [(r["iterations"], round(r["minimize_s"],2), r.get("coarse_iterations")) for r in d["per_replica"]])
PY
