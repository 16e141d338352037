#!/bin/bash
# round 2, call 21 (1 GPU): far-field kernel staged through shared memory (FP64, one rsqrt): tests, coarse evaluation time,
# 16 ensemble members
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_driver.py -m gpu -q -x -k "far_field or cutoff or two_stage or coarse or ensemble" > gpurun_out/r2c21_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c21_pytest.log
timeout 240 python scripts/gpu_cutoff_ab.py 0.5 > gpurun_out/r2c21_cutoff_ab.jsonl 2> gpurun_out/r2c21_cutoff_ab.err
timeout 600 python scripts/gpu_ensemble.py 16 0 0.5 > gpurun_out/r2c21_ensemble_16x1.log 2>&1; echo "ens rc=$?"
tail -n 4 gpurun_out/r2c21_pytest.log
cat gpurun_out/r2c21_cutoff_ab.jsonl
python - <<'PY'
import json
d=json.load(open("gpurun_out/ensemble_16x1gpu_0.5.json")); print(d["structures_per_hour"], d["wall_seconds"], [(r["iterations"], round(r["minimize_s"],2), r.get("coarse_iterations"), r.get("coarse_rounds")) for r in d["per_replica"]])
PY
