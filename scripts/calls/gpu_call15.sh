#!/bin/bash
# round 2, call 15 (1 GPU): full GPU suite after the amd integrator / device-side force negation / half-stage items;
# small systems; short bench (e2e)
mkdir -p gpurun_out
export OMP_NUM_THREADS=$(nproc)
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2c15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c15_pytest.log
timeout 300 python scripts/gpu_small_systems.py > gpurun_out/r2c15_small_systems.json 2> gpurun_out/r2c15_small_systems.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-minimize-full --no-ensemble --minimize-iters 5 > gpurun_out/r2c15_bench_short.json 2> gpurun_out/r2c15_bench_short.err
tail -n 15 gpurun_out/r2c15_pytest.log
cat gpurun_out/r2c15_small_systems.json | cut -c1-1500
python - <<'PY'
import json
for l in open("gpurun_out/r2c15_bench_short.json"):
    if l.startswith("{"):
        d=json.loads(l); print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches")}, d["roofline"]["frac_nominal"])
PY
