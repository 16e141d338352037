#!/bin/bash
# round 2, call 30 (1 GPU, the last seconds of the budget): bench.py's ensemble object on the pipelined driver
mkdir -p gpurun_out
timeout 40 python bench.py --steps 3 --warmup 3 --no-cpu --no-minimize-full --minimize-iters 5 --ensemble-members 1 > gpurun_out/r2c30_bench.json 2> gpurun_out/r2c30_bench.err; echo "rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2c30_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["value"], d["e2e"]["value"], d["minimize"], {k: d["ensemble"].get(k) for k in ("structures_per_hour","error","members")})
PY
tail -n 3 gpurun_out/r2c30_bench.err
