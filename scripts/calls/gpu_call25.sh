#!/bin/bash
# round 2, call 25 (2 GPUs): the driver's N = 2 bench command with the final code, and the real 2-GPU tests
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2c25_gpus.txt
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2c25_pytest_two_gpus.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c25_pytest_two_gpus.log
timeout 420 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c25_bench_2gpu.json 2> gpurun_out/r2c25_bench_2gpu.err; echo "bench2 rc=$?"
tail -n 3 gpurun_out/r2c25_pytest_two_gpus.log
python - <<'PY'
import json
for l in open("gpurun_out/r2c25_bench_2gpu.json"):
    if l.startswith("{"):
        d=json.loads(l); print({k:d[k] for k in ("value","n_gpus","ms_per_step","e2e")}, d["ensemble"]["structures_per_hour"], d["decomposed"]["ms_per_evaluation"], d["decomposed"]["exchange_ms"])
PY
