#!/bin/bash
# round 2, call 5 (2 GPUs): the real two-GPU tests, the bench line at N = 2 (replicas + ensemble + sharded stress system),
# the ensemble driver on two GPUs.  Every command under its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2c5_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_analysis.py -m gpu -q -rA > gpurun_out/r2c5_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c5_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2c5_bench_2gpu.json 2> gpurun_out/r2c5_bench_2gpu.err; echo "bench rc=$?" >> gpurun_out/r2c5_bench_2gpu.err
timeout 600 python scripts/gpu_ensemble.py 4 0,1 0.5 > gpurun_out/r2c5_ensemble_2gpu.log 2>&1; echo "ens rc=$?" >> gpurun_out/r2c5_ensemble_2gpu.log
tail -n 4 gpurun_out/r2c5_pytest_multi.log
