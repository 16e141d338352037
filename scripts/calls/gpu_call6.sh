#!/bin/bash
# round 2, call 6 (2 GPUs): warp-per-item cut-off kernel (parity + timing, on GPU 0), the global NVLink work queue
# (real 2-GPU tests; sharded stress system with the queue and with static dealing)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cutoff or sharded or graph or surrogate" > gpurun_out/r2c6_pytest_cutoff.log 2>&1; echo "rc=$?" >> gpurun_out/r2c6_pytest_cutoff.log
timeout 600 python scripts/gpu_cutoff_timing.py 0.5 gw > gpurun_out/r2c6_cutoff_timing.json 2> gpurun_out/r2c6_cutoff_timing.err
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -rA -s -k "one_system" > gpurun_out/r2c6_pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r2c6_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 5 --warmup 3 --no-ensemble > gpurun_out/r2c6_bench_2gpu_queue.json 2> gpurun_out/r2c6_bench_2gpu_queue.err
MMM_DIST_STATIC=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 5 --warmup 3 --no-ensemble > gpurun_out/r2c6_bench_2gpu_static.json 2> gpurun_out/r2c6_bench_2gpu_static.err
tail -n 3 gpurun_out/r2c6_pytest_cutoff.log gpurun_out/r2c6_pytest_multi.log
