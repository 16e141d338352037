#!/bin/bash
# round 2, call 12 (8 GPUs): configs[3] — the genome-wide ensemble of 64 structures over 8 GPUs (two-stage with the far
# field on clusters), and the bench line at N = 8 with the final code
mkdir -p gpurun_out
echo "start $(date +%s)" > gpurun_out/r2c12_times.txt
timeout 400 python scripts/gpu_ensemble.py 64 0,1,2,3,4,5,6,7 0.5 > gpurun_out/r2c12_ensemble_64x8.log 2>&1; echo "ens64 rc=$? $(date +%s)" >> gpurun_out/r2c12_times.txt
timeout 420 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8 --master-port 29641 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2c12_bench_8gpu.json 2> gpurun_out/r2c12_bench_8gpu.err; echo "bench8 rc=$? $(date +%s)" >> gpurun_out/r2c12_times.txt
cat gpurun_out/r2c12_times.txt
