#!/bin/bash
# round 2, call 9 (8 GPUs): the bench line at N = 8 and N = 4 (replicas + ensemble + sharded stress system, work queue vs
# static dealing), the ensemble driver over 8 GPUs (two-stage and exact).  Every command under a tight timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2c9_gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "start $(date +%s)" > gpurun_out/r2c9_times.txt
timeout 420 $TR --nproc-per-node 8 --master-port 29631 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2c9_bench_8gpu.json 2> gpurun_out/r2c9_bench_8gpu.err; echo "bench8 rc=$? $(date +%s)" >> gpurun_out/r2c9_times.txt
MMM_DIST_STATIC=1 timeout 300 $TR --nproc-per-node 8 --master-port 29632 bench.py --gpus 8 --steps 5 --warmup 3 --no-ensemble > gpurun_out/r2c9_bench_8gpu_static.json 2> gpurun_out/r2c9_bench_8gpu_static.err; echo "bench8 static rc=$? $(date +%s)" >> gpurun_out/r2c9_times.txt
timeout 300 python scripts/gpu_ensemble.py 16 0,1,2,3,4,5,6,7 0.5 > gpurun_out/r2c9_ensemble_16x8.log 2>&1; echo "ens16 rc=$? $(date +%s)" >> gpurun_out/r2c9_times.txt
timeout 420 $TR --nproc-per-node 4 --master-port 29633 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2c9_bench_4gpu.json 2> gpurun_out/r2c9_bench_4gpu.err; echo "bench4 rc=$? $(date +%s)" >> gpurun_out/r2c9_times.txt
timeout 300 python scripts/gpu_ensemble.py 8 0,1,2,3,4,5,6,7 0.0 > gpurun_out/r2c9_ensemble_8x8_exact.log 2>&1; echo "ens8 exact rc=$? $(date +%s)" >> gpurun_out/r2c9_times.txt
cat gpurun_out/r2c9_times.txt
