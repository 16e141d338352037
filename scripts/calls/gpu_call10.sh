#!/bin/bash
# round 2, call 10 (1 GPU): warm start of the exact stage (test + two-stage timing over several seeds), the 1-GPU baseline of the
# N = 2e6 stress system, 8 ensemble members on one GPU (the 1-GPU leg of configs[3])
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "warm or graph or cutoff" > gpurun_out/r2c10_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c10_pytest.log
timeout 600 python scripts/gpu_cutoff_timing.py 0.5 gw > gpurun_out/r2c10_cutoff_timing.json 2> gpurun_out/r2c10_cutoff_timing.err
timeout 300 python bench.py --workload stress --steps 3 --warmup 3 --no-cpu --no-ensemble --no-minimize-full --minimize-iters 0 > gpurun_out/r2c10_bench_stress_1gpu.json 2> gpurun_out/r2c10_bench_stress_1gpu.err
timeout 600 python scripts/gpu_ensemble.py 8 0 0.5 > gpurun_out/r2c10_ensemble_8x1.log 2>&1; echo "ens rc=$?" >> gpurun_out/r2c10_ensemble_8x1.log
tail -n 4 gpurun_out/r2c10_pytest.log
