#!/bin/bash
# round 2, call 8 (1 GPU): finer items of the cut-off kernel (parity + timing), default bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cutoff or sharded or graph or surrogate" > gpurun_out/r2c8_pytest_cutoff.log 2>&1; echo "rc=$?" >> gpurun_out/r2c8_pytest_cutoff.log
timeout 600 python scripts/gpu_cutoff_timing.py 0.5 gw > gpurun_out/r2c8_cutoff_timing.json 2> gpurun_out/r2c8_cutoff_timing.err
timeout 900 python bench.py > gpurun_out/r2c8_bench.json 2> gpurun_out/r2c8_bench.err; echo "bench rc=$?" >> gpurun_out/r2c8_bench.err
tail -n 3 gpurun_out/r2c8_pytest_cutoff.log
