#!/bin/bash
# round 2, call 24 (1 GPU): tensor-core feasibility micro-benchmark (mma.sync TF32 rate; hot body with r^2 / i-side force by MMA)
mkdir -p gpurun_out
timeout 120 scripts/bin/microbench_mma > gpurun_out/r2c24_microbench_mma.log 2>&1
cat gpurun_out/r2c24_microbench_mma.log
