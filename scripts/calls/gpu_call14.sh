#!/bin/bash
# round 2, call 14 (1 GPU): coarse / exact-probe rounds: 16 ensemble members on one GPU; phase-ordered hot body in the micro-benchmark
mkdir -p gpurun_out
timeout 120 scripts/bin/microbench_issue > gpurun_out/r2c14_microbench_issue.log 2>&1
timeout 900 python scripts/gpu_ensemble.py 16 0 0.5 > gpurun_out/r2c14_ensemble_16x1.log 2>&1; echo "ens rc=$?" >> gpurun_out/r2c14_ensemble_16x1.log
grep "hot" gpurun_out/r2c14_microbench_issue.log | tail -6
