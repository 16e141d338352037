#!/bin/bash
# round 2, call 1: full GPU test suite, default bench line, issue-port micro-benchmark, reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2c1_gpu.txt 2>&1
nproc >> gpurun_out/r2c1_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -rA --durations=15 > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c1_pytest.log
timeout 120 scripts/bin/microbench_issue > gpurun_out/r2c1_microbench_issue.log 2>&1
timeout 900 python bench.py > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; echo "bench rc=$?" >> gpurun_out/r2c1_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c1_bench_ref.json 2> gpurun_out/r2c1_bench_ref.err
tail -n 3 gpurun_out/r2c1_pytest.log
