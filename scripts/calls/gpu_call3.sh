#!/bin/bash
# round 2, call 3: launch list of cut-off evaluations (ncu), sharded cut-off test, S1 minimisation test
mkdir -p gpurun_out
timeout 300 python scripts/ncu_target_cutoff.py > gpurun_out/r2c3_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2c3_launches_cutoff.csv python scripts/ncu_target_cutoff.py > gpurun_out/r2c3_ncu.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sharded" > gpurun_out/r2c3_pytest_sharded.log 2>&1; echo "rc=$?" >> gpurun_out/r2c3_pytest_sharded.log
timeout 900 python -m pytest tests/test_gpu_parity_full.py -m gpu -q -x -k "s1_minimised" > gpurun_out/r2c3_pytest_s1.log 2>&1; echo "rc=$?" >> gpurun_out/r2c3_pytest_s1.log
tail -n 3 gpurun_out/r2c3_pytest_sharded.log gpurun_out/r2c3_pytest_s1.log
