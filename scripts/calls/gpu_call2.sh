#!/bin/bash
# round 2, call 2: new cut-off path (parity + timing), issue micro-benchmark with MUFU-saving variants, full suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cutoff" > gpurun_out/r2c2_pytest_cutoff.log 2>&1; echo "rc=$?" >> gpurun_out/r2c2_pytest_cutoff.log
timeout 600 python scripts/gpu_cutoff_timing.py 0.5 gw > gpurun_out/r2c2_cutoff_timing.json 2> gpurun_out/r2c2_cutoff_timing.err
timeout 1500 python -m pytest tests -m gpu -q -rA --durations=10 > gpurun_out/r2c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c2_pytest.log
tail -n 3 gpurun_out/r2c2_pytest.log
