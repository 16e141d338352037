#!/bin/bash
# round 2, call 13 (1 GPU): coarse stage converged to half the tolerance: 16 ensemble members on one GPU; MD / driver tests
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_md.py tests/test_gpu_driver.py -m gpu -q -x > gpurun_out/r2c13_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c13_pytest.log
timeout 900 python scripts/gpu_ensemble.py 16 0 0.5 > gpurun_out/r2c13_ensemble_16x1.log 2>&1; echo "ens rc=$?" >> gpurun_out/r2c13_ensemble_16x1.log
tail -n 3 gpurun_out/r2c13_pytest.log
