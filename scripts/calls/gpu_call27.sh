#!/bin/bash
# round 2, call 27 (1 GPU): ensemble workers with the early CUDA start (two workers on GPU 0), A/B against MMM_NO_EARLY_CUDA=1;
# the ensemble driver tests
mkdir -p gpurun_out
timeout 150 python scripts/gpu_ensemble_startup.py > gpurun_out/r2c27_startup.json 2> gpurun_out/r2c27_startup.err; echo "startup rc=$?"
timeout 100 python -m pytest tests/test_gpu_driver.py -m gpu -q -x -k "ensemble" > gpurun_out/r2c27_pytest.log 2>&1; echo "pytest rc=$?"
cat gpurun_out/r2c27_startup.json; tail -n 3 gpurun_out/r2c27_startup.err; tail -n 2 gpurun_out/r2c27_pytest.log
