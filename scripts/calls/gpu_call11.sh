#!/bin/bash
# round 2, call 11 (1 GPU): far field on clusters (EV tail + CHB) in the coarse stage: parity tests, two-stage timing, 8 members
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "far_field or graph or cutoff or sharded" > gpurun_out/r2c11_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c11_pytest.log
timeout 600 python scripts/gpu_cutoff_timing.py 0.5 gw > gpurun_out/r2c11_cutoff_timing.json 2> gpurun_out/r2c11_cutoff_timing.err
timeout 600 python scripts/gpu_ensemble.py 8 0 0.5 > gpurun_out/r2c11_ensemble_8x1.log 2>&1; echo "ens rc=$?" >> gpurun_out/r2c11_ensemble_8x1.log
tail -n 4 gpurun_out/r2c11_pytest.log
