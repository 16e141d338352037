#!/bin/bash
# round 2, call 22 (1 GPU): generic functional forms on the Newton-3 machinery: parity tests, timing against the gather kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_reference_expressions.py tests/test_gpu_multi.py -m gpu -q -k "not s1_minimised" > gpurun_out/r2c22_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c22_pytest.log
timeout 300 python scripts/gpu_generic_timing.py 50000 > gpurun_out/r2c22_generic_timing.json 2> gpurun_out/r2c22_generic_timing.err
tail -n 12 gpurun_out/r2c22_pytest.log; cat gpurun_out/r2c22_generic_timing.json; tail -n 3 gpurun_out/r2c22_generic_timing.err
