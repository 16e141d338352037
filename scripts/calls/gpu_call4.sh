#!/bin/bash
# round 2, call 4: packed near body + CHB clusters + graph replay: parity suite, cut-off timing, small systems, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_driver.py tests/test_zz_gpu_reference_expressions.py -m gpu -q -rA > gpurun_out/r2c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c4_pytest.log
timeout 300 python scripts/gpu_small_systems.py > gpurun_out/r2c4_small_systems.json 2> gpurun_out/r2c4_small_systems.err
timeout 600 python scripts/gpu_cutoff_timing.py 0.5 gw > gpurun_out/r2c4_cutoff_timing.json 2> gpurun_out/r2c4_cutoff_timing.err
timeout 900 python bench.py > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err; echo "bench rc=$?" >> gpurun_out/r2c4_bench.err
tail -n 3 gpurun_out/r2c4_pytest.log
