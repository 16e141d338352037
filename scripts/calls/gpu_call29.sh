#!/bin/bash
# round 2, call 29 (1 GPU): last check of the final state: the GPU suite without its two long oracle tests, smoke()
mkdir -p gpurun_out
timeout 75 python -m pytest tests -m gpu -q -x -k "not s1_minimised and not baseline_config" > gpurun_out/r2c29_pytest.log 2>&1; echo "pytest rc=$?"
timeout 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c29_smoke.log 2>&1; echo "smoke rc=$?"
tail -n 3 gpurun_out/r2c29_pytest.log; cat gpurun_out/r2c29_smoke.log
