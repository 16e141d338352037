#!/bin/bash
# round 2, call 16 (1 GPU): cut-off pass A/B — previous kernel / flat tile cull + prefetch / the same with 128-thread CTAs
mkdir -p gpurun_out
for lib in libmultimm_b200_cwold.so libmultimm_b200.so libmultimm_b200_cw128.so; do
  MMM_LIB_NAME=$lib timeout 240 python scripts/gpu_cutoff_ab.py 0.5 >> gpurun_out/r2c16_cutoff_ab.jsonl 2>> gpurun_out/r2c16_cutoff_ab.err
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -q -x -k "far_field or cutoff or sharded or cells or outlier or one_system" > gpurun_out/r2c16_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c16_pytest.log
cat gpurun_out/r2c16_cutoff_ab.jsonl
tail -n 4 gpurun_out/r2c16_pytest.log
