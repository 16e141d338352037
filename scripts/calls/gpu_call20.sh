#!/bin/bash
# round 2, call 20 (1 GPU): launch list of the coarse stage (where does an iteration's 1.27 ms go?)
mkdir -p gpurun_out
timeout 200 python scripts/ncu_target_coarse.py 0.5 20 > gpurun_out/r2c20_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c20_launches_coarse.csv python scripts/ncu_target_coarse.py 0.5 20 > gpurun_out/r2c20_ncu.log 2>&1
tail -n 2 gpurun_out/r2c20_plain.log; wc -l gpurun_out/r2c20_launches_coarse.csv
