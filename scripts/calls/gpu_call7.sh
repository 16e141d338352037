#!/bin/bash
# round 2, call 7 (1 GPU): finer Morton keys (parity + timing); ncu launch list of the bench command and --set full
# captures of the exact pair kernel and of the cut-off kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cutoff or sharded or graph or surrogate" > gpurun_out/r2c7_pytest_cutoff.log 2>&1; echo "rc=$?" >> gpurun_out/r2c7_pytest_cutoff.log
timeout 600 python scripts/gpu_cutoff_timing.py 0.5 gw > gpurun_out/r2c7_cutoff_timing.json 2> gpurun_out/r2c7_cutoff_timing.err
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu --no-minimize-full --no-ensemble --minimize-iters 5"
timeout 300 $BENCH > gpurun_out/r2c7_bench_short.json 2> gpurun_out/r2c7_bench_short.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c7_launches_bench.csv $BENCH > gpurun_out/r2c7_ncu1.log 2>&1
timeout 300 python scripts/ncu_target.py gw > gpurun_out/r2c7_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pair_n3 -s 1 -c 1 -o gpurun_out/r2c7_pair_n3 -f python scripts/ncu_target.py gw > gpurun_out/r2c7_ncu2.log 2>&1
timeout 300 python scripts/ncu_target_cutoff.py 0.5 > gpurun_out/r2c7_plain3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pair_cut_warp -s 2 -c 1 -o gpurun_out/r2c7_pair_cut_warp -f python scripts/ncu_target_cutoff.py 0.5 > gpurun_out/r2c7_ncu3.log 2>&1
tail -n 3 gpurun_out/r2c7_pytest_cutoff.log
