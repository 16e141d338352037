#!/bin/bash
# round 2, call 23 (1 GPU): exact pair kernel A/B — j-side columns as a plain store instead of a read-modify-write
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-minimize-full --no-ensemble --minimize-iters 0"
for rep in 1 2; do
for lib in libmultimm_b200_old.so libmultimm_b200.so; do
  MMM_LIB_NAME=$lib timeout 200 $B 2>> gpurun_out/r2c23.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$lib', d['ms_per_step'], d['roofline']['pair_kernel_ms'], d['roofline']['frac_nominal'], d['energy_terms']['EV'])
" | tee -a gpurun_out/r2c23_ab.txt
done; done
