#!/bin/bash
# ncu --set full of the two tree-gravity kernels (one launch each, after warm-up)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_grav_pairs -s 3 -c 1 -f -o gpurun_out/prof_pairs python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_pairs.log 2>&1; echo "ncu pairs rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_grav_walk -s 3 -c 1 -f -o gpurun_out/prof_walk python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_walk.log 2>&1; echo "ncu walk rc=$?"
