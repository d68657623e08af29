#!/bin/bash
# ncu --set full of one tree-gravity kernel launch (after warm-up).  usage: ncu_walk.sh <kernel regex> <out name> [bench args...]
mkdir -p gpurun_out
K=$1; O=$2; shift 2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/$O python bench.py --steps 1 --warmup 3 --no-cpu --no-hydro "$@" > gpurun_out/$O.log 2>&1; echo "ncu $K rc=$?"
