#!/bin/bash
# ncu --set full of one launch of a kernel inside tools/walk_probe.py.  usage: ncu_probe.sh <kernel regex> <out name> <skip> [probe args...]
mkdir -p gpurun_out
K=$1; O=$2; S=$3; shift 3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o gpurun_out/$O python tools/walk_probe.py "$@" > gpurun_out/$O.log 2>&1; echo "ncu $K rc=$?"
