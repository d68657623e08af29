#!/bin/bash
# ncu --set full of one launch of an SPH kernel inside tools/sph_prof.py.  usage: ncu_sph.sh <kernel regex> <out name> <skip>
mkdir -p gpurun_out
K=$1; O=$2; S=$3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o gpurun_out/$O python tools/sph_prof.py 128 1 > gpurun_out/$O.log 2>&1; echo "ncu $K rc=$?"
