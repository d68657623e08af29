#!/bin/bash
# the SPH timing probe over experiment builds: sph_var.sh <variant|new> ...
mkdir -p gpurun_out
{
for v in "$@"; do
  if [ "$v" == "new" ]; then unset B200_LIB; else export B200_LIB=$PWD/build_variants/$v/libb200force.so; fi
  echo "== $v"; timeout 600 python tools/sph_prof.py 128 2 2>&1 | tail -2
done
} 2>&1 | tee -a gpurun_out/r2_sphvar.log
