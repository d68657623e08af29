#!/bin/bash
# one GPU call over experiment builds: r2_var.sh <ng> <state> <variant> [variant...]   ("new" = the in-tree library)
mkdir -p gpurun_out
NG=$1; ST=$2; shift 2
{
for v in "$@"; do
  if [ "$v" == "new" ]; then unset B200_LIB; else export B200_LIB=$PWD/build_variants/$v/libb200force.so; fi
  echo "== $v"
  timeout 300 python tools/walk_probe.py $NG $ST 4 2>&1 | tail -3
done
} 2>&1 | tee -a gpurun_out/r2_var.log
