#!/bin/bash
# one GPU call: the walk / pair experiment builds on the displaced 128^3 and 256^3 boxes
mkdir -p gpurun_out
{
for v in base "" nseg8 nseg2 nseg1 pair3 walk5; do
  if [ -z "$v" ]; then unset B200_LIB; echo "== new"; else export B200_LIB=$PWD/build_variants/$v/libb200force.so; echo "== $v"; fi
  timeout 300 python tools/walk_probe.py 128 displaced 4 2>&1 | tail -4
done
for v in base "" nseg8; do
  if [ -z "$v" ]; then unset B200_LIB; echo "== new 256"; else export B200_LIB=$PWD/build_variants/$v/libb200force.so; echo "== $v 256"; fi
  timeout 300 python tools/walk_probe.py 256 displaced 4 2>&1 | tail -3
done
unset B200_LIB
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_config_parity.py -m gpu -x -q 2>&1 | tail -5
} 2>&1 | tee gpurun_out/r2_walkvar.log
