#!/bin/bash
# SPH GPU tests + the density / hydro timing probe.  args: values of B200_SPH_TPW to try (0 = adaptive)
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_sph.py tests/test_dropin.py -m gpu -x -q 2>&1 | tail -3
for v in "$@"; do
  export B200_SPH_TPW=$v
  echo "== tpw $v"; timeout 600 python tools/sph_prof.py 128 2 2>&1 | tail -2
done
} 2>&1 | tee -a gpurun_out/r2_sph.log
