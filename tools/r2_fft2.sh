#!/bin/bash
# pm_fft.cu tile-width / occupancy variants at 768^3 (build_variants/t4*: tools/build_variant.sh NAME pm_fft.cu "-DFFT_T=4 ...")
mkdir -p gpurun_out
export PM_PROBE_ONLY=own
( echo "T8 256"; timeout 200 python tools/pm_probe.py 256 768 2>&1 | tail -1
  for v in t4 t4b3; do for th in 256 128; do
    echo "$v $th"; B200_LIB=build_variants/$v/libb200force.so B200_FFT_THREADS=$th timeout 200 python tools/pm_probe.py 256 768 2>&1 | tail -1
  done; done ) | tee gpurun_out/r2_fft_variants.log
