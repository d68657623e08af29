#!/bin/bash
# First hardware run of csrc/pm_fft.cu: PM parity tests, timing against cuFFT at 768^3, one ncu capture of its kernels.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_pm_golden.py -m gpu -q -k "pm or empty" 2>&1 | tail -8 | tee gpurun_out/r2_fft_tests.log
timeout 300 python tools/pm_probe.py 256 768 2>&1 | tail -4 | tee gpurun_out/r2_fft_probe.log
B200_FFT_THREADS=128 PM_PROBE_ONLY=own timeout 300 python tools/pm_probe.py 256 768 2>&1 | tail -1 | tee -a gpurun_out/r2_fft_probe.log
B200_FFT_THREADS=192 PM_PROBE_ONLY=own timeout 300 python tools/pm_probe.py 256 768 2>&1 | tail -1 | tee -a gpurun_out/r2_fft_probe.log
PM_PROBE_ONLY=own PM_PROBE_ITERS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_fft -c 5 -o gpurun_out/r2_fft_ncu -f python tools/pm_probe.py 256 768 > gpurun_out/r2_fft_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python -m pytest tests/test_config_parity.py -m gpu -q 2>&1 | tail -4 | tee -a gpurun_out/r2_fft_tests.log
