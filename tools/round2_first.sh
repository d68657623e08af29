#!/bin/bash
# First GPU call of the next round: the step-loop CUDA code has only run under the CPU emulation so far.
# 1. hardware parity of csrc/steploop.cu against the reference goldens, 2. the usual GPU suite,
# 3. step-loop timings.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_step_gpu.py tests/test_domain_keys.py -q 2>&1 | tail -15 | tee gpurun_out/step_gpu.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/gpu_tests.log
timeout 900 python tools/steploop_bench.py 128 256 2>&1 | tail -30 | tee gpurun_out/steploop_bench.log
# memcheck of the new kernels on the hardware (the emulation ran them under AddressSanitizer only on the CPU)
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_step_gpu.py tests/test_domain_keys.py -q -k "primitives or domain_keys" 2>&1 | tail -15 | tee gpurun_out/step_memcheck.log
# and of a cross-section of the hardware-verified suite (never done in round 1); memcheck slows kernels 10-50x, so bounded
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py tests/test_pm_golden.py tests/test_sph.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/gpu_memcheck.log
# launch list of the step-loop kernels (per-launch durations; not a bench value)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_step_|k_domain_" -c 400 --csv --log-file gpurun_out/steploop_launches.csv python tools/steploop_bench.py 128 > gpurun_out/steploop_ncu.log 2>&1
tail -3 gpurun_out/steploop_ncu.log
