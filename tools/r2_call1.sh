#!/bin/bash
# Round 2, first GPU call: hardware run of the step-loop / domain-key kernels (so far emulation-only) + the full GPU suite.
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/r2c1_gpus.log
export B200_RUN_UNVERIFIED=1
timeout 900 python -m pytest tests/test_step_gpu.py tests/test_domain_keys.py -q 2>&1 | tail -40 | tee gpurun_out/r2c1_step_gpu.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/r2c1_gpu_tests.log
timeout 600 python tools/steploop_bench.py 128 256 2>&1 | tail -30 | tee gpurun_out/r2c1_steploop_bench.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_step_gpu.py tests/test_domain_keys.py -q -k "primitives or domain_keys" 2>&1 | tail -15 | tee gpurun_out/r2c1_step_memcheck.log
