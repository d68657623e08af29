#!/bin/bash
# first GPU contact: smoke, parity tests, quick timing at 128^3 / 256^3
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python scripts_quicktime.py > gpurun_out/quicktime.log 2>&1; echo "quick rc=$?" >> gpurun_out/quicktime.log
tail -5 gpurun_out/smoke.log; tail -30 gpurun_out/pytest_gpu.log; tail -40 gpurun_out/quicktime.log
