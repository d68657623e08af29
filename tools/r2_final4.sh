#!/bin/bash
# Round-2 closing verification on one GPU (after the PM step went back into stream order): whole GPU suite + the full bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02c_pytest_gpu.log
timeout 900 python bench.py 2>gpurun_out/r02c_bench.err | tee gpurun_out/r02c_bench.json | cut -c1-300
