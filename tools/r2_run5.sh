#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded.py -m gpu -x -q 2>&1 | tail -3
tools/r2_run4.sh ${1:-128}
