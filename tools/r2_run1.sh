#!/bin/bash
tools/r2_var.sh 128 displaced new
tools/r2_var.sh 128 z9 new
tools/r2_var.sh 128 clustered new
tools/r2_var.sh 256 displaced new
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_config_parity.py -m gpu -x -q 2>&1 | tail -5
