#!/bin/bash
tools/r2_var.sh 128 displaced new qr12 qr28 qr33
tools/r2_var.sh 128 clustered new qr12 qr28 qr33
tools/r2_var.sh 128 z9 new
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_config_parity.py -m gpu -x -q 2>&1 | tail -5
