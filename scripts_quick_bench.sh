#!/bin/bash
# usage: scripts_quick_bench.sh [pytest]
if [ "$1" == "pytest" ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -5; fi
python bench.py --no-cpu --steps 3 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e', d['e2e']['value']); print({k:round(v,2) for k,v in d['phases_ms'].items()})"
