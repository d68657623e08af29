#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_config_parity.py tests/test_sph.py -m gpu -x -q 2>&1 | tail -5
for b in 0 1; do
  if [ $b == 1 ]; then export B200_E2E_BULK=1; else unset B200_E2E_BULK; fi
  timeout 600 python bench.py --no-cpu --no-hydro --no-states --no-steploop --steps 3 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
e=d['e2e']; print('bulk=$b dev ms', round(d['ms_per_step'],2), 'e2e ms', round(e['ms_per_step'],2), 'h2d_ms', round(e['h2d_ms'],2), 'd2h_ms', round(e['d2h_ms'],2), 'bytes', e['h2d_bytes_per_step'], e['d2h_bytes_per_step'], 'chk', e['check_vs_device_arm'])
"
done
