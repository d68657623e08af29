#!/bin/bash
mkdir -p gpurun_out
for v in $@; do
B200_WALK_VARIANT=$v timeout 600 python bench.py --no-cpu --steps 3 > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err; echo "variant $v bench rc=$?"; tail -3 gpurun_out/bench_v$v.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_v$v.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'], 'walk', d['phases_ms']['walk'])
PY
done
