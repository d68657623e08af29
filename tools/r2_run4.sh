#!/bin/bash
# 2-GPU: the sharded bench line with the SPH entry at a reduced size
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 1 --ng ${1:-64} --sharded-hydro --no-parity 2>gpurun_out/r2_sh2.err | tee gpurun_out/r2_sh2.json | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step']); print(json.dumps(d.get('hydro'))[:1500])
"
grep -v "^$" gpurun_out/r2_sh2.err | grep -v "OMP_NUM\|\*\*\*\*" | head -30 | cut -c1-300
