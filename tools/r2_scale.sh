#!/bin/bash
# multi-GPU bench line: r2_scale.sh <ngpus> [extra bench args]
mkdir -p gpurun_out
N=$1; shift
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 "$@" 2>gpurun_out/r02_scale$N.err | tee gpurun_out/r02_scale$N.json | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('N', d['n_gpus'], 'ms', round(d['ms_per_step'],2), 'value %.4g' % d['value'], 'e2e ms', round(d['e2e']['ms_per_step'],2), 'parity', (d.get('parity') or {}).get('ok'), (d.get('parity') or {}).get('acc_max_err_over_mean_all_ranks'))
        print('phases', {k: round(v,1) for k,v in d['phases_ms'].items() if v and v < 1e5}); print('sharded', d.get('sharded_phases_ms'))
        print(json.dumps(d.get('hydro'))[:1200])
"
grep -v "^$" gpurun_out/r02_scale$N.err | grep -v "OMP_NUM\|\*\*\*\*" | head -12 | cut -c1-300
