#!/bin/bash
# usage: tools/scale.sh NGPUS NG
N=$1; NG=${2:-256}
cd "$(dirname "$0")/.." && python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --ng $NG --no-cpu 2>&1 | tail -25
