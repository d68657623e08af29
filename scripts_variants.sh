#!/bin/bash
for v in "" _mb6 _mb8; do
  echo "variant [$v]"
  B200_LIB=$PWD/mp-gadget_b200/libb200force$v.so python bench.py --no-cpu --steps 3 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'walk',d['phases_ms']['walk'])"
done
