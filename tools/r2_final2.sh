#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k_|Radix|fft' -c 500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-hydro --no-states --no-steploop > gpurun_out/r02_ncu_list.log 2>&1; echo "ncu list rc=$?"
grep -c k_grav gpurun_out/r02_launches.csv
