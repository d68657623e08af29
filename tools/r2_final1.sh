#!/bin/bash
# Round-2 verification on one GPU: the whole GPU suite, the full bench line, the reference arm, ncu launch list and captures.
mkdir -p gpurun_out
nvidia-smi -L | head -2
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py 2>gpurun_out/r02_bench.err | tee gpurun_out/r02_bench.json | cut -c1-300
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>gpurun_out/r02_bench_ref.err | tee gpurun_out/r02_bench_reference.json | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-hydro --no-states --no-steploop > gpurun_out/r02_ncu_list.log 2>&1; echo "ncu list rc=$?"
tools/ncu_probe.sh k_grav_walk r02_prof_walk 2 256 displaced 3
tools/ncu_probe.sh k_grav_pairs r02_prof_pairs 2 256 displaced 3
