#!/bin/bash
# Round-2 closing verification on one GPU after the PM transform passes went in: whole GPU suite, full bench line, ncu launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02b_pytest_gpu.log
timeout 900 python bench.py 2>gpurun_out/r02b_bench.err | tee gpurun_out/r02b_bench.json | cut -c1-400
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-hydro --no-states --no-steploop > gpurun_out/r02b_ncu_list.log 2>&1; echo "ncu list rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
