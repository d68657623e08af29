#!/bin/bash
# Round-end measurement set: GPU parity tests, the default bench line, the reference arm,
# the ncu launch list of one bench step and full captures of the two tree-gravity kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err; cut -c1-600 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$?"; grep -v "^\[" gpurun_out/bench_reference.json | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-hydro > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_grav_pairs -s 3 -c 1 -f -o gpurun_out/prof_pairs python bench.py --steps 1 --warmup 3 --no-cpu --no-hydro > gpurun_out/ncu_pairs.log 2>&1; echo "ncu pairs rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_grav_walk -s 3 -c 1 -f -o gpurun_out/prof_walk python bench.py --steps 1 --warmup 3 --no-cpu --no-hydro > gpurun_out/ncu_walk.log 2>&1; echo "ncu walk rc=$?"
ls -la gpurun_out | head -30
