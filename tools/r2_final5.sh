#!/bin/bash
# ncu launch list of one bench step (our kernels + CUB only) and a check of the step-loop entry with its warm-up sub-step
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_|Radix|fft" -c 400 --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-hydro --no-states --no-steploop --no-extras > gpurun_out/r02c_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 300 python bench.py --no-cpu --no-hydro --no-states --no-extras --steps 2 --warmup 3 2>/dev/null | python -c '
import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); s=d.get("steploop",{}); print("ms_per_step", round(d["ms_per_step"],2), "steploop", [(round(x["wall_ms"],1), x.get("stages_ms")) for x in s.get("substeps",[])][:3] or s)' | tee gpurun_out/r02c_steploop.log
