#!/bin/bash
mkdir -p gpurun_out
pick='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); s=d.get("steploop",{}); print(sys.argv[1], "ms_per_step", round(d["ms_per_step"],2), "steploop", [(round(x["wall_ms"],1), x.get("stages_ms")) for x in s.get("substeps",[])][:2] or s)'
( timeout 400 python bench.py --no-cpu --no-states 2>/dev/null | python -c "$pick" own-after-hydro ) | tee gpurun_out/r2_diag2.log
