#!/bin/bash
# GPU parity tests + the bench's hydro section (SPH density / hydro timings)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --no-cpu --steps 2 > gpurun_out/bench_quick.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
h = d["hydro"]
print(d["ms_per_step"], "density", h["density_ms"], "hydro", h["hydro_ms"], h["density_passes_mean"], h["neighbours_mean"])
PY
