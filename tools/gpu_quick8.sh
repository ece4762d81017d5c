#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
for b in 96 1776 5328; do DRC_DEBUG_EPS=1e-14 timeout 300 python bench.py --steps 3 --warmup 3 --batch $b > gpurun_out/${tag}_lat_b$b.json 2>gpurun_out/${tag}_lat_b$b.err; done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_lat*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["roofline"]["stage_ms"], "iters", d["mean_admm_iters"])
    except Exception as e: print(f, "ERR", e)
PY
