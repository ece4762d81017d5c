#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
for mb in 2 4; do DRC_ADMM_MINB=$mb timeout 300 python bench.py --steps 8 --warmup 3 > gpurun_out/${tag}_bench_admm_mb${mb}.json 2>/dev/null; done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"]/1e6,3), "Mcyc/s nohint", round(d["value_no_schedule_hint"]/1e6,3), d["roofline"]["stage_ms"], "e2e", round(d["e2e"]["value"]/1e6,3))
    except Exception as e: print(f, "ERR", e)
PY
bash tools/gpu_round.sh r01d
