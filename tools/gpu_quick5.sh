#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
for mb in 3 4; do for ct in 64 128; do DRC_COL_MINB=$mb DRC_COL_THREADS=$ct timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_colmb${mb}_t${ct}.json 2>/dev/null; done; done
DRC_COL_THREADS=64 timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_colmb2_t64.json 2>/dev/null
tail -3 gpurun_out/${tag}_smoke.log; tail -6 gpurun_out/${tag}_pytest.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"]/1e6,3), "Mcyc/s nohint", round(d["value_no_schedule_hint"]/1e6,3), d["roofline"]["stage_ms"], "e2e", round(d["e2e"]["value"]/1e6,3), "launches", d["gpu_launches"])
    except Exception as e: print(f, "ERR", e)
PY
