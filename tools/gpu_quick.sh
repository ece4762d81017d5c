#!/bin/bash
# quick GPU pass: parity tests + bench + launch-config experiments
tag=${1:-q}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
for ct in 32 128; do DRC_COL_THREADS=$ct timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_col${ct}.json 2>/dev/null; done
for jt in 32 128; do DRC_JOB_THREADS=$jt timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_job${jt}.json 2>/dev/null; done
tail -5 gpurun_out/${tag}_pytest.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"]/1e6,2), "Mcyc/s", d["roofline"]["stage_ms"], "e2e", round(d["e2e"]["value"]/1e6,2))
    except Exception as e: print(f, "ERR", e)
PY
