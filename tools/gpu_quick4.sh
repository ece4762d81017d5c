#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
for w in fr3_qpid husky_qpik xls_qpik; do timeout 600 python bench.py --steps 5 --warmup 3 --workload $w > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err; done
tail -8 gpurun_out/${tag}_pytest.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"]/1e6,3), "Mcyc/s nohint", round(d["value_no_schedule_hint"]/1e6,3), d["roofline"]["stage_ms"], "e2e", round(d["e2e"]["value"]/1e6,3), "solved", d["solved_fraction"], "iters", d["mean_admm_iters"], "launches", d["gpu_launches"])
    except Exception as e: print(f, "ERR", e)
PY
