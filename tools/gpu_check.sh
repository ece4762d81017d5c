#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --workload fr3_qpid > gpurun_out/${tag}_bench_fr3_qpid.json 2> gpurun_out/${tag}_bench_fr3_qpid.err
tail -3 gpurun_out/${tag}_smoke.log; tail -6 gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"]/1e6,3), "Mcyc/s nohint", round(d["value_no_schedule_hint"]/1e6,3), d["ms_per_step"], d["roofline"]["stage_ms"], "e2e", round(d["e2e"]["value"]/1e6,3), "launches", d["gpu_launches"], "solved", d["solved_fraction"], d["mean_admm_iters"])
    except Exception as e: print(f, "ERR", e)
PY
