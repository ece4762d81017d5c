#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ur5e.py -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
for mi in 200 400 1000; do DRC_DEBUG_MAX_ITER=$mi timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_maxiter$mi.json 2>/dev/null; done
timeout 600 python bench.py --steps 5 --warmup 3 --workload fr3_qpid > gpurun_out/${tag}_bench_fr3_qpid.json 2> gpurun_out/${tag}_bench_fr3_qpid.err
tail -8 gpurun_out/${tag}_pytest.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"]/1e6,3), "Mcyc/s", d["roofline"]["stage_ms"], "e2e", round(d["e2e"]["value"]/1e6,3), "solved", d["solved_fraction"], "iters", d["mean_admm_iters"])
    except Exception as e: print(f, "ERR", e)
PY
