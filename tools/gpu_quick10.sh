#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --workload ur5e_clik_osf --batch 4096 > gpurun_out/${tag}_bench_ur5e.json 2> gpurun_out/${tag}_bench_ur5e.err
timeout 600 python bench.py --steps 5 --warmup 3 --workload ur5e_clik_osf --batch 65536 > gpurun_out/${tag}_bench_ur5e_64k.json 2> gpurun_out/${tag}_bench_ur5e_64k.err
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_bench_ur5e.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"]/1e6,3), "Mcyc/s", d["ms_per_step"], "e2e", round(d["e2e"]["value"]/1e6,3), "cpu", round(d["cpu_baseline"]["value"]), "cpu1", round(d["cpu_baseline"]["single_thread"]["value"]), "launches", d["gpu_launches"])
    except Exception as e: print(f, "ERR", e)
PY
