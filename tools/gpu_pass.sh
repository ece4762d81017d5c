#!/bin/bash
# round-2 GPU pass: reference-library probe, ADMM lab A/B, parity tests, bench (usage: tools/gpu_pass.sh <tag>)
TAG=${1:-r2}
mkdir -p gpurun_out
python tools/dump_reference_golden.py --out gpurun_out/golden_probe > gpurun_out/${TAG}_ref_probe.json 2>&1
bash tools/lab/run_ab.sh > gpurun_out/${TAG}_ab.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_ref_probe.json; cat gpurun_out/${TAG}_ab.log; tail -5 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],'nohint',d['value_no_schedule_hint'],'ms',d['ms_per_step'],d['roofline']['stage_ms'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])"
