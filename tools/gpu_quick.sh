#!/bin/bash
# quick GPU pass: all GPU tests, then the headline bench without siblings (usage: tools/gpu_quick.sh <tag>)
TAG=${1:-q}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-siblings > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -4 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],'nohint',d['value_no_schedule_hint'],'ms',d['ms_per_step'],d['roofline']['stage_ms'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac']); print(d['roofline']['trace_ms'])"
