#!/bin/bash
# development pass on a --dev build (7-dof only): FR3 parity tests, narrow-phase timing, headline bench without siblings
TAG=${1:-dev}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rollout.py -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
python tools/lab/col_time.py 2>&1 | tee gpurun_out/${TAG}_col.txt
timeout 600 python bench.py --steps 30 --warmup 3 --no-siblings > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1]); print('value',d['value'],'nohint',d['value_no_schedule_hint'],'ms',d['ms_per_step'],d['roofline']['stage_ms'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac']); print(d['roofline']['trace_ms'])"
