#!/bin/bash
# round-2 GPU pass A: reference-library probe, ADMM lab A/B, parity tests, bench
mkdir -p gpurun_out
python tools/dump_reference_golden.py --out gpurun_out/golden_probe > gpurun_out/r2a_ref_probe.json 2>&1
bash tools/lab/run_ab.sh > gpurun_out/r2a_ab.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
cat gpurun_out/r2a_ref_probe.json; cat gpurun_out/r2a_ab.log; tail -5 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],'nohint',d['value_no_schedule_hint'],'ms',d['ms_per_step'],d['roofline']['stage_ms'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])"
