#!/bin/bash
# quick GPU-box pass: all parity tests + the default bench line (with siblings)   usage: tools/gpu_pass2.sh <tag>
tag=${1:-p}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print('value %.4g nohint %.4g ms %.3f e2e %.4g frac %.3f' % (d['value'], d['value_no_schedule_hint'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']), d['roofline']['stage_ms'])
print(d['roofline'].get('trace_ms'))
for k,v in d.get('siblings',{}).items():
    print(k, v['batch'], '%.4g' % v['value'], 'ms %.3f' % v['ms_per_step'], 'e2e %.4g' % v['e2e']['value'], v['roofline'].get('stage_ms'))
PY
