#!/bin/bash
# MoMa parity tests + whole-body benches on the full build, then the FR3 lab variants (tools/lab/variants/*.so) on the same box
tag=${1:-p3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_moma.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
for wl in husky_qpik husky_qpid xls_qpik; do
  b=262144; [ $wl = xls_qpik ] && b=1048576
  timeout 300 python bench.py --workload $wl --batch $b --steps 5 --warmup 3 2>> gpurun_out/${tag}_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', d['config']['batch_per_gpu'], 'value %.4g ms %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['roofline'].get('stage_ms'))"
done
bash tools/lab/run_so_variants.sh 2>&1 | grep -v min_distance
