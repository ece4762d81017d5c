#!/bin/bash
# LAB: same box, several builds of libdrc_b200.so (tools/lab/variants/*.so): FR3 QPID bench (config 3) of each, twice
mkdir -p gpurun_out
LIB=dyros_robot_controller_b200/libdrc_b200.so
cp $LIB /tmp/lib_orig.so
for v in tools/lab/variants/*.so; do
  n=$(basename $v .so)
  cp $v $LIB
  echo "== $n"
  for rep in 1 2; do
  timeout 600 python bench.py --workload fr3_qpid --steps 10 --warmup 3 --no-siblings 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.4g nohint %.4g ms %.3f' % (d['value'],d['value_no_schedule_hint'],d['ms_per_step']), d['roofline']['stage_ms'], 'alone', d['roofline']['alone'])"
  done
done 2>&1 | tee gpurun_out/so_variants_qpid.txt
cp /tmp/lib_orig.so $LIB
