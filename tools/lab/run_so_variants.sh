#!/bin/bash
# LAB: same box, several builds of libdrc_b200.so (tools/lab/variants/*.so): narrow-phase timing + headline bench of each
mkdir -p gpurun_out
LIB=dyros_robot_controller_b200/libdrc_b200.so
cp $LIB /tmp/lib_orig.so
for v in tools/lab/variants/*.so; do
  n=$(basename $v .so)
  cp $v $LIB
  echo "== $n"
  python tools/lab/col_time.py 2>&1 | tail -1
  for rep in 1 2; do
  timeout 600 python bench.py --steps 30 --warmup 3 --no-siblings 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.4g nohint %.4g ms %.3f' % (d['value'],d['value_no_schedule_hint'],d['ms_per_step']), d['roofline']['stage_ms'])"
  done
done 2>&1 | tee gpurun_out/so_variants.txt
cp /tmp/lib_orig.so $LIB
