#!/bin/bash
# GPU-box pass for the whole-body (mobile manipulator) path: its parity tests, the sibling benches at their BASELINE batches, the launch
# list of one Husky-FR3 / XLS-FR3 QPIK step and ncu --set full of the kernels that carry it.   usage: tools/gpu_moma_prof.sh <tag>
tag=${1:-r02m}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_moma.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
for wl in husky_qpik husky_qpid xls_qpik xls_qpid; do
  b=262144; case $wl in xls*) b=1048576;; esac
  timeout 300 python bench.py --workload $wl --batch $b --steps 5 --warmup 3 > gpurun_out/${tag}_${wl}.json 2>> gpurun_out/${tag}_bench.err
  python -c "
import json
d=json.loads(open('gpurun_out/${tag}_${wl}.json').read().strip().splitlines()[-1]); print('$wl', d['config']['batch_per_gpu'], 'value %.4g ms %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['roofline'].get('stage_ms'))"
done
for wl in husky_qpik xls_qpik; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_${wl}_launches.csv \
    python bench.py --workload $wl --batch 65536 --steps 1 --warmup 3 > gpurun_out/${tag}_ncu_${wl}.log 2>&1
done
# XLS-FR3 step at 65536: the QP-build job (k_robot_job, flags 4132), the dynamics-only job (4101) and k_pinv_list
for k in k_pinv_list k_robot_job; do
  skip=3; [ "$k" = "k_robot_job" ] && skip=7   # set-up: 2 state launches; per step: F_STORE, build, build redo, dynamics
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^${k}" --launch-skip $skip -c 3 -f -o gpurun_out/${tag}_${k} \
    python bench.py --workload xls_qpik --batch 65536 --steps 1 --warmup 3 > gpurun_out/${tag}_ncu_${k}.log 2>&1
  ncu -i gpurun_out/${tag}_${k}.ncu-rep --page raw --csv > gpurun_out/${tag}_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${tag}_${k}.ncu-rep --page details > gpurun_out/${tag}_${k}_details.txt 2>/dev/null
  rm -f gpurun_out/${tag}_${k}.ncu-rep
done
du -sh gpurun_out
