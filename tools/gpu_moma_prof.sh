#!/bin/bash
# GPU-box pass for the whole-body (mobile manipulator) build kernel: parity tests, default bench line, ncu --set full of k_robot_job
# in the Husky-FR3 QPIK cycle (source-level CSV for tools/ncu_lines.py).   usage: tools/gpu_moma_prof.sh <tag> [skip_tests]
tag=${1:-r02m}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
if [ -z "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
  timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
fi
for wl in husky_qpik xls_qpik; do
  timeout 300 python bench.py --workload $wl --batch 65536 --steps 5 --warmup 3 > gpurun_out/${tag}_${wl}_b65536.json 2>> gpurun_out/${tag}_bench.err
done
# launches of one step: F_STORE job, collision kernels, the K_IK build job, k_admm ...
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_husky_launches.csv \
  python bench.py --workload husky_qpik --batch 65536 --steps 1 --warmup 3 > gpurun_out/${tag}_ncu_husky.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_robot_job" --launch-skip 3 -c 1 -f -o gpurun_out/${tag}_moma_job \
  python bench.py --workload husky_qpik --batch 65536 --steps 1 --warmup 3 > gpurun_out/${tag}_ncu_moma_job.log 2>&1
ncu -i gpurun_out/${tag}_moma_job.ncu-rep --page raw --csv > gpurun_out/${tag}_moma_job_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_moma_job.ncu-rep --page source --csv > gpurun_out/${tag}_moma_job_src.csv 2>/dev/null
ncu -i gpurun_out/${tag}_moma_job.ncu-rep --page details > gpurun_out/${tag}_moma_job_details.txt 2>/dev/null
rm -f gpurun_out/${tag}_moma_job.ncu-rep
du -sh gpurun_out
tail -3 gpurun_out/${tag}_pytest.log 2>/dev/null; head -c 1500 gpurun_out/${tag}_bench.json 2>/dev/null; echo; cat gpurun_out/${tag}_husky_qpik_b65536.json | head -c 600
