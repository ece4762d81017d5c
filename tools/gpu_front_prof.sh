#!/bin/bash
# ncu --set full of the main-pipeline front-stage kernels of the second control tick, with per-source-line pages
tag=${1:-front}
mkdir -p gpurun_out
for k in k_collision k_robot_job; do
  skip=3; [ "$k" = "k_robot_job" ] && skip=10
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^${k}\$" --launch-skip $skip -c 1 -f -o gpurun_out/${tag}_${k} \
    python tools/prof_cycle.py 65536 2 > gpurun_out/${tag}_ncu_${k}.log 2>&1
  ncu -i gpurun_out/${tag}_${k}.ncu-rep --page raw --csv > gpurun_out/${tag}_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${tag}_${k}.ncu-rep --page source --csv > gpurun_out/${tag}_${k}_src.csv 2>/dev/null
  ncu -i gpurun_out/${tag}_${k}.ncu-rep --page details > gpurun_out/${tag}_${k}_details.txt 2>/dev/null
  rm -f gpurun_out/${tag}_${k}.ncu-rep
done
ls -la gpurun_out/${tag}_*; true
