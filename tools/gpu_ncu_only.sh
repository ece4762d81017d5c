#!/bin/bash
tag=${1:-r01}
mkdir -p gpurun_out
for k in k_admm k_collision k_robot_job; do
  skip=3; [ "$k" = "k_robot_job" ] && skip=9
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^${k}\$" --launch-skip $skip -c 1 -f -o gpurun_out/${tag}_${k} \
    python tools/prof_cycle.py 65536 2 > gpurun_out/${tag}_ncu_${k}.log 2>&1
  ncu -i gpurun_out/${tag}_${k}.ncu-rep --page raw --csv > gpurun_out/${tag}_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${tag}_${k}.ncu-rep --page details > gpurun_out/${tag}_${k}_details.txt 2>/dev/null
  rm -f gpurun_out/${tag}_${k}.ncu-rep
done
grep -h "gpu__time_duration.sum\|Duration" gpurun_out/${tag}_k_*_details.txt | head
