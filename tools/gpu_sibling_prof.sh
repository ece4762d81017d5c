#!/bin/bash
# ncu --set full of the solver launch (k_admm) of the sibling configurations at 65 536 robots, plus the whole-body EPA pass and
# dynamics-only job of Husky-FR3.   usage: tools/gpu_sibling_prof.sh <tag>     (outputs: gpurun_out/<tag>_<workload>_<kernel>_*)
tag=${1:-sib}
mkdir -p gpurun_out
cap() {   # workload kernel-regex launch-skip count
  wl=$1; k=$2; skip=$3; cnt=$4; name=${5:-$2}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^${k}" --launch-skip $skip -c $cnt -f -o gpurun_out/${tag}_${wl}_${name} \
    python bench.py --workload $wl --batch 65536 --steps 1 --warmup 3 --no-siblings > gpurun_out/${tag}_ncu_${wl}_${name}.log 2>&1
  ncu -i gpurun_out/${tag}_${wl}_${name}.ncu-rep --page raw --csv > gpurun_out/${tag}_k_${wl}_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${tag}_${wl}_${name}.ncu-rep --page details > gpurun_out/${tag}_${wl}_${name}_details.txt 2>/dev/null
  rm -f gpurun_out/${tag}_${wl}_${name}.ncu-rep
}
# per step: fr3_qpid = priority launch, EPA-pending robots, main launch; whole-body = EPA-pending robots, main launch
cap fr3_qpid   k_admm 9 3
cap husky_qpik k_admm 6 2
cap xls_qpid   k_admm 6 2
cap husky_qpik k_collision_epa 3 1
cap husky_qpik "k_robot_job<12, 0, 4101" 3 1 k_dyn_job
du -sh gpurun_out
