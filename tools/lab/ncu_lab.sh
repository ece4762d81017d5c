#!/bin/bash
# LAB: ncu --set full of k_admm alone (first launch of admm_lab = the benchmark records in input order)
mkdir -p gpurun_out
tag=${1:-lab}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_admm --launch-skip 1 -c 1 -f -o gpurun_out/${tag}_k_admm \
  tools/lab/admm_lab tools/lab/records.bin 4096 16 > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}_k_admm.ncu-rep --page raw --csv > gpurun_out/${tag}_k_admm_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_k_admm.ncu-rep --page source --csv > gpurun_out/${tag}_k_admm_src.csv 2>/dev/null
ncu -i gpurun_out/${tag}_k_admm.ncu-rep --page details > gpurun_out/${tag}_k_admm_details.txt 2>/dev/null
rm -f gpurun_out/${tag}_k_admm.ncu-rep
tail -3 gpurun_out/${tag}_ncu.log; ls -la gpurun_out/${tag}_*
