#!/bin/bash
# LAB: A/B of the ADMM kernel alone (old = build of the previous commit, new = working tree) on the same box
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/ab_gpu.txt 2>&1
for v in old new; do
  bin=tools/lab/admm_lab; [ $v = old ] && bin=tools/lab/admm_lab_old
  [ -x $bin ] || continue
  timeout 300 $bin tools/lab/records.bin 4096 16 > gpurun_out/ab_admm_$v.txt 2>&1
done
tail -n +1 gpurun_out/ab_admm_*.txt
