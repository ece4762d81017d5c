#!/bin/bash
# LAB: launch-bound variants of the ADMM kernel alone (same records, same box)
mkdir -p gpurun_out
for b in admm_lab admm_lab_m4w1 admm_lab_m2w1 admm_lab_m3w2; do
  [ -x tools/lab/$b ] || continue
  echo "== $b"; timeout 300 tools/lab/$b tools/lab/records.bin 4096 16 2>&1 | head -4
done > gpurun_out/lab_variants.txt 2>&1
cat gpurun_out/lab_variants.txt
