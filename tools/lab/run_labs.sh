#!/bin/bash
# LAB: all lab measurements into gpurun_out/ (copied to profiles/ by hand)
mkdir -p gpurun_out
tools/lab/admm_lab tools/lab/records.bin 4096 16 > gpurun_out/lab_admm_kernel.txt 2>&1
{ tools/lab/lat_lab; tools/lab/lat_lab2; } > gpurun_out/lab_latencies.txt 2>&1
{ python tools/lab/stage_times.py; python tools/lab/col_time.py; } > gpurun_out/lab_stage_times.txt 2>&1
tail -4 gpurun_out/lab_stage_times.txt
