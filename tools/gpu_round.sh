#!/bin/bash
# One GPU-box pass: parity tests, both bench arms, ncu launch list, ncu --set full of the hot kernels.
# usage: tools/gpu_round.sh <tag>      (outputs under gpurun_out/<tag>_*)
tag=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 300 python tools/lab/rollout_time.py > gpurun_out/${tag}_rollout.txt 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 > gpurun_out/${tag}_ncu_bench.log 2>&1
for k in k_admm k_collision_closed k_collision k_robot_job; do
  # MAIN-pipeline launch of the second control tick (the ADMM schedule then has the previous tick's iteration counts).
  # Launch order per tick: priority pipeline (FK store, collision, build, ADMM) then main pipeline (same kernels);
  # tools/prof_cycle.py adds two k_robot_job launches for its set-up.
  skip=3; [ "$k" = "k_robot_job" ] && skip=11   # per tick: main FK, prio FK, prio build, main build, its follow-up, dynamics-only (+2 set-up launches)
  [ "$k" = "k_admm" ] && skip=5                    # per tick: priority launch, EPA-pending robots, main launch
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^${k}\$" --launch-skip $skip -c 1 -f -o gpurun_out/${tag}_${k} \
    python tools/prof_cycle.py 65536 2 > gpurun_out/${tag}_ncu_${k}.log 2>&1
  # gpurun_out/ is capped at 64 MiB: keep the CSV pages, drop the report
  ncu -i gpurun_out/${tag}_${k}.ncu-rep --page raw --csv > gpurun_out/${tag}_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${tag}_${k}.ncu-rep --page source --csv > gpurun_out/${tag}_${k}_src.csv 2>/dev/null
  ncu -i gpurun_out/${tag}_${k}.ncu-rep --page details > gpurun_out/${tag}_${k}_details.txt 2>/dev/null
  rm -f gpurun_out/${tag}_${k}.ncu-rep
done
du -sh gpurun_out
tail -3 gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_bench.json
