#!/bin/bash
# weak-scaling bench on N GPUs of one box (torchrun, one rank per GPU, NCCL only for the barrier / max-reduction), the reference arm,
# and the STRONG-scaling run BASELINE config 5 describes (1 M XLS-FR3 robots split over the N GPUs)
n=${1:-2}; tag=${2:-scale}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
python bench.py --gpus 1 --steps 10 --warmup 3 --no-siblings > gpurun_out/${tag}_n1.json 2> gpurun_out/${tag}_n1.err
$TR --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_n$n.json 2> gpurun_out/${tag}_n$n.err
$TR --master-port 29512 bench.py --gpus $n --steps 3 --warmup 1 --impl reference > gpurun_out/${tag}_ref_n$n.json 2> gpurun_out/${tag}_ref_n$n.err
$TR --master-port 29513 bench.py --gpus $n --steps 5 --warmup 3 --workload xls_qpik --batch $((1048576 / n)) > gpurun_out/${tag}_xls_strong_n$n.json 2> gpurun_out/${tag}_xls_strong_n$n.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d.get("n_gpus"), round(d["value"]/1e6,3), "Mcyc/s", d["ms_per_step"], "e2e", round(d["e2e"]["value"]/1e6,3))
        for i, r in enumerate(d.get("per_rank") or []): print("   rank", i, r)
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/${tag}_n$n.err
