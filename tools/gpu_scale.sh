#!/bin/bash
# weak-scaling bench on N GPUs of one box (torchrun, one rank per GPU, NCCL only for the barrier / max-reduction)
n=${1:-2}; tag=${2:-scale}
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/${tag}_n1.json 2> gpurun_out/${tag}_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_n$n.json 2> gpurun_out/${tag}_n$n.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 3 --warmup 1 --impl reference > gpurun_out/${tag}_ref_n$n.json 2> gpurun_out/${tag}_ref_n$n.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d.get("n_gpus"), round(d["value"]/1e6,3), "Mcyc/s", d["ms_per_step"], "e2e", round(d["e2e"]["value"]/1e6,3))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/${tag}_n$n.err
