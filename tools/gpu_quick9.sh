#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/${tag}_ncu_bench.log 2>&1
python - <<PY
import csv, collections
rows=list(csv.reader(open("gpurun_out/${tag}_launches.csv")))
for i,r in enumerate(rows):
    if r and r[0]=='ID': hdr=r; data=rows[i+1:]; break
ix={k:i for i,k in enumerate(hdr)}
agg=collections.defaultdict(list)
for r in data:
    if len(r)<len(hdr): continue
    name=r[ix['Kernel Name']][:90]+' grid='+r[ix['Grid Size']]
    agg[name].append(float(r[ix['Metric Value']])/1e6)
for k,v in agg.items(): print(f"{k:120s} n={len(v):3d} mean={sum(v)/len(v):8.3f} ms  last={v[-1]:.3f}")
PY
