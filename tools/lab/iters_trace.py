#!/usr/bin/env python3
"""LAB: ADMM iteration counts of the benchmark batch over consecutive control ticks (same ticks as bench.py) -> npz,
plus the library's stage trace of every tick.  Used to study how well the previous tick predicts the next."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import dyros_robot_controller_b200 as drc
from bench import LINK, make_workload

B, T = 65536, int(sys.argv[1]) if len(sys.argv) > 1 else 24
model = drc.Model(drc.FR3_URDF, drc.FR3_SRDF)
ctx = drc.Context(model, B)
q, qd, q_t, xd = make_workload(model, B, 0)
ctx.update_state(q_t, qd)
x_t = ctx.get_frame(LINK, want=("pose",))["pose"]
dev = torch.device("cuda", 0)
tq, tqd, txt, txd = (torch.from_numpy(a).to(dev) for a in (q, qd, x_t, xd))
ctx.enable_timing(True)
its, sts, traces = [], [], []
for k in range(T):
    r = ctx.cycle_qpik_step(tq + (k * 1e-3) * tqd, tqd, txt, txd, LINK)
    torch.cuda.synchronize()
    its.append(r["iters"].cpu().numpy().copy()); sts.append(r["status"].cpu().numpy().copy())
    traces.append(ctx.last_trace())
np.savez_compressed("gpurun_out/iters_trace.npz", iters=np.stack(its), status=np.stack(sts))
json.dump(traces, open("gpurun_out/iters_trace_stages.json", "w"))
for t in traces[-6:]:
    d = dict(t)
    print({k: round(d[k], 3) for k in ("prio_collision", "prio_build", "prio_admm", "collision", "build", "epa_join", "admm", "end")})
