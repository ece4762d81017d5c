#!/usr/bin/env python3
"""LAB: what the ADMM convergence tail costs one fused control cycle.  Steps are timed back to back like bench.py (states advance
between steps), for several iteration caps -- max_iter below 4000 CHANGES RESULTS (robots stop early); this is a measurement
of the critical path only -- and with the schedule hint on / off."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import dyros_robot_controller_b200 as drc
from bench import LINK, make_workload

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
model = drc.Model(drc.FR3_URDF, drc.FR3_SRDF)
ctx = drc.Context(model, B)
q, qd, q_t, xd = make_workload(model, B, 0)
ctx.update_state(q_t, qd)
x_t = ctx.get_frame(LINK, want=("pose",))["pose"]
dev = torch.device("cuda", 0)
tq0, tqd, txt, txd = [torch.from_numpy(a).to(dev) for a in (q, qd, x_t, xd)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for hint in (1, 0):
    for cap in (4000, 2000, 1000, 500, 250, 100):
        ctx.set_params(schedule_hint=hint, max_iter=cap)
        tq = tq0.clone()
        ev = []
        K, W = 12, 3
        for it in range(K + W):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = ctx.cycle_qpik_step(tq, tqd, txt, txd, LINK)
            e1.record()
            ev.append((e0, e1))
            tq = tq + 1e-3 * r["out"]
        torch.cuda.synchronize()
        ts = np.array([a.elapsed_time(b) for a, b in ev[W:]])
        it_ = r["iters"].cpu().numpy()
        print(f"hint={hint} max_iter={cap:5d}: step {ts.mean():.3f} ms (min {ts.min():.3f} max {ts.max():.3f})  mean iters {it_.mean():.1f}  "
              f">=300: {(it_ >= 300).sum()}  >=1000: {(it_ >= 1000).sum()}  at cap: {(it_ >= cap).sum()}", flush=True)
