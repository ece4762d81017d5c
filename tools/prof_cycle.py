#!/usr/bin/env python3
"""Profiling driver: a few FR3 QPIK control-cycle launches at the benchmark batch (for ncu)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import dyros_robot_controller_b200 as drc
from bench import LINK, make_workload

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
model = drc.Model(drc.FR3_URDF, drc.FR3_SRDF)
ctx = drc.Context(model, B)
q, qd, q_t, xd = make_workload(model, B, 0)
ctx.update_state(q_t, qd)
x_t = ctx.get_frame(LINK, want=("pose",))["pose"]
dev = torch.device("cuda", 0)
t = [torch.from_numpy(a).to(dev) for a in (q, qd, x_t, xd)]
for _ in range(reps):
    r = ctx.cycle_qpik_step(*t, LINK)
torch.cuda.synchronize()
print("ok", float(r["out"].abs().sum()), int((r["status"] == 1).sum()))
