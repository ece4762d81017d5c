#!/usr/bin/env python3
"""LAB: stage times (CUDA events inside the library) of the fused cycle vs update_state + QPIKStep-from-cache."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import dyros_robot_controller_b200 as drc
from bench import LINK, make_workload

B = 65536
model = drc.Model(drc.FR3_URDF, drc.FR3_SRDF)
ctx = drc.Context(model, B)
q, qd, q_t, xd = make_workload(model, B, 0)
ctx.update_state(q_t, qd)
x_t = ctx.get_frame(LINK, want=("pose",))["pose"]
dev = torch.device("cuda", 0)
tq, tqd, txt, txd = [torch.from_numpy(a).to(dev) for a in (q, qd, x_t, xd)]
ctx.enable_timing(True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush2 = torch.ones(256 << 20, dtype=torch.uint8, device=dev)
for hint, fl in ((1, "write"), (1, "write+read"), (1, "none"), (0, "write"), (0, "write+read"), (0, "none")):
    ctx.set_params(schedule_hint=hint)
    for mode in ("fused",):
        acc = np.zeros(4); n = 0
        for it in range(8):
            if fl != "none":
                flush.zero_()
            if fl == "write+read":
                sink = flush2.view(torch.int64).sum()
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            if mode == "fused":
                r = ctx.cycle_qpik_step(tq, tqd, txt, txd, LINK)
            else:
                ctx.update_state(tq, tqd)
                e1.record()
                r = ctx.qpik_step(txt, txd, LINK)
            e2.record()
            torch.cuda.synchronize()
            if it >= 3:
                t_ = ctx.last_timing(); acc += np.array([t_["collision_ms"], t_["build_ms"], t_["admm_ms"], t_["total_ms"]]); n += 1
                tot = e0.elapsed_time(e2); upd = e0.elapsed_time(e1) if mode == "split" else 0.0
        print(f"hint={hint} flush={fl} {mode}: stages [col, build, admm, total] = {np.round(acc / n, 3)}  outer total {tot:.3f} ms  update_state {upd:.3f} ms")
