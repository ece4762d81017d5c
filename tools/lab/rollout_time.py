#!/usr/bin/env python3
"""LAB: time per tick of drc_batch_rollout_qpik at the benchmark batch (device tensors, CUDA events) next to the same ticks as
T fused cycle calls + an integrate step on the caller's side."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import dyros_robot_controller_b200 as drc
from bench import LINK, make_workload

B, T = 65536, 20
model = drc.Model(drc.FR3_URDF, drc.FR3_SRDF)
ctx = drc.Context(model, B)
q, qd, q_t, xd = make_workload(model, B, 0)
ctx.update_state(q_t, qd)
x_t = ctx.get_frame(LINK, want=("pose",))["pose"]
dev = torch.device("cuda", 0)
for fused, warm in ((0, 0), (1, 0), (0, 1)):
    ctx.set_params(rollout_fused=fused, rollout_warm_start=warm)
    for rep in range(3):
        tq, tqd, txt, txd = (torch.from_numpy(a).to(dev) for a in (q, qd, x_t, xd))
        torch.cuda.synchronize()
        l0 = ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = ctx.rollout_qpik(tq, tqd, txt, txd, LINK, T, 1e-3)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"rollout [{'two launches per tick' if fused else 'pipeline per tick'}{', warm start' if warm else ''}] B={B} T={T}: {ms:.2f} ms total, "
              f"{ms / T:.3f} ms / tick, {B * T / ms / 1e3:.2f} M cycles/s, {(ctx.launch_count - l0) / T:.2f} launches / tick, "
              f"failed ticks {int(r['fail_ticks'].sum())}, mean iterations / tick {float(r['iters_total'].double().mean()) / T:.1f}")
ctx.set_params(rollout_fused=0, rollout_warm_start=0)
for rep in range(2):
    tq, tqd, txt, txd = (torch.from_numpy(a).to(dev) for a in (q, qd, x_t, xd))
    torch.cuda.synchronize()
    l0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(T):
        r = ctx.cycle_qpik_step(tq, tqd, txt, txd, LINK)
        tq = tq + 1e-3 * r["out"]
        tqd = r["out"]
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"T fused cycle calls + torch integrate: {ms:.2f} ms total, {ms / T:.3f} ms / tick, {(ctx.launch_count - l0) / T:.2f} library launches / tick")
