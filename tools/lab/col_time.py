#!/usr/bin/env python3
"""LAB: the self-collision stage alone, in situ (device pointers, CUDA events on the launch stream)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import dyros_robot_controller_b200 as drc
from dyros_robot_controller_b200._capi import lib
from bench import make_workload

B = 65536
model = drc.Model(drc.FR3_URDF, drc.FR3_SRDF)
ctx = drc.Context(model, B)
q, qd, q_t, xd = make_workload(model, B, 0)
dev = torch.device("cuda", 0)
tq, tqd = torch.from_numpy(q).to(dev), torch.from_numpy(qd).to(dev)
d = torch.empty(B, dtype=torch.float64, device=dev)
g = torch.empty((B, 7), dtype=torch.float64, device=dev)
p = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream or 1)
for what in ("update_state", "min_distance"):
    ts = []
    for it in range(8):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if what == "update_state":
            rc = lib().drc_batch_update_state(ctx._h, B, p(tq), p(tqd), 0, st)
        else:
            rc = lib().drc_batch_get_min_distance(ctx._h, B, 0, p(d), p(g), None, None, 0, st)
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0
        ts.append(e0.elapsed_time(e1))
    print(what, "ms:", np.round(ts[3:], 3))
