#!/usr/bin/env python3
"""Mobile base kernels (k_mobile_fk / k_mobile_ik) at a large batch: bases/s and the HBM roofline fraction.
Algorithmic bytes per base: FK  8*(w + w) in [caster; differential / mecanum read the wheel velocities only: 8*w] + 8*3 out,
IK 8*3 in (+ 8*w steering angles for casters) + 8*w out; Jacobian outputs are not requested (want_J=False)."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import dyros_robot_controller_b200 as drc
from tests.test_mobile_cpu import KINS, wheels_of

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
peak = 6562.9
try:
    peak = float(json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs", peak))
except Exception:
    pass
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, kin in KINS.items():
    w = wheels_of(kin)
    base = drc.MobileBase(kin)
    caster = kin["type"] == "Caster"
    wp = torch.rand((B, w), dtype=torch.float64, device=dev) * 6.28 - 3.14
    wv = torch.rand((B, w), dtype=torch.float64, device=dev) * 4 - 2
    bv = torch.randn((B, 3), dtype=torch.float64, device=dev)
    for what in ("fk", "ik"):
        ts = []
        for it in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if what == "fk":
                base.fk(wp, wv, want_J=False)
            else:
                base.ik(wp, bv, saturate=True, want_J=False)
            e1.record(); e1.synchronize()
            if it >= 3:
                ts.append(e0.elapsed_time(e1))
        ms = float(np.mean(ts))
        nbytes = 8 * ((w if what == "fk" else 3) + (w if caster else 0) + (3 if what == "fk" else w))
        print(json.dumps({"kernel": f"k_mobile_{what}", "drive": name, "wheels": w, "batch": B, "ms": round(ms, 4),
                          "bases_per_s": B / ms * 1e3, "algorithmic_bytes_per_base": nbytes,
                          "roofline": {"bound": "hbm", "achieved": B * nbytes / ms / 1e6, "peak": peak, "unit": "GB/s",
                                       "frac": B * nbytes / ms / 1e6 / peak}}))
