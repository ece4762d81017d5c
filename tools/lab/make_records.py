"""Lab input: FR3 QPIK records (the per-robot QP data k_admm reads) for the benchmark workload, built on the CPU by the
kernel-body emulation (tests/kernel_emu).  Writes tools/lab/records.bin (float64, N x STRIDE) + the emulation's iteration
counts / solutions as the reference for variant checks."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from tests.conftest import URDF, SRDF, workload  # noqa: E402
from tests.emu import Emu  # noqa: E402
from oracle.c_oracle import Oracle  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
e = Emu(URDF, SRDF)
o = Oracle(URDF, SRDF, threads=8)
f = o.frame_id("fr3_link8")
q, qd, q_t, xd = workload(o.model, N, 0)
x_t = o.update_state(q_t, qd, f)["pose"]
r = e.cycle(1, q, qd, x_t, xd, e.frame_id("fr3_link8"), want_records=True)
out = Path(__file__).resolve().parent
r["records"].astype(np.float64).tofile(out / "records.bin")
np.savez(out / "records_ref.npz", iters=r["iters"], status=r["status"], out=r["out"])
print("records", r["records"].shape, "mean iters", r["iters"].mean(), "max", r["iters"].max(), "solved", (r["status"] == 1).mean())
