#!/usr/bin/env python3
"""Generate tests/golden/fr3_golden.npz -- committed input/output vectors of the FR3 hot path.

The reference (YoungWook0533/dyros_robot_controller) cannot be built or imported in this image
(needs catkin, Pinocchio, hpp-fcl, OSQP, OsqpEigen, eigenpy: SURVEY.md section 8c) and holds no tests
or golden vectors of its own, so these vectors come from the CPU oracle (oracle/src, C++) AFTER it
has been cross-checked against the independent numpy restatement (oracle/np_oracle.py) and the
URDF-derived analytic anchors (tests/test_oracle_anchors.py).  PARITY UNPINNED by the reference.

    python tools/make_golden.py        # rewrites tests/golden/fr3_golden.npz
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle.c_oracle import Oracle  # noqa: E402
from tests.conftest import LINK, SRDF, URDF, workload  # noqa: E402


def main():
    o = Oracle(URDF, SRDF, threads=1)
    f = o.frame_id(LINK)
    B = 24
    q, qd, q_t, xdot_t = workload(o.model, B, 2024, stress=True)
    # anchors first: q = 0 and the home posture (fr3_controller.cpp:94)
    q[0] = 0.0
    q[1] = [0, 0, 0, -np.pi / 2, 0, np.pi / 2, np.pi / 4]
    q_t[:2] = q[:2] + 0.03
    st = o.update_state(q, qd, f)
    x_t = o.update_state(q_t, qd, f)["pose"]
    mani, mgrad, mgraddot = o.manipulability(q, qd, f, with_graddot=True)
    md = o.min_distance(q, qd, with_graddot=True)
    out = dict(q=q, qd=qd, x_target=x_t, xdot_target=xdot_t, pose=st["pose"], J=st["J"], Jdot=st["Jdot"], M=st["M"],
               Minv=st["Minv"], g=st["g"], nle=st["nle"], mani=mani, mani_grad=mgrad, mani_graddot=mgraddot,
               dist=md["d"], dist_grad=md["grad"], dist_graddot=md["grad_dot"], dist_pair=md["pair"])
    for mode, name in ((1, "qpik_step"), (3, "qpid_step")):
        r = o.cycle(mode, q, qd, x_t, xdot_t, f, want_x=True)
        out[name + "_out"] = r["out"]; out[name + "_status"] = r["status"]; out[name + "_iters"] = r["iters"]
        out[name + "_x"] = r["x"]
    des = 0.2 * np.random.default_rng(5).normal(size=(B, 6))
    r = o.cycle(0, q, qd, None, des, f)
    out["des"] = des; out["qpik_out"] = r["out"]; out["qpik_iters"] = r["iters"]; out["qpik_status"] = r["status"]
    null = np.random.default_rng(6).normal(size=(B, 7))
    out["null"] = null
    out["clik"] = o.taskspace(0, q, qd, x_t, xdot_t, f, null_vec=null)
    out["osf_step"] = o.taskspace(1, q, qd, x_t, xdot_t, f, null_vec=null)
    out["osf"] = o.taskspace(2, q, qd, None, des, f)
    out["pd_torque"] = o.joint_torque_step(q, qd, q_t, 0.5 * qd)
    dst = ROOT / "tests" / "golden" / "fr3_golden.npz"
    dst.parent.mkdir(exist_ok=True)
    np.savez_compressed(dst, **out)
    print("wrote", dst, dst.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
