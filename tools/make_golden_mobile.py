#!/usr/bin/env python3
"""Generate tests/golden/mobile_golden.npz -- committed input/output vectors of the mobile base path
(Mobile::RobotData / Mobile::RobotController for the differential, mecanum and powered-caster drives) and of the
powered-caster mobile manipulator state.  Source: the CPU oracle (oracle/src/omoma.h), after it has been checked against the
reference's closed forms recomputed in numpy (tests/test_mobile_cpu.py).  PARITY UNPINNED by the reference (it holds no
tests or golden vectors and cannot be built here).

    python tools/make_golden_mobile.py        # rewrites tests/golden/mobile_golden.npz
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle.c_oracle import MomaOracle, mobile_base  # noqa: E402
from tests.conftest import MOMA, moma_workload  # noqa: E402
from tests.test_mobile_cpu import KINS, wheels_of  # noqa: E402


def main():
    out = {}
    rng = np.random.default_rng(2025)
    for name, kin in KINS.items():
        w = wheels_of(kin)
        B = 12
        wp = rng.uniform(-np.pi, np.pi, (B, w)); wv = rng.uniform(-3, 3, (B, w))
        bv = rng.normal(size=(B, 3)) * np.array([1.5, 1.5, 3.0])
        bv[0] = [1e-5, -2e-5, 0.3]; bv[1] = [10.0, 0.0, -9.0]
        J, vel = mobile_base(kin, True, wp, wv)
        Ji, wheel = mobile_base(kin, False, wp, bv)
        _, wheel_sat = mobile_base(kin, False, wp, bv, saturate=True)
        for k, v in dict(wheel_pos=wp, wheel_vel=wv, base_vel_des=bv, J_fk=J, base_vel=vel, J_ik=Ji, wheel_cmd=wheel,
                         wheel_cmd_saturated=wheel_sat).items():
            out[f"{name}_{k}"] = v
    d = MOMA["pcv_fr3"]
    o = MomaOracle(d["urdf"], d["srdf"], d["kin"], d["joint_idx"], d["actuator_idx"], threads=1)
    f = o.frame_id("fr3_link8")
    q, qd, q_t, xd = moma_workload(o.model, o.w, 10, 2026)
    ms = o.moma_update_state(q, qd, f)
    x_t = o.update_state(q_t, qd, f)["pose"]
    r = o.moma_cycle(1, q, qd, x_t, xd, f)
    out.update(pcv_q=q, pcv_qd=qd, pcv_x_target=x_t, pcv_xdot_target=xd, pcv_M=ms["M"], pcv_g=ms["g"], pcv_J=ms["J"], pcv_Jdot=ms["Jdot"],
               pcv_mani=ms["mani"], pcv_qpik_step_out=r["out"], pcv_qpik_step_status=r["status"], pcv_qpik_step_iters=r["iters"])
    dst = ROOT / "tests" / "golden" / "mobile_golden.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, dst.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
