"""GPU parity of BASELINE config 2: UR5e (6 dof, synthesized URDF) updateState + CLIKStep + OSFStep at batch 4096."""
import numpy as np
import pytest

from tests.conftest import ROBOTS, workload

pytestmark = pytest.mark.gpu
URDF, SRDF, LINK = str(ROBOTS / "ur5e" / "ur5e.urdf"), str(ROBOTS / "ur5e" / "ur5e.srdf"), "tool0"


def test_ur5e_clik_osf_batch_4096():
    import dyros_robot_controller_b200 as drc
    from oracle.c_oracle import Oracle
    o = Oracle(URDF, SRDF, threads=8)
    model = drc.Model(URDF, SRDF)
    assert model.dof == 6
    ctx = drc.Context(model, 4096, device=0)
    B = 4096
    q, qd, q_t, xdot_t = workload(o.model, B, 71)
    f = o.frame_id(LINK)
    ref = o.update_state(q, qd, f)
    x_t = o.update_state(q_t, qd, f)["pose"]
    ctx.update_state(q, qd)
    fr, dy = ctx.get_frame(LINK), ctx.get_dynamics()
    rel = lambda a, b: np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
    assert rel(fr["pose"], ref["pose"]) < 1e-12 and rel(fr["J"], ref["J"]) < 1e-12 and rel(fr["Jdot"], ref["Jdot"]) < 1e-11
    assert rel(dy["M"], ref["M"]) < 1e-9 and rel(dy["g"], ref["g"]) < 1e-9 and rel(dy["nle"], ref["nle"]) < 1e-9
    # away from wrist singularities the two pseudo-inverse implementations agree to rounding; near them (sigma_min of the
    # 6x6 Jacobian small) the error scales with 1/sigma_min: compare relative to each robot's output norm
    a, b = ctx.clik_step(x_t, xdot_t, LINK), o.taskspace(0, q, qd, x_t, xdot_t, f)
    err = np.abs(a - b).max(axis=1) / np.maximum(1.0, np.abs(b).max(axis=1))
    assert (err < 1e-7).mean() > 0.99 and err.max() < 1e-3
    a, b = ctx.osf_step(x_t, xdot_t, LINK), o.taskspace(1, q, qd, x_t, xdot_t, f)
    err = np.abs(a - b).max(axis=1) / np.maximum(1.0, np.abs(b).max(axis=1))
    assert (err < 1e-7).mean() > 0.99 and err.max() < 1e-3
    # the 6-dof QPIK (20 variables / 34 rows) runs too
    r = ctx.cycle_qpik_step(q[:512], qd[:512], x_t[:512], xdot_t[:512], LINK)
    ref_c = o.cycle(1, q[:512], qd[:512], x_t[:512], xdot_t[:512], f)
    same = (r["iters"] == ref_c["iters"]) & (r["status"] == ref_c["status"])
    assert same.mean() > 0.95
    assert (np.abs(r["out"] - ref_c["out"]).max(axis=1)[same] < 1e-3).mean() > 0.99
