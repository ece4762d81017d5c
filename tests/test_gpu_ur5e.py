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


@pytest.mark.parametrize("robot", ["ur5e", "fr3"])
def test_fused_clik_osf_cycle_equals_the_three_calls(robot):
    """drc_{host,batch}_cycle_clik_osf_step = updateState + CLIKStep + OSFStep in one launch: the same commands as the three
    separate calls (CLIK forms Kp e + xdot_target, OSF forms Kp e + Kv edot, robot_controller.cpp:169,238), the same state cache
    afterwards, host and device paths, and the oracle's values."""
    import torch
    import dyros_robot_controller_b200 as drc
    from oracle.c_oracle import Oracle
    urdf, srdf, link = (URDF, SRDF, LINK) if robot == "ur5e" else (drc.FR3_URDF, drc.FR3_SRDF, "fr3_link8")
    o = Oracle(urdf, srdf, threads=8)
    model = drc.Model(urdf, srdf)
    ctx, ctx2 = drc.Context(model, 4096, device=0), drc.Context(model, 4096, device=0)
    B = 4096
    q, qd, q_t, xdot_t = workload(o.model, B, 72)
    f = o.frame_id(link)
    x_t = o.update_state(q_t, qd, f)["pose"]
    ctx.update_state(q, qd)
    a_clik, a_osf = ctx.clik_step(x_t, xdot_t, link), ctx.osf_step(x_t, xdot_t, link)
    r = ctx2.cycle_clik_osf_step(q, qd, x_t, xdot_t, link)
    # two instantiations of the same arithmetic (FMA contraction may differ): rounding level, amplified by the conditioning of
    # J Minv J' for the torque (measured: 5e-10 on |tau| up to 230)
    rowrel = lambda a, b: (np.abs(a - b).max(axis=1) / np.maximum(1.0, np.abs(b).max(axis=1))).max()
    assert rowrel(r["qdot"], a_clik) < 1e-10
    assert rowrel(r["tau"], a_osf) < 1e-8
    d1, d2 = ctx.get_dynamics(), ctx2.get_dynamics()
    close = lambda a, b: np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(b).max())   # two instantiations of the same arithmetic
    for k in ("M", "Minv", "g", "nle"):
        assert close(d1[k], d2[k]), k
    f1, f2 = ctx.get_frame(link), ctx2.get_frame(link)
    assert close(f1["pose"], f2["pose"]) and close(f1["J"], f2["J"])
    # device path
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(a).to(dev) for a in (q, qd, x_t, xdot_t)]
    rt = ctx2.cycle_clik_osf_step(*t, link)
    torch.cuda.synchronize()
    assert np.array_equal(rt["qdot"].cpu().numpy(), r["qdot"]) and np.array_equal(rt["tau"].cpu().numpy(), r["tau"])
    # oracle
    for got, mode in ((r["qdot"], 0), (r["tau"], 1)):
        b = o.taskspace(mode, q, qd, x_t, xdot_t, f)
        err = np.abs(got - b).max(axis=1) / np.maximum(1.0, np.abs(b).max(axis=1))
        assert (err < 1e-7).mean() > 0.99 and err.max() < 1e-3
