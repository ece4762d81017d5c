"""GPU parity tests of the mobile-manipulator path through the C ABI against the oracle."""
import numpy as np
import pytest

from tests.conftest import MOMA, moma_workload

pytestmark = pytest.mark.gpu
LINK = "fr3_link8"


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module", params=["husky_fr3", "xls_fr3"])
def rig(request):
    import dyros_robot_controller_b200 as drc
    from oracle.c_oracle import MomaOracle
    if drc.device_count() < 1:
        pytest.fail("GPU tests need a CUDA device; the product path has no CPU fallback")
    d = MOMA[request.param]
    o = MomaOracle(d["urdf"], d["srdf"], d["kin"], d["joint_idx"], d["actuator_idx"], threads=8)
    model = drc.Model(d["urdf"], d["srdf"]).attach_mobile_base(d["kin"], d["joint_idx"], d["actuator_idx"])
    ctx = drc.Context(model, 8192, device=0)
    return request.param, d, o, model, ctx


def consistent(o, q, qd):
    J, bv = o.mobile_state(q[:, 3:3 + o.w], qd[:, 3:3 + o.w])
    c, s = np.cos(q[:, 2]), np.sin(q[:, 2])
    qd = qd.copy()
    qd[:, 0], qd[:, 1], qd[:, 2] = c * bv[:, 0] - s * bv[:, 1], s * bv[:, 0] + c * bv[:, 1], bv[:, 2]
    return qd


def test_moma_model_and_state(rig):
    name, d, o, model, ctx = rig
    assert model.actuated_dof == o.act and model.mani_dof == 7 and model.wheel_num == o.w
    assert np.abs(model.base_jacobian() - o.mobile_state(np.zeros((1, o.w)), np.zeros((1, o.w)))[0][0]).max() < 1e-12
    f = o.frame_id(LINK)
    q, qd, _, _ = moma_workload(o.model, o.w, 1000, 31)
    ref, full = o.moma_update_state(q, qd, f), o.update_state(q, qd, f)
    ctx.moma_update_state(q, qd)
    r = ctx.moma_get_state(LINK)
    assert rel(r["pose"], full["pose"]) < 1e-12
    assert rel(r["J"], ref["J"]) < 1e-12 and rel(r["Jdot"], ref["Jdot"]) < 1e-11
    assert rel(r["vel"], np.einsum("bij,bj->bi", full["J"], qd)) < 1e-11
    assert rel(r["M"], ref["M"]) < 1e-9 and rel(r["g"], ref["g"]) < 1e-9 and rel(r["nle"], ref["nle"]) < 1e-9
    assert rel(r["Minv"], ref["Minv"]) < 1e-7
    assert np.abs(r["mani"] - ref["mani"]).max() < 1e-11
    assert np.abs(r["mani_grad"] - ref["mani_grad"]).max() < 1e-9 * max(1.0, np.abs(ref["mani_grad"]).max())
    assert np.abs(r["mani_graddot"] - ref["mani_graddot"]).max() < 1e-8 * max(1.0, np.abs(ref["mani_graddot"]).max())


def test_full_model_pinv_cod_against_lapack(rig):
    """PinvCOD(M) of the full model (robot_data.cpp:118, COD threshold 1e-6) as k_pinv_list computes it with 16 lanes per robot,
    against an independent statement: LAPACK's column-pivoted QR (scipy), the rank rule |R_kk| > 1e-6 max|R_kk|, numpy's pinv of the
    kept rows.  Husky-FR3 keeps every pivot (the result is the inverse), XLS-FR3 drops one (a rank-13 pseudo-inverse)."""
    import scipy.linalg as sl
    name, d, o, model, ctx = rig
    q, qd, _, _ = moma_workload(o.model, o.w, 300, 77)
    ctx.moma_update_state(q, qd)
    dyn = ctx.get_dynamics(want=("M", "Minv"))
    ranks = []
    for M, Mi in zip(dyn["M"], dyn["Minv"]):
        Q, R, P = sl.qr(M, pivoting=True)
        dg = np.abs(np.diag(R))
        r = int((dg > 1e-6 * dg.max()).sum())
        ranks.append(r)
        ref = np.zeros_like(M)
        ref[P, :] = np.linalg.pinv(R[:r, :]) @ Q[:, :r].T
        assert np.abs(Mi - ref).max() < 1e-7 * np.abs(ref).max()
        if r == M.shape[0]:   # (the truncated COD pseudo-inverse is not symmetric: dropping R_22 perturbs M unsymmetrically)
            assert np.abs(Mi - Mi.T).max() < 1e-7 * np.abs(ref).max()
    n = dyn["M"].shape[1]
    assert set(ranks) == ({n} if name == "husky_fr3" else {n - 1})


@pytest.mark.parametrize("kind", ["ik", "id"])
def test_moma_cycle_leaves_the_full_state_cache(rig, kind):
    """updateState's part of a fused cycle (mobile_manipulator/robot_data.cpp:83-144): the QP-build jobs compute only what their record
    reads and a dynamics-only job behind the solver launch completes the cache -- every getter must see the NEW state afterwards."""
    name, d, o, model, ctx = rig
    f = o.frame_id(LINK)
    q0, qd0, _, _ = moma_workload(o.model, o.w, 1500, 61)
    q, qd, q_t, xd = moma_workload(o.model, o.w, 1500, 62)
    qd = consistent(o, q, qd)
    x_t = o.update_state(q_t, qd, f)["pose"]
    ctx.moma_update_state(q0, qd0)                      # an OLD state in the cache
    ctx.moma_cycle(kind, q, qd, x_t, xd, LINK)
    ref, full = o.moma_update_state(q, qd, f), o.update_state(q, qd, f)
    # the cache as the cycle left it (these getters only copy; moma_get_state below re-evaluates the state from the cached q)
    dyn = ctx.get_dynamics(want=("M", "Minv", "g", "nle"))   # full-model quantities (Manipulator::RobotData getters)
    assert rel(dyn["M"], full["M"]) < 1e-9 and rel(dyn["g"], full["g"]) < 1e-9 and rel(dyn["nle"], full["nle"]) < 1e-9
    # PinvCOD with the reference's 1e-6 rank threshold: either the inverse (Husky-FR3, every pivot kept) or the truncated
    # pseudo-inverse (XLS-FR3); the oracle applies the same rule
    assert rel(dyn["Minv"], full["Minv"]) < 1e-6
    r = ctx.moma_get_state(LINK)
    assert rel(r["pose"], full["pose"]) < 1e-12 and rel(r["J"], ref["J"]) < 1e-12
    assert rel(r["M"], ref["M"]) < 1e-9 and rel(r["g"], ref["g"]) < 1e-9 and rel(r["nle"], ref["nle"]) < 1e-9
    assert rel(r["Minv"], ref["Minv"]) < 1e-7


@pytest.mark.parametrize("mode,B", [(1, 2000), (3, 1000), (0, 500), (2, 500)])
def test_moma_control_cycle_matches_oracle(rig, mode, B):
    name, d, o, model, ctx = rig
    f = o.frame_id(LINK)
    q, qd, q_t, xd = moma_workload(o.model, o.w, B, 40 + mode)
    qd = consistent(o, q, qd)
    x_t = o.update_state(q_t, qd, f)["pose"] if mode in (1, 3) else None
    des = xd if mode in (1, 3) else 3.0 * xd
    ref = o.moma_cycle(mode, q, qd, x_t, des, f)
    if mode == 1:
        r = ctx.moma_cycle("ik", q, qd, x_t, des, LINK)
    elif mode == 3:
        r = ctx.moma_cycle("id", q, qd, x_t, des, LINK)
    else:
        ctx.moma_update_state(q, qd)
        r = ctx.moma_qpik(des, LINK) if mode == 0 else ctx.moma_qpid(des, LINK)
    assert (r["status"] == ref["status"]).mean() > 0.98
    same = (r["iters"] == ref["iters"]) & (r["status"] == ref["status"])
    assert same.mean() > 0.97
    scale = max(1.0, np.abs(ref["out"]).max())
    err = np.abs(r["out"] - ref["out"]).max(axis=1)[same]
    # hard-constraint whole-body QPs amplify the GJK witness noise of an ACTIVE self-collision row more than the slack
    # formulations do: a handful of robots per thousand may differ by a few per cent of the torque scale
    assert (err < 1e-4 * scale).mean() > 0.99 and err.max() < (1e-2 if mode <= 1 else 5e-2) * scale
    if mode >= 2:
        e2 = np.abs(r["etadot"] - ref["out2"]).max(axis=1)[same]
        assert (e2 < 1e-4 * max(1.0, np.abs(ref["out2"]).max())).mean() > 0.99
    # fallbacks: zeros (QPIK) / actuated gravity with zero eta_dot (QPID)
    bad = r["status"] != 1
    if bad.any():
        if mode <= 1:
            assert np.abs(r["out"][bad]).max() == 0.0
        else:
            assert np.abs(r["etadot"][bad]).max() == 0.0


def test_moma_reference_api_mirror(rig):
    """dyros_robot_controller_b200.drc.mobile_manipulator mirrors the reference classes: six state vectors in,
    (mobile, manipulator) pairs out (mobile_manipulator/robot_controller.cpp:147-231)."""
    from dyros_robot_controller_b200.drc.mobile_manipulator import RobotController, RobotData
    from oracle import c_oracle
    name, d, o, model, ctx = rig
    rd = RobotData(d["kin"], d["joint_idx"], d["actuator_idx"], d["urdf"], d["srdf"], max_batch=32)
    rc = RobotController(0.001, rd)
    f = o.frame_id(LINK)
    q, qd, q_t, xd = moma_workload(o.model, o.w, 32, 50)
    qd = consistent(o, q, qd)
    w = o.w
    x_t = o.update_state(q_t, qd, f)["pose"]
    # the reference's MobileManipulator controller defaults to task gains 400 / 40 (mobile_manipulator/robot_controller.cpp:15-16)
    o.set_task_gains(np.full(6, 400.0), np.full(6, 40.0))
    try:
        ref = o.moma_cycle(1, q, qd, x_t, xd, f)
        ref3 = o.moma_cycle(3, q, qd, x_t, xd, f)
    finally:
        o.set_task_gains(np.full(6, 100.0), np.full(6, 20.0))
    # one robot, reference shapes
    assert rd.update_state(q[0, :3], q[0, 3:3 + w], q[0, 3 + w:], qd[0, :3], qd[0, 3:3 + w], qd[0, 3 + w:]) is True
    assert rd.get_pose(LINK).shape == (4, 4) and rd.get_jacobian_actuated(LINK).shape == (6, o.act)
    assert rd.get_mass_matrix_actuated().shape == (o.act, o.act) and rd.get_gravity_actuated().shape == (o.act,)
    assert np.abs(rd.get_mobile_base_vel() - rd.get_mobile_FK_jacobian() @ qd[0, 3:3 + w]).max() < 1e-14
    mob, mani = rc.QPIK_step(c_oracle.pose44(x_t[0]), xd[0], LINK)
    assert mob.shape == (w,) and mani.shape == (7,)
    assert np.abs(np.concatenate([mob, mani]) - ref["out"][0]).max() < 1e-4
    # full-dof getters inherited from Manipulator::RobotData (mobile_manipulator/robot_data.h:42) and the selection matrix
    full, st = o.update_state(q[:1], qd[:1], f), o.moma_update_state(q[:1], qd[:1], f)
    n = o.nv
    assert rd.get_jacobian(LINK).shape == (6, n) and rel(rd.get_jacobian(LINK), full["J"][0]) < 1e-12
    assert rel(rd.get_jacobian_time_variation(LINK), full["Jdot"][0]) < 1e-11
    assert rel(rd.get_mass_matrix(), full["M"][0]) < 1e-9 and rel(rd.get_gravity(), full["g"][0]) < 1e-9
    assert rel(rd.get_nonlinear_effects(), full["nle"][0]) < 1e-9 and rel(rd.get_coriolis(), full["nle"][0] - full["g"][0]) < 1e-8
    # (PinvCOD with the reference's 1e-6 rank threshold, robot_data.cpp:118: the 14-dof model's M is numerically rank deficient,
    # so M^-1 M = I is NOT expected; the oracle applies the same rule)
    assert rel(rd.get_mass_matrix_inv(), full["Minv"][0]) < 1e-6
    S = rd.get_selection_matrix()
    assert S.shape == (n, o.act) and np.abs(S - st["S"][0]).max() < 1e-12
    assert np.abs(rd.get_jacobian_actuated(LINK) - rd.get_jacobian(LINK) @ S).max() < 1e-10      # J~ = J S (robot_data.cpp:407-410)
    md, mref = rd.get_min_distance(True, True), o.min_distance(q[:1], qd[:1], with_graddot=True)
    assert abs(md.distance - mref["d"][0]) < 1e-8 and md.grad.shape == (n,) and np.abs(md.grad - mref["grad"][0]).max() < 1e-4
    lo, hi = rd.get_joint_position_limit()
    assert lo.shape == (n,) and np.abs(rd.get_manipulator_joint_position() - q[0, 3 + w:]).max() == 0
    assert np.abs(rd.get_virtual_joint_velocity() - qd[0, :3]).max() == 0 and np.abs(rd.get_mobile_joint_position() - q[0, 3:3 + w]).max() == 0
    # stateless twins: same values as the cached getters at the same state, cache untouched
    args = (q[1, :3], q[1, 3:3 + w], q[1, 3 + w:])
    vargs = (qd[1, :3], qd[1, 3:3 + w], qd[1, 3 + w:])
    full1 = o.update_state(q[1:2], qd[1:2], f)
    assert rel(rd.compute_mass_matrix(*args), full1["M"][0]) < 1e-9 and rel(rd.compute_gravity(*args), full1["g"][0]) < 1e-9
    assert rel(rd.compute_nonlinear_effects(*args, *vargs), full1["nle"][0]) < 1e-9
    assert rel(rd.compute_jacobian(*args, LINK), full1["J"][0]) < 1e-12
    assert rel(rd.compute_jacobian_time_variation(*args, *vargs, LINK), full1["Jdot"][0]) < 1e-11
    assert np.abs(rd.compute_pose(*args, LINK)[:3] - full1["pose"][0].reshape(3, 4)).max() < 1e-12
    S1 = rd.compute_selection_matrix(q[1, :3], q[1, 3:3 + w])
    assert np.abs(S1 - o.moma_update_state(q[1:2], qd[1:2], f)["S"][0]).max() < 1e-12
    # the reference's actuated twins evaluate at q_virtual = 0 and multiply by S(q_virtual) (robot_data.cpp:185-232, 382-405)
    q0 = q[1:2].copy(); q0[:, :3] = 0
    f0 = o.update_state(q0, np.zeros_like(q0), f)
    assert rel(rd.compute_jacobian_actuated(*args, LINK), f0["J"][0] @ S1) < 1e-11
    assert rel(rd.compute_mass_matrix_actuated(*args), S1.T @ f0["M"][0] @ S1) < 1e-9
    assert rel(rd.compute_gravity_actuated(*args), S1.T @ f0["g"][0]) < 1e-9
    mm = rd.compute_manipulability(q[1, 3 + w:], qd[1, 3 + w:], True, True, LINK)
    assert abs(mm.manipulability - o.moma_update_state(q[1:2], qd[1:2], f)["mani"][0]) < 1e-11 and mm.grad.shape == (7,)
    assert np.abs(rd.get_pose(LINK)[:3] - full["pose"][0].reshape(3, 4)).max() < 1e-12            # cache still holds robot 0
    # manipulator joint-space helpers (robot_controller.cpp:78-145): M_mani qddot + g_mani with the PD acceleration
    ms = d["joint_idx"]["mani_start"]
    qt, qdt = q[0, 3 + w:] + 0.05, 0.5 * qd[0, 3 + w:]
    acc = 400.0 * (qt - q[0, 3 + w:]) + 40.0 * (qdt - qd[0, 3 + w:])
    tau_ref = full["M"][0][ms:ms + 7, ms:ms + 7] @ acc + full["g"][0][ms:ms + 7]
    assert rel(rc.move_manipulator_joint_torque_step(qt, qdt), tau_ref) < 1e-9
    assert rel(rc.move_manipulator_joint_torque_step(qddot_mani_target=acc), tau_ref) < 1e-9
    assert np.abs(rc.move_manipulator_joint_position_cubic(qt, qdt, q[0, 3 + w:], qd[0, 3 + w:], 2.0, 0.0, 1.0) - qt).max() == 0
    with pytest.raises(RuntimeError):
        rc.set_manipulator_joint_gain(np.ones(3), np.ones(7))
    # batch
    rd.update_state(q[:, :3], q[:, 3:3 + w], q[:, 3 + w:], qd[:, :3], qd[:, 3:3 + w], qd[:, 3 + w:])
    mob, mani = rc.QPIK_step(c_oracle.pose44(x_t), xd, LINK)
    same = rc.last_iters == ref["iters"]
    assert same.mean() > 0.9 and np.abs(np.concatenate([mob, mani], axis=1) - ref["out"])[same].max() < 1e-4
    acc, tau = rc.QPID_step(c_oracle.pose44(x_t), xd, LINK)
    same = rc.last_iters == ref3["iters"]
    assert acc.shape == (32, w) and tau.shape == (32, 7)
    assert np.abs(tau - ref3["out"][:, w:])[same].max() < 1e-4 * max(1.0, np.abs(ref3["out"]).max())
    assert np.abs(acc - ref3["out2"][:, :w])[same].max() < 1e-4 * max(1.0, np.abs(ref3["out2"]).max())
    assert rd.get_jacobian(LINK).shape == (32, 6, n) and rd.get_selection_matrix().shape == (32, n, o.act)


def test_moma_full_size_properties(rig):
    """BASELINE config 4/5 batch on one GPU shard (8192 here): determinism, permutation equivariance, fallbacks."""
    name, d, o, model, ctx = rig
    B = 8192
    q, qd, q_t, xd = moma_workload(o.model, o.w, B, 60)
    qd = consistent(o, q, qd)
    ctx.moma_update_state(q_t, qd)
    x_t = ctx.moma_get_state(LINK, want=("pose",))["pose"]
    r = ctx.moma_cycle("ik", q, qd, x_t, xd, LINK)
    assert (r["status"] == 1).mean() > 0.9
    assert (r["iters"] % 25 == 0).all()
    r2 = ctx.moma_cycle("ik", q, qd, x_t, xd, LINK)
    assert np.array_equal(r2["out"], r["out"]) and (r2["iters"] == r["iters"]).all()
    perm = np.random.default_rng(0).permutation(B)
    r3 = ctx.moma_cycle("ik", q[perm], qd[perm], x_t[perm], xd[perm], LINK)
    assert np.array_equal(r3["out"], r["out"][perm])
