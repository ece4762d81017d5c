"""Mobile base + mobile manipulator on the CPU: the oracle restatement pinned by analytic
anchors and an independent numpy computation, and the product's kernel bodies (host emulation) against the oracle."""
import numpy as np
import pytest

from tests.conftest import MOMA, moma_workload

LINK = "fr3_link8"


@pytest.fixture(scope="module", params=["husky_fr3", "xls_fr3"])
def robots(request):
    from oracle.c_oracle import MomaOracle
    from tests.emu import MomaEmu
    d = MOMA[request.param]
    o = MomaOracle(d["urdf"], d["srdf"], d["kin"], d["joint_idx"], d["actuator_idx"], threads=8)
    e = MomaEmu(d["urdf"], d["srdf"], d["kin"], d["joint_idx"], d["actuator_idx"])
    return request.param, d, o, e


def consistent_base_state(o, q, qd):
    """virtual-joint velocities consistent with the wheels: qdot_virtual = Rz(yaw) J_mobile qdot_wheel."""
    J, bv = o.mobile_state(q[:, 3:3 + o.w], qd[:, 3:3 + o.w])
    c, s = np.cos(q[:, 2]), np.sin(q[:, 2])
    qd = qd.copy()
    qd[:, 0] = c * bv[:, 0] - s * bv[:, 1]
    qd[:, 1] = s * bv[:, 0] + c * bv[:, 1]
    qd[:, 2] = bv[:, 2]
    return qd


def test_base_jacobians_analytic(robots):
    name, d, o, e = robots
    J, bv = o.mobile_state(np.zeros((1, o.w)), np.ones((1, o.w)))
    if name == "husky_fr3":          # mobile/robot_data.cpp:138-147
        r, b = 0.1651, 0.555
        assert np.allclose(J[0], [[r / 2, r / 2], [0, 0], [-r / b, r / b]])
        assert np.allclose(bv[0], [r, 0, 0])                  # both wheels forward -> pure translation
    else:                            # mecanum: J = pinv(J_inv), J_inv rows (1/r)[1, tan(g)] [[1,0,-py],[0,1,px]]  (:149-177)
        r = 0.120
        Ji = np.array([[1, np.tan(g), -py + np.tan(g) * px] for g, (px, py) in zip(d["kin"]["roller_angles"], d["kin"]["base2wheel_positions"])]) / r
        assert np.allclose(J[0], np.linalg.pinv(Ji), atol=1e-12)
        assert np.allclose(J[0] @ Ji, np.eye(3), atol=1e-12)
        assert np.allclose(bv[0], [r, 0, 0], atol=1e-12)      # all wheels forward -> pure translation
    assert np.abs(e.base_jacobian() - J[0]).max() < 1e-12    # product's host model compiler (csrc/model.cpp:299-350)


def test_oracle_moma_state_matches_numpy(robots):
    """S, M~ = S'MS, g~, nle~, J~ = J S recomputed in numpy from the full-model quantities."""
    name, d, o, e = robots
    f = o.frame_id(LINK)
    q, qd, _, _ = moma_workload(o.model, o.w, 5, 3)
    full = o.update_state(q, qd, f)
    ms = o.moma_update_state(q, qd, f)
    ji, ai = d["joint_idx"], d["actuator_idx"]
    Jm, _ = o.mobile_state(q[:, 3:3 + o.w], qd[:, 3:3 + o.w])
    for b in range(5):
        S = np.zeros((o.nv, o.act))
        S[ji["mani_start"]:ji["mani_start"] + o.mani, ai["mani_start"]:ai["mani_start"] + o.mani] = np.eye(o.mani)
        S[ji["mobi_start"]:ji["mobi_start"] + o.w, ai["mobi_start"]:ai["mobi_start"] + o.w] = np.eye(o.w)
        c, s = np.cos(q[b, 2]), np.sin(q[b, 2])
        S[0:3, ai["mobi_start"]:ai["mobi_start"] + o.w] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]]) @ Jm[b]
        assert np.abs(ms["S"][b] - S).max() < 1e-13
        assert np.abs(ms["M"][b] - S.T @ full["M"][b] @ S).max() < 1e-10
        assert np.abs(ms["g"][b] - S.T @ full["g"][b]).max() < 1e-10
        assert np.abs(ms["nle"][b] - S.T @ full["nle"][b]).max() < 1e-10
        assert np.abs(ms["J"][b] - full["J"][b] @ S).max() < 1e-12
        assert np.abs(ms["Jdot"][b] - full["Jdot"][b] @ S).max() < 1e-11
        assert np.abs(ms["Minv"][b] @ ms["M"][b] - np.eye(o.act)).max() < 1e-8
        # manipulability of the arm does not depend on the base (mobile_manipulator/robot_data.cpp:447)
        Jarm = full["J"][b][:, ji["mani_start"]:ji["mani_start"] + o.mani]
        assert abs(ms["mani"][b] - np.sqrt(np.linalg.det(Jarm @ Jarm.T))) < 1e-10
    # the full model carries the whole moving mass: base + wheels + arm
    assert abs(o.model.mass.sum() - {"husky_fr3": 46.0 + 2 * 2.637, "xls_fr3": 125.0 + 4 * 6.5}[name] - 15.4948010262 - 2.3966) < 0.05


def test_oracle_moma_qp_structure(robots):
    name, d, o, e = robots
    f = o.frame_id(LINK)
    q, qd, _, xd = moma_workload(o.model, o.w, 1, 4)
    P, qv, A, l, u = o.moma_build_qp(0, q[0], qd[0], xd[0], f)
    act, k = o.act, o.mani
    assert P.shape == (act, act) and A.shape == (act + 2 * k + 2, act)          # 9/25 and 11/27
    assert np.allclose(A[:act], np.eye(act)) and (l[:act] <= -1e30).all() and (u[:act] >= 1e30).all()   # free bounds
    ms = o.moma_update_state(q, qd, f)
    assert np.abs(P - (2 * ms["J"][0].T @ ms["J"][0] + 0.01 * np.eye(act))).max() < 1e-10              # QP_IK.cpp:70-71
    am = d["actuator_idx"]["mani_start"]
    assert np.allclose(A[act:act + k, am:am + k], np.eye(k)) and np.allclose(A[act + k:act + 2 * k, am:am + k], -np.eye(k))
    assert np.abs(A[act:, :am]).max() == 0.0                                     # CBF rows touch manipulator columns only
    P2, q2, A2, l2, u2 = o.moma_build_qp(1, q[0], qd[0], xd[0], f)
    assert P2.shape == (2 * act, 2 * act) and A2.shape == (4 * k + 2 + act, 2 * act)   # 18/39 and 22/41, no bound rows
    assert np.abs(A2[4 * k + 2:, :act] - ms["M"][0]).max() < 1e-12 and np.allclose(A2[4 * k + 2:, act:], -np.eye(act))
    assert np.allclose(l2[4 * k + 2:], -ms["g"][0]) and np.allclose(u2[4 * k + 2:], -ms["g"][0])


def test_kernel_bodies_moma_state(robots):
    name, d, o, e = robots
    f = o.frame_id(LINK)
    q, qd, _, _ = moma_workload(o.model, o.w, 200, 5)
    ref = o.moma_update_state(q, qd, f)
    full = o.update_state(q, qd, f)
    r = e.moma_state(q, qd, e.frame_id(LINK))
    rel = lambda a, b: np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
    assert rel(r["pose"], full["pose"]) < 1e-12
    assert rel(r["J"], ref["J"]) < 1e-12 and rel(r["Jdot"], ref["Jdot"]) < 1e-11
    assert rel(r["M"], ref["M"]) < 1e-9 and rel(r["g"], ref["g"]) < 1e-9 and rel(r["nle"], ref["nle"]) < 1e-9
    assert rel(r["Minv"], ref["Minv"]) < 1e-7
    assert np.abs(r["mani"] - ref["mani"]).max() < 1e-11
    assert np.abs(r["mani_grad"] - ref["mani_grad"]).max() < 1e-9 * max(1.0, np.abs(ref["mani_grad"]).max())
    assert np.abs(r["mani_graddot"] - ref["mani_graddot"]).max() < 1e-8 * max(1.0, np.abs(ref["mani_graddot"]).max())


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_kernel_bodies_moma_cycle(robots, mode):
    """whole-body QPIK / QPID: structured ADMM without slacks (hard CBF rows) == dense OSQP restatement."""
    name, d, o, e = robots
    f = o.frame_id(LINK)
    B = 150
    q, qd, q_t, xd = moma_workload(o.model, o.w, B, 6 + mode)
    qd = consistent_base_state(o, q, qd)
    x_t = o.update_state(q_t, qd, f)["pose"] if mode in (1, 3) else None
    des = xd if mode in (1, 3) else 3.0 * xd
    ref = o.moma_cycle(mode, q, qd, x_t, des, f)
    r = e.moma_cycle(mode, q, qd, x_t, des, e.frame_id(LINK))
    assert (r["status"] == ref["status"]).mean() > 0.98
    same = (r["iters"] == ref["iters"]) & (r["status"] == ref["status"])
    assert same.mean() > 0.97, (np.bincount(ref["status"]), np.bincount(r["status"]))
    scale = max(1.0, np.abs(ref["out"]).max())
    err = np.abs(r["out"] - ref["out"]).max(axis=1)[same]
    assert (err < 1e-5 * scale).mean() > 0.99 and err.max() < 1e-3 * scale
    if mode >= 2:
        err2 = np.abs(r["out2"] - ref["out2"]).max(axis=1)[same]
        assert (err2 < 1e-5 * max(1.0, np.abs(ref["out2"]).max())).mean() > 0.99
    assert (ref["status"] == 1).mean() > 0.9
