"""Mobile base alone on the CPU: the oracle restatement of Mobile::RobotData /
Mobile::RobotController pinned by the closed forms of the reference (recomputed in numpy with numpy.linalg.pinv) and by
the forward / inverse consistency of the Jacobians; the product's kernel bodies (host emulation) against the oracle; the
powered-caster whole-body path (state-dependent base Jacobian inside the mobile-manipulator kernels)."""
import numpy as np
import pytest

from tests.conftest import MOMA, moma_workload

KINS = {
    # Husky-class differential drive
    "differential": dict(type="Differential", wheel_radius=0.1651, base_width=0.555, max_lin_speed=1.0, max_ang_speed=2.0),
    # the reference's XLS example (examples/C++/src/xls_controller.cpp:18-27)
    "mecanum": dict(type="Mecanum", wheel_radius=0.120, max_lin_speed=2.0, max_ang_speed=2.0,
                    roller_angles=[-np.pi / 4, np.pi / 4, np.pi / 4, -np.pi / 4],
                    base2wheel_positions=[(0.2225, 0.2045), (0.2225, -0.2045), (-0.2225, 0.2045), (-0.2225, -0.2045)],
                    base2wheel_angles=[0.0, 0.0, 0.0, 0.0]),
    # powered-caster vehicle with four casters (8 joints: steer, roll per caster)
    "caster4": dict(type="Caster", wheel_radius=0.055, wheel_offset=0.020, max_lin_speed=1.5, max_ang_speed=3.0,
                    base2wheel_positions=[(0.215, 0.125), (0.215, -0.125), (-0.215, 0.125), (-0.215, -0.125)]),
    "caster2": dict(type="Caster", wheel_radius=0.055, wheel_offset=0.020, max_lin_speed=1.5, max_ang_speed=3.0,
                    base2wheel_positions=[(0.215, 0.125), (-0.215, -0.125)]),
}


def wheels_of(kin):
    return 2 if kin["type"] == "Differential" else (len(kin["roller_angles"]) if kin["type"] == "Mecanum"
                                                   else 2 * len(kin["base2wheel_positions"]))


def numpy_fk_jacobian(kin, wheel_pos):
    """mobile/robot_data.cpp:138-203, literally, with numpy.linalg.pinv in place of PinvCOD"""
    r = kin["wheel_radius"]
    if kin["type"] == "Differential":
        b = kin["base_width"]
        return np.array([[r / 2, r / 2], [0, 0], [-r / b, r / b]])
    if kin["type"] == "Mecanum":
        return np.linalg.pinv(numpy_ik_jacobian(kin, wheel_pos))
    w = wheels_of(kin)
    b = kin["wheel_offset"]
    Jp, Jq = np.zeros((w, 3)), np.zeros((w, w))
    for i, (px, py) in enumerate(kin["base2wheel_positions"]):
        phi = wheel_pos[2 * i]
        Jp[2 * i:2 * i + 2] = [[1, 0, -(py + b * np.sin(phi))], [0, 1, px + b * np.cos(phi)]]
        Jq[2 * i:2 * i + 2, 2 * i:2 * i + 2] = [[b * np.sin(phi), r * np.cos(phi)], [-b * np.cos(phi), r * np.sin(phi)]]
    return np.linalg.pinv(Jp.T @ Jp) @ Jp.T @ Jq


def numpy_ik_jacobian(kin, wheel_pos):
    """mobile/robot_controller.cpp:65-123"""
    r = kin["wheel_radius"]
    if kin["type"] == "Differential":
        b = kin["base_width"]
        return np.array([[1 / r, 0, -b / (2 * r)], [1 / r, 0, b / (2 * r)]])
    if kin["type"] == "Mecanum":
        rows = []
        for g, (px, py), pt in zip(kin["roller_angles"], kin["base2wheel_positions"], kin["base2wheel_angles"]):
            A1 = np.array([[1, 0, -py], [0, 1, px]])
            A2 = np.array([[np.cos(pt), np.sin(pt)], [-np.sin(pt), np.cos(pt)]])
            rows.append((np.array([[1, np.tan(g)]]) @ A2 @ A1)[0] / r)
        return np.array(rows)
    b = kin["wheel_offset"]
    rows = []
    for i, (px, py) in enumerate(kin["base2wheel_positions"]):
        phi = wheel_pos[2 * i]
        rows.append([-np.sin(phi) / b, np.cos(phi) / b, (px * np.cos(phi) + py * np.sin(phi)) / b - 1])
        rows.append([np.cos(phi) / r, np.sin(phi) / r, (px * np.sin(phi) - py * np.cos(phi)) / r])
    return np.array(rows)


def numpy_saturate(kin, v):
    """mobile/robot_controller.cpp:14-41"""
    sp = np.hypot(v[0], v[1])
    d = np.zeros(2) if abs(sp) < 1e-4 else v[:2] / sp
    sp = min(max(sp, -kin["max_lin_speed"]), kin["max_lin_speed"])
    return np.array([d[0] * sp, d[1] * sp, min(max(v[2], -kin["max_ang_speed"]), kin["max_ang_speed"])])


@pytest.fixture(scope="module", params=list(KINS))
def base(request):
    kin = KINS[request.param]
    w = wheels_of(kin)
    rng = np.random.default_rng(11)
    B = 64
    wp = rng.uniform(-np.pi, np.pi, (B, w))
    wv = rng.uniform(-3, 3, (B, w))
    bv = rng.normal(size=(B, 3)) * np.array([1.5, 1.5, 3.0])
    bv[0] = [1e-5, -2e-5, 0.3]     # below the 1e-4 speed threshold: the direction is dropped
    bv[1] = [10.0, 0.0, -9.0]      # far above both limits
    return request.param, kin, w, wp, wv, bv


def test_oracle_mobile_matches_reference_closed_forms(base):
    from oracle.c_oracle import mobile_base
    name, kin, w, wp, wv, bv = base
    J, vel = mobile_base(kin, True, wp, wv)
    Ji, wheel = mobile_base(kin, False, wp, bv)
    _, wheel_sat = mobile_base(kin, False, wp, bv, saturate=True)
    for b in range(len(wp)):
        Jn, Jin = numpy_fk_jacobian(kin, wp[b]), numpy_ik_jacobian(kin, wp[b])
        assert np.abs(J[b] - Jn).max() < 1e-12 * max(1.0, np.abs(Jn).max())
        assert np.abs(Ji[b] - Jin).max() < 1e-12 * max(1.0, np.abs(Jin).max())
        assert np.abs(vel[b] - Jn @ wv[b]).max() < 1e-12 * max(1.0, np.abs(Jn @ wv[b]).max())
        assert np.abs(wheel[b] - Jin @ bv[b]).max() < 1e-11 * max(1.0, np.abs(Jin @ bv[b]).max())
        vs = numpy_saturate(kin, bv[b])
        assert np.abs(wheel_sat[b] - Jin @ vs).max() < 1e-11 * max(1.0, np.abs(Jin @ vs).max())


def test_forward_and_inverse_jacobians_are_consistent(base):
    """J_fk J_ik = I for the mecanum drive (J_fk is the pseudo-inverse of J_ik); the differential drive reproduces
    (vx, 0, omega).  The reference's two CASTER Jacobians are not mutual inverses: CasterFKJacobian's J_q^-1 places the
    contact point at +b along the steering direction (mobile/robot_data.cpp:193-199) while CasterIKJacobian trails it
    (mobile/robot_controller.cpp:118-119) -- reproduced as written (quirk Q13); what holds for casters is
    J_fk = pinv(J_p~) J_q^-1 and that a pure steering-axis twist gives J_ik rows as in the reference."""
    from oracle.c_oracle import mobile_base
    name, kin, w, wp, wv, bv = base
    J, _ = mobile_base(kin, True, wp, None)
    Ji, _ = mobile_base(kin, False, wp, None)
    if name.startswith("caster"):
        r, b = kin["wheel_radius"], kin["wheel_offset"]
        for k in range(8):
            Jp, Jq = np.zeros((w, 3)), np.zeros((w, w))
            for i, (px, py) in enumerate(kin["base2wheel_positions"]):
                phi = wp[k, 2 * i]
                Jp[2 * i:2 * i + 2] = [[1, 0, -(py + b * np.sin(phi))], [0, 1, px + b * np.cos(phi)]]
                Jq[2 * i:2 * i + 2, 2 * i:2 * i + 2] = [[b * np.sin(phi), r * np.cos(phi)], [-b * np.cos(phi), r * np.sin(phi)]]
            assert np.abs(J[k] - np.linalg.pinv(Jp) @ Jq).max() < 1e-12
            assert np.abs(Jp @ J[k] - Jp @ np.linalg.pinv(Jp) @ Jq).max() < 1e-12
        return
    want = np.diag([1.0, 0.0, 1.0]) if name == "differential" else np.eye(3)
    for b in range(len(wp)):
        assert np.abs(J[b] @ Ji[b] - want).max() < 1e-10


def test_kernel_bodies_mobile_match_oracle(base):
    from oracle.c_oracle import mobile_base
    from tests import emu
    name, kin, w, wp, wv, bv = base
    for fk, vec, sat in ((True, wv, False), (False, bv, False), (False, bv, True)):
        Jr, outr = mobile_base(kin, fk, wp, vec, saturate=sat)
        Je, oute = emu.mobile_base(kin, fk, wp, vec, saturate=sat)
        assert np.abs(Je - Jr).max() < 1e-12 * max(1.0, np.abs(Jr).max())
        assert np.abs(oute - outr).max() < 1e-11 * max(1.0, np.abs(outr).max())


def test_saturation_properties():
    from tests import emu
    kin = KINS["mecanum"]
    rng = np.random.default_rng(3)
    bv = rng.normal(size=(500, 3)) * 4.0
    Ji, wheel = emu.mobile_base(kin, False, np.zeros((500, 4)), bv, saturate=True)
    Jf, _ = emu.mobile_base(kin, True, np.zeros((500, 4)), None)
    got = np.einsum("brk,bk->br", Jf, wheel)         # the base velocity those wheel speeds produce
    assert (np.hypot(got[:, 0], got[:, 1]) <= kin["max_lin_speed"] + 1e-9).all()
    assert (np.abs(got[:, 2]) <= kin["max_ang_speed"] + 1e-9).all()
    inside = (np.hypot(bv[:, 0], bv[:, 1]) <= kin["max_lin_speed"]) & (np.abs(bv[:, 2]) <= kin["max_ang_speed"])
    assert inside.any() and np.abs(got[inside] - bv[inside]).max() < 1e-9   # commands inside the limits pass unchanged
    # direction of the planar velocity is kept
    big = np.hypot(bv[:, 0], bv[:, 1]) > kin["max_lin_speed"]
    cross = got[big, 0] * bv[big, 1] - got[big, 1] * bv[big, 0]
    assert np.abs(cross).max() < 1e-9


# ---- powered-caster mobile manipulator: state-dependent base Jacobian inside the whole-body kernels
@pytest.fixture(scope="module")
def pcv():
    from oracle.c_oracle import MomaOracle
    from tests.emu import MomaEmu
    d = MOMA["pcv_fr3"]
    o = MomaOracle(d["urdf"], d["srdf"], d["kin"], d["joint_idx"], d["actuator_idx"], threads=8)
    e = MomaEmu(d["urdf"], d["srdf"], d["kin"], d["joint_idx"], d["actuator_idx"])
    return d, o, e


def test_caster_moma_oracle_selection_matrix(pcv):
    d, o, e = pcv
    f = o.frame_id("fr3_link8")
    q, qd, _, _ = moma_workload(o.model, o.w, 6, 21)
    ms = o.moma_update_state(q, qd, f)
    full = o.update_state(q, qd, f)
    for b in range(6):
        Jm = numpy_fk_jacobian(KINS["caster2"], q[b, 3:7])
        c, s = np.cos(q[b, 2]), np.sin(q[b, 2])
        S = np.zeros((o.nv, o.act))
        S[7:14, 4:11] = np.eye(7)
        S[3:7, 0:4] = np.eye(4)
        S[0:3, 0:4] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]]) @ Jm
        assert np.abs(ms["S"][b] - S).max() < 1e-12
        assert np.abs(ms["M"][b] - S.T @ full["M"][b] @ S).max() < 1e-10
        assert np.abs(ms["J"][b] - full["J"][b] @ S).max() < 1e-12


def test_caster_moma_kernel_bodies_state(pcv):
    d, o, e = pcv
    f = o.frame_id("fr3_link8")
    q, qd, _, _ = moma_workload(o.model, o.w, 200, 22)
    ref = o.moma_update_state(q, qd, f)
    r = e.moma_state(q, qd, e.frame_id("fr3_link8"))
    rel = lambda a, b: np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
    assert rel(r["J"], ref["J"]) < 1e-12 and rel(r["Jdot"], ref["Jdot"]) < 1e-11
    assert rel(r["M"], ref["M"]) < 1e-9 and rel(r["g"], ref["g"]) < 1e-9 and rel(r["nle"], ref["nle"]) < 1e-9
    assert rel(r["Minv"], ref["Minv"]) < 1e-7
    assert np.abs(r["mani"] - ref["mani"]).max() < 1e-11


@pytest.mark.parametrize("mode", [1, 3])
def test_caster_moma_kernel_bodies_cycle(pcv, mode):
    d, o, e = pcv
    f = o.frame_id("fr3_link8")
    B = 100
    q, qd, q_t, xd = moma_workload(o.model, o.w, B, 23 + mode)
    x_t = o.update_state(q_t, qd, f)["pose"]
    ref = o.moma_cycle(mode, q, qd, x_t, xd, f)
    r = e.moma_cycle(mode, q, qd, x_t, xd, e.frame_id("fr3_link8"))
    same = (r["iters"] == ref["iters"]) & (r["status"] == ref["status"])
    assert same.mean() > 0.97, (np.bincount(ref["status"]), np.bincount(r["status"]))
    scale = max(1.0, np.abs(ref["out"]).max())
    err = np.abs(r["out"] - ref["out"]).max(axis=1)[same]
    # the caster base Jacobian has entries of order r, b ~ 1e-2: the unregularised whole-body QPID (QP_ID.cpp has no
    # regulariser) is worse conditioned than on the other bases, a few loosely converged iterates differ by 1e-4 relative
    assert (err < 1e-5 * scale).mean() > 0.95 and err.max() < 1e-3 * scale
    assert (ref["status"] == 1).mean() > 0.8


# ---- committed golden vectors (tools/make_golden_mobile.py): regression pin of the oracle and of the kernel bodies
def test_oracle_and_kernel_bodies_reproduce_mobile_golden_vectors(pcv):
    from pathlib import Path
    from oracle.c_oracle import mobile_base
    from tests import emu
    G = np.load(Path(__file__).resolve().parent / "golden" / "mobile_golden.npz")
    for name, kin in KINS.items():
        g = lambda k: G[f"{name}_{k}"]
        for impl in (mobile_base, emu.mobile_base):
            J, vel = impl(kin, True, g("wheel_pos"), g("wheel_vel"))
            Ji, wheel = impl(kin, False, g("wheel_pos"), g("base_vel_des"))
            _, wsat = impl(kin, False, g("wheel_pos"), g("base_vel_des"), saturate=True)
            for got, key in ((J, "J_fk"), (vel, "base_vel"), (Ji, "J_ik"), (wheel, "wheel_cmd"), (wsat, "wheel_cmd_saturated")):
                assert np.abs(got - g(key)).max() < 1e-11 * max(1.0, np.abs(g(key)).max()), (name, key)
    d, o, e = pcv
    f = o.frame_id("fr3_link8")
    ms = o.moma_update_state(G["pcv_q"], G["pcv_qd"], f)
    r = e.moma_state(G["pcv_q"], G["pcv_qd"], e.frame_id("fr3_link8"))
    for src in (ms, r):
        assert np.abs(src["M"] - G["pcv_M"]).max() < 1e-9 * np.abs(G["pcv_M"]).max()
        assert np.abs(src["J"] - G["pcv_J"]).max() < 1e-12 and np.abs(src["mani"] - G["pcv_mani"]).max() < 1e-11
    c = o.moma_cycle(1, G["pcv_q"], G["pcv_qd"], G["pcv_x_target"], G["pcv_xdot_target"], f)
    assert (c["iters"] == G["pcv_qpik_step_iters"]).all() and np.abs(c["out"] - G["pcv_qpik_step_out"]).max() < 1e-9
