"""GPU parity tests of the mobile base path through the C ABI against the oracle:
Mobile::RobotData (FK Jacobian, base velocity) and Mobile::RobotController (IK Jacobian, wheel velocities,
VelocityCommand) for differential / mecanum / powered-caster bases, and the powered-caster mobile manipulator
(state-dependent base Jacobian inside the whole-body kernels)."""
import numpy as np
import pytest

from tests.conftest import MOMA, moma_workload
from tests.test_mobile_cpu import KINS, wheels_of

pytestmark = pytest.mark.gpu
LINK = "fr3_link8"


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module", params=list(KINS))
def rig(request):
    import dyros_robot_controller_b200 as drc
    if drc.device_count() < 1:
        pytest.fail("GPU tests need a CUDA device; the product path has no CPU fallback")
    kin = KINS[request.param]
    return request.param, kin, wheels_of(kin), drc.MobileBase(kin, device=0)


def workload(w, B, seed):
    rng = np.random.default_rng(seed)
    wp = rng.uniform(-np.pi, np.pi, (B, w))
    wv = rng.uniform(-3, 3, (B, w))
    bv = rng.normal(size=(B, 3)) * np.array([1.5, 1.5, 3.0])
    bv[0] = [1e-5, -2e-5, 0.3]
    bv[1] = [10.0, 0.0, -9.0]
    return wp, wv, bv


def test_mobile_base_matches_oracle_host_buffers(rig):
    from oracle.c_oracle import mobile_base
    name, kin, w, base = rig
    wp, wv, bv = workload(w, 5000, 71)
    n0 = base.launch_count()
    J, vel = base.fk(wp, wv)
    Jr, velr = mobile_base(kin, True, wp, wv)
    assert rel(J, Jr) < 1e-12 and rel(vel, velr) < 1e-12
    for sat in (False, True):
        Ji, wheel = base.ik(wp, bv, saturate=sat)
        Jir, wheelr = mobile_base(kin, False, wp, bv, saturate=sat)
        assert rel(Ji, Jir) < 1e-12 and rel(wheel, wheelr) < 1e-11
    assert base.launch_count() == n0 + 3          # the CUDA kernels ran (no host path exists)


def test_mobile_base_device_tensors_and_full_size(rig):
    """device pointers on the current torch stream; 1 M bases: determinism, permutation equivariance, and the identity
    J_fk (J_ik v) = v for the mecanum drive / (vx, 0, omega) for the differential drive."""
    import torch
    from oracle.c_oracle import mobile_base
    name, kin, w, base = rig
    B = 1 << 20
    wp, wv, bv = workload(w, B, 72)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a).to(dev)
    J, vel = base.fk(t(wp), t(wv))
    Ji, wheel = base.ik(t(wp), t(bv), saturate=True)
    torch.cuda.synchronize()
    vel, wheel = vel.cpu().numpy(), wheel.cpu().numpy()
    sub = np.random.default_rng(0).choice(B, 2000, replace=False)
    _, velr = mobile_base(kin, True, wp[sub], wv[sub])
    _, wheelr = mobile_base(kin, False, wp[sub], bv[sub], saturate=True)
    assert rel(vel[sub], velr) < 1e-12 and rel(wheel[sub], wheelr) < 1e-11
    perm = np.random.default_rng(1).permutation(B)
    _, vel2 = base.fk(t(wp[perm]), t(wv[perm]), want_J=False)
    assert np.array_equal(vel2.cpu().numpy(), vel[perm])
    if not name.startswith("caster"):
        _, back = base.fk(t(wp), torch.from_numpy(wheel).to(dev), want_J=False)
        back = back.cpu().numpy()
        sp = np.hypot(bv[:, 0], bv[:, 1])
        d = np.where((sp < 1e-4)[:, None], 0.0, bv[:, :2] / np.maximum(sp, 1e-300)[:, None])
        want = np.concatenate([d * np.minimum(sp, kin["max_lin_speed"])[:, None],
                               np.clip(bv[:, 2:3], -kin["max_ang_speed"], kin["max_ang_speed"])], axis=1)
        if name == "differential":
            want[:, 1] = 0.0
        assert np.abs(back - want).max() < 1e-9


def test_mobile_reference_api_mirror(rig):
    """dyros_robot_controller_b200.drc.mobile mirrors the reference classes (drc/mobile/robot_data.py,
    drc/mobile/robot_controller.py): single-base arrays in and out, or a leading batch axis."""
    from dyros_robot_controller_b200.drc import DriveType, KinematicParam
    from dyros_robot_controller_b200.drc.mobile import RobotController, RobotData
    from oracle.c_oracle import mobile_base
    name, kin, w, base = rig
    kp = KinematicParam(type=DriveType[kin["type"]], wheel_radius=kin["wheel_radius"], max_lin_speed=kin["max_lin_speed"],
                        max_ang_speed=kin["max_ang_speed"], base_width=kin.get("base_width"),
                        roller_angles=kin.get("roller_angles"), base2wheel_positions=kin.get("base2wheel_positions"),
                        base2wheel_angles=kin.get("base2wheel_angles"), wheel_offset=kin.get("wheel_offset"))
    rd = RobotData(kp)
    rc = RobotController(0.001, rd)
    wp, wv, bv = workload(w, 16, 73)
    assert rd.get_wheel_num() == w
    assert rd.update_state(wp[3], wv[3]) is True
    Jr, velr = mobile_base(kin, True, wp[3:4], wv[3:4])
    Jir, wheelr = mobile_base(kin, False, wp[3:4], bv[3:4], saturate=True)
    assert rd.get_FK_jacobian().shape == (3, w) and rel(rd.get_FK_jacobian(), Jr[0]) < 1e-12
    assert rd.get_base_vel().shape == (3,) and rel(rd.get_base_vel(), velr[0]) < 1e-12
    assert rel(rd.compute_fk_jacobian(wp[5]), mobile_base(kin, True, wp[5:6], None)[0][0]) < 1e-12
    assert rel(rd.compute_base_vel(wp[5], wv[5]), mobile_base(kin, True, wp[5:6], wv[5:6])[1][0]) < 1e-12
    assert rc.compute_IK_jacobian().shape == (w, 3) and rel(rc.compute_IK_jacobian(), Jir[0]) < 1e-12
    assert rc.velocity_command(bv[3]).shape == (w,) and rel(rc.velocity_command(bv[3]), wheelr[0]) < 1e-11
    assert rel(rc.compute_wheel_vel(bv[3]), mobile_base(kin, False, wp[3:4], bv[3:4])[1][0]) < 1e-11
    # batch
    rd.update_state(wp, wv)
    assert rel(rd.get_base_vel(), mobile_base(kin, True, wp, wv)[1]) < 1e-12
    assert rel(rc.velocity_command(bv), mobile_base(kin, False, wp, bv, saturate=True)[1]) < 1e-11


def test_mobile_create_rejects_bad_parameters():
    import dyros_robot_controller_b200 as drc
    with pytest.raises(RuntimeError):
        drc.MobileBase(dict(type="Caster", wheel_radius=0.05, wheel_offset=0.02, base2wheel_positions=[(0.2, 0.1)]))  # one caster
    with pytest.raises(RuntimeError):
        drc.MobileBase(dict(type="Differential", wheel_radius=0.1))       # no base width
    with pytest.raises(RuntimeError):
        drc.MobileBase(dict(type="Caster", wheel_radius=0.05, wheel_offset=0.0, base2wheel_positions=[(0.2, 0.1), (-0.2, -0.1)]))


# ---- powered-caster mobile manipulator
@pytest.fixture(scope="module")
def pcv():
    import dyros_robot_controller_b200 as drc
    from oracle.c_oracle import MomaOracle
    d = MOMA["pcv_fr3"]
    o = MomaOracle(d["urdf"], d["srdf"], d["kin"], d["joint_idx"], d["actuator_idx"], threads=8)
    model = drc.Model(d["urdf"], d["srdf"]).attach_mobile_base(d["kin"], d["joint_idx"], d["actuator_idx"])
    ctx = drc.Context(model, 4096, device=0)
    return d, o, model, ctx


def test_caster_moma_state(pcv):
    d, o, model, ctx = pcv
    assert model.actuated_dof == 11 and model.wheel_num == 4 and model.drive_type == 2
    f = o.frame_id(LINK)
    q, qd, _, _ = moma_workload(o.model, o.w, 1000, 81)
    ref, full = o.moma_update_state(q, qd, f), o.update_state(q, qd, f)
    ctx.moma_update_state(q, qd)
    r = ctx.moma_get_state(LINK)
    assert rel(r["pose"], full["pose"]) < 1e-12
    assert rel(r["J"], ref["J"]) < 1e-12 and rel(r["Jdot"], ref["Jdot"]) < 1e-11
    assert rel(r["M"], ref["M"]) < 1e-9 and rel(r["g"], ref["g"]) < 1e-9 and rel(r["nle"], ref["nle"]) < 1e-9
    assert rel(r["Minv"], ref["Minv"]) < 1e-7
    assert np.abs(r["mani"] - ref["mani"]).max() < 1e-11


@pytest.mark.parametrize("mode,B", [(1, 2000), (3, 1000)])
def test_caster_moma_control_cycle(pcv, mode, B):
    d, o, model, ctx = pcv
    f = o.frame_id(LINK)
    q, qd, q_t, xd = moma_workload(o.model, o.w, B, 82 + mode)
    x_t = o.update_state(q_t, qd, f)["pose"]
    ref = o.moma_cycle(mode, q, qd, x_t, xd, f)
    r = ctx.moma_cycle("ik" if mode == 1 else "id", q, qd, x_t, xd, LINK)
    assert (r["status"] == ref["status"]).mean() > 0.98
    same = (r["iters"] == ref["iters"]) & (r["status"] == ref["status"])
    assert same.mean() > 0.97
    scale = max(1.0, np.abs(ref["out"]).max())
    err = np.abs(r["out"] - ref["out"]).max(axis=1)[same]
    assert (err < 1e-4 * scale).mean() > 0.98 and err.max() < 5e-2 * scale
    assert (ref["status"] == 1).mean() > 0.8


def test_caster_moma_reference_api_mirror(pcv):
    from dyros_robot_controller_b200.drc import ActuatorIndex, DriveType, JointIndex, KinematicParam
    from dyros_robot_controller_b200.drc.mobile_manipulator import RobotController, RobotData
    from oracle import c_oracle
    d, o, model, ctx = pcv
    kp = KinematicParam(type=DriveType.Caster, wheel_radius=d["kin"]["wheel_radius"], wheel_offset=d["kin"]["wheel_offset"],
                        base2wheel_positions=d["kin"]["base2wheel_positions"])
    rd = RobotData(kp, JointIndex(**d["joint_idx"]), ActuatorIndex(**d["actuator_idx"]), d["urdf"], d["srdf"], max_batch=16)
    rc = RobotController(0.001, rd)
    f = o.frame_id(LINK)
    q, qd, q_t, xd = moma_workload(o.model, o.w, 16, 85)
    x_t = o.update_state(q_t, qd, f)["pose"]
    o.set_task_gains(np.full(6, 400.0), np.full(6, 40.0))   # MobileManipulator::RobotController defaults (robot_controller.cpp:15-16)
    ref = o.moma_cycle(1, q, qd, x_t, xd, f)
    o.set_task_gains(np.full(6, 100.0), np.full(6, 20.0))
    rd.update_state(q[0, :3], q[0, 3:7], q[0, 7:], qd[0, :3], qd[0, 3:7], qd[0, 7:])
    Jm, bv = o.mobile_state(q[0:1, 3:7], qd[0:1, 3:7])
    assert rd.get_FK_jacobian().shape == (3, 4) and rel(rd.get_FK_jacobian(), Jm[0]) < 1e-12
    assert rel(rd.get_base_vel(), bv[0]) < 1e-12
    mob, mani = rc.QPIK_step(c_oracle.pose44(x_t[0]), xd[0], LINK)
    assert mob.shape == (4,) and mani.shape == (7,)
    if ref["status"][0] == 1:
        assert np.abs(np.concatenate([mob, mani]) - ref["out"][0]).max() < 1e-4 * max(1.0, np.abs(ref["out"][0]).max())
