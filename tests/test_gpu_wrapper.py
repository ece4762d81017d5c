"""GPU test of the reference-named extension module `dyros_robot_controller_cpp_wrapper` (pybind11 over the C ABI): the reference's
own Python classes subclass the extension classes and forward snake_case methods to camelCase ones
(/root/reference/drc/manipulator/robot_data.py:6-60); this test does exactly that and replays the call sequence of the
reference's example controller (examples/python/fr3_controller.py:100-176: update_state -> get_pose / get_velocity ->
move_joint_*_cubic / QPIK_cubic -> move_joint_torque_step -> get_gravity) for a few closed-loop ticks against the oracle."""
import sys
from pathlib import Path

import numpy as np
import pytest

from tests.conftest import LINK, MOMA, SRDF, URDF, moma_workload

pytestmark = pytest.mark.gpu
PKG = Path(__file__).resolve().parents[1] / "dyros_robot_controller_b200"


def wrapper():
    if str(PKG) not in sys.path:
        sys.path.insert(0, str(PKG))
    import dyros_robot_controller_cpp_wrapper as w
    return w


def test_fr3_controller_call_sequence(oracle):
    from oracle import c_oracle
    w = wrapper()

    class RobotData(w.ManipulatorRobotData):          # the reference's wrapper pattern
        def __init__(self, urdf_path, srdf_path="", packages_path=""):
            super().__init__(urdf_path, srdf_path, packages_path)

        def update_state(self, q, qdot): return super().updateState(q, qdot)
        def get_pose(self, link_name): return super().getPose(link_name)
        def get_velocity(self, link_name): return super().getVelocity(link_name)
        def get_gravity(self): return super().getGravity()
        def get_dof(self): return super().getDof()
        def get_verbose(self): return super().getVerbose()

    class RobotController(w.ManipulatorRobotController):
        def __init__(self, dt, robot_data):
            super().__init__(dt, robot_data)

        def move_joint_position_cubic(self, q_target, qdot_target, q_init, qdot_init, current_time, init_time, duration):
            return super().moveJointPositionCubic(q_target, qdot_target, q_init, qdot_init, current_time, init_time, duration)

        def move_joint_velocity_cubic(self, q_target, qdot_target, q_init, qdot_init, current_time, init_time, duration):
            return super().moveJointVelocityCubic(q_target, qdot_target, q_init, qdot_init, current_time, init_time, duration)

        def move_joint_torque_step(self, q_target=None, qdot_target=None, qddot_target=None):
            if qddot_target is not None:
                return super().moveJointTorqueStep(qddot_target)
            return super().moveJointTorqueStep(q_target, qdot_target)

        def QPIK_cubic(self, x_target, xdot_target, x_init, xdot_init, current_time, init_time, duration, link_name):
            return super().QPIKCubic(x_target, xdot_target, x_init, xdot_init, current_time, init_time, duration, link_name)

    dt = 0.001
    rd = RobotData(URDF, SRDF, "")
    rc = RobotController(dt, rd)
    assert rd.get_dof() == 7 and "fr3_joint1" in rd.get_verbose()
    f = oracle.frame_id(LINK)
    q = np.array([0.0, 0.0, 0.0, -np.pi / 2.0, 0.0, np.pi / 2.0, np.pi / 4.0]) + 0.1
    qdot = np.zeros(7)
    # ---- Home mode
    rd.update_state(q, qdot)
    x = rd.get_pose(LINK)
    xdot = rd.get_velocity(LINK)
    ref = oracle.update_state(q[None], qdot[None], f)
    assert x.shape == (4, 4) and np.abs(x[:3] - ref["pose"][0].reshape(3, 4)).max() < 1e-12 and np.allclose(x[3], [0, 0, 0, 1])
    q_home = np.array([0.0, 0.0, 0.0, -np.pi / 2.0, 0.0, np.pi / 2.0, np.pi / 4.0])
    q_des = rc.move_joint_position_cubic(q_target=q_home, qdot_target=np.zeros(7), q_init=q, qdot_init=qdot, init_time=0.0, current_time=1.0, duration=3.0)
    qd_des = rc.move_joint_velocity_cubic(q_target=q_home, qdot_target=np.zeros(7), q_init=q, qdot_init=qdot, init_time=0.0, current_time=1.0, duration=3.0)
    s = 1.0 / 3.0
    assert np.abs(q_des - (q + (q_home - q) * (3 * s * s - 2 * s ** 3))).max() < 1e-12
    tau = rc.move_joint_torque_step(q_target=q_des, qdot_target=qd_des)
    assert np.abs(tau - oracle.joint_torque_step(q[None], qdot[None], q_des[None], qd_des[None])[0]).max() < 1e-9 * max(1.0, np.abs(tau).max())
    # ---- QPIK mode: closed loop with ideal tracking, a few ticks
    x_init, xdot_init = x.copy(), xdot.copy()
    target_x = x_init.copy()
    target_x[:3, 3] += np.array([0.0, 0.1, 0.1])
    t = 0.0
    for k in range(5):
        rd.update_state(q, qdot)
        qdot_des = rc.QPIK_cubic(x_target=target_x, xdot_target=np.zeros(6), x_init=x_init, xdot_init=xdot_init, init_time=0.0, current_time=t,
                                 duration=2.0, link_name=LINK)
        x_des, xd_des = c_oracle.task_space_cubic(target_x, np.zeros(6), x_init, xdot_init, t, 0.0, 2.0)
        r = oracle.cycle(1, q[None], qdot[None], c_oracle.pose12(x_des)[None], xd_des[None], f)
        assert r["status"][0] == 1 and np.abs(qdot_des - r["out"][0]).max() < 1e-4
        q_desired = q + qdot_des * dt
        tau = rc.move_joint_torque_step(q_target=q_desired, qdot_target=qdot_des)
        assert np.abs(tau - oracle.joint_torque_step(q[None], qdot[None], q_desired[None], qdot_des[None])[0]).max() < 1e-8 * max(1.0, np.abs(tau).max())
        q, qdot, t = q_desired, qdot_des, t + dt
    # ---- gravity compensation mode
    rd.update_state(q, qdot)
    assert np.abs(rd.get_gravity() - oracle.update_state(q[None], qdot[None], f)["g"][0]).max() < 1e-9 * 50
    # unknown link: stderr + neutral value, like the reference (robot_data.cpp:380-384)
    assert np.abs(rd.get_pose("no_such_link") - np.eye(4)).max() == 0


def test_mobile_manipulator_classes():
    """MobileManipulatorRobotData / RobotController of the extension module: six state vectors in, (mobile, manipulator) tuples out."""
    from oracle import c_oracle
    from oracle.c_oracle import MomaOracle
    w = wrapper()
    d = MOMA["xls_fr3"]
    o = MomaOracle(d["urdf"], d["srdf"], d["kin"], d["joint_idx"], d["actuator_idx"], threads=2)
    kp = w.KinematicParam()
    kp.type = w.DriveType.Mecanum
    kp.wheel_radius = d["kin"]["wheel_radius"]
    kp.roller_angles = list(d["kin"]["roller_angles"])
    kp.base2wheel_positions = [np.asarray(p, float) for p in d["kin"]["base2wheel_positions"]]
    kp.base2wheel_angles = list(d["kin"]["base2wheel_angles"])
    ji, ai = w.JointIndex(), w.ActuatorIndex()
    ji.virtual_start, ji.mobi_start, ji.mani_start = d["joint_idx"]["virtual_start"], d["joint_idx"]["mobi_start"], d["joint_idx"]["mani_start"]
    ai.mobi_start, ai.mani_start = d["actuator_idx"]["mobi_start"], d["actuator_idx"]["mani_start"]
    rd = w.MobileManipulatorRobotData(kp, ji, ai, d["urdf"], d["srdf"], "")
    rc = w.MobileManipulatorRobotController(0.001, rd)
    f, wn = o.frame_id(LINK), o.w
    q, qd, q_t, xd = moma_workload(o.model, wn, 2, 91)
    J, bv = o.mobile_state(q[:, 3:3 + wn], qd[:, 3:3 + wn])
    c, s = np.cos(q[:, 2]), np.sin(q[:, 2])
    qd[:, 0], qd[:, 1], qd[:, 2] = c * bv[:, 0] - s * bv[:, 1], s * bv[:, 0] + c * bv[:, 1], bv[:, 2]
    x_t = o.update_state(q_t, qd, f)["pose"]
    assert rd.updateState(q[0, :3], q[0, 3:3 + wn], q[0, 3 + wn:], qd[0, :3], qd[0, 3:3 + wn], qd[0, 3 + wn:]) is True
    assert rd.getDof() == o.nv and rd.getActuatordDof() == o.act and rd.getManipulatorDof() == 7 and rd.getMobileDof() == wn
    st, full = o.moma_update_state(q[:1], qd[:1], f), o.update_state(q[:1], qd[:1], f)
    rel = lambda a, b: np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
    assert rel(rd.getJacobianActuated(LINK), st["J"][0]) < 1e-12 and rel(rd.getMassMatrixActuated(), st["M"][0]) < 1e-9
    assert rel(rd.getJacobian(LINK), full["J"][0]) < 1e-12 and rel(rd.getMassMatrix(), full["M"][0]) < 1e-9
    assert np.abs(rd.getSelectionMatrix() - st["S"][0]).max() < 1e-12
    assert np.abs(rd.getMobileBaseVel() - bv[0]).max() < 1e-12 and rd.getMobileFKJacobian().shape == (3, wn)
    assert abs(rd.getManipulability(True, True, LINK).manipulability - st["mani"][0]) < 1e-11
    o.set_task_gains(np.full(6, 400.0), np.full(6, 40.0))     # the controller's defaults (robot_controller.cpp:15-16)
    ref, ref3 = o.moma_cycle(1, q[:1], qd[:1], x_t[:1], xd[:1], f), o.moma_cycle(3, q[:1], qd[:1], x_t[:1], xd[:1], f)
    mob, mani = rc.QPIKStep(c_oracle.pose44(x_t[0]), xd[0], LINK)
    assert mob.shape == (wn,) and mani.shape == (7,) and np.abs(np.concatenate([mob, mani]) - ref["out"][0]).max() < 1e-4
    acc, tau = rc.QPIDStep(c_oracle.pose44(x_t[0]), xd[0], LINK)
    assert np.abs(tau - ref3["out"][0, wn:]).max() < 1e-4 * max(1.0, np.abs(ref3["out"]).max())
    assert np.abs(acc - ref3["out2"][0, :wn]).max() < 1e-4 * max(1.0, np.abs(ref3["out2"]).max())
    # stateless twin leaves the cache alone
    M1 = rd.computeMassMatrix(q[1, :3], q[1, 3:3 + wn], q[1, 3 + wn:])
    assert rel(M1, o.update_state(q[1:2], qd[1:2], f)["M"][0]) < 1e-9 and rel(rd.getMassMatrix(), full["M"][0]) < 1e-9


def test_mobile_classes():
    w = wrapper()
    from oracle import c_oracle
    kp = w.KinematicParam()
    kp.type = w.DriveType.Differential
    kp.wheel_radius, kp.base_width, kp.max_lin_speed, kp.max_ang_speed = 0.1651, 0.555, 1.0, 1.0
    rd = w.MobileRobotData(kp)
    rc = w.MobileRobotController(0.001, rd)
    assert rd.getWheelNum() == 2
    rd.updateState(np.zeros(2), np.array([1.0, 2.0]))
    kin = dict(type="Differential", wheel_radius=0.1651, base_width=0.555, max_lin_speed=1.0, max_ang_speed=1.0)
    Jr, bvr = c_oracle.mobile_base(kin, True, np.zeros((1, 2)), np.array([[1.0, 2.0]]))
    assert np.abs(rd.getFKJacobian() - Jr[0]).max() < 1e-14 and np.abs(rd.getBaseVel() - bvr[0]).max() < 1e-14
    Ji, wv = c_oracle.mobile_base(kin, False, np.zeros((1, 2)), np.array([[3.0, 0.0, 2.0]]), saturate=True)
    assert np.abs(rc.computeIKJacobian() - Ji[0]).max() < 1e-14 and np.abs(rc.VelocityCommand(np.array([3.0, 0.0, 2.0])) - wv[0]).max() < 1e-12
