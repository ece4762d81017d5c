"""Manipulator RobotController with the reference's interface (reference drc/manipulator/robot_controller.py:6-517
wrapping src/manipulator/robot_controller.cpp) on top of the batched engine.  All methods act on the state cached
by RobotData.update_state, as in the reference."""
from __future__ import annotations

import sys

import numpy as np

from ... import engine
from .robot_data import RobotData


def _cubic(t, t0, tf, x0, xf, v0, vf):
    """DyrosMath::cubic / cubicDot (include/math_type_define.h:62-144) on arrays."""
    x0, xf, v0, vf = (np.asarray(a, np.float64) for a in (x0, xf, v0, vf))
    if t < t0:
        return x0, v0
    if t > tf:
        return xf, vf
    e, T = t - t0, tf - t0
    a2 = 3 * (xf - x0) / T ** 2 - 2 * v0 / T - vf / T
    a3 = -2 * (xf - x0) / T ** 3 + (v0 + vf) / T ** 2
    return x0 + v0 * e + a2 * e ** 2 + a3 * e ** 3, v0 + 2 * a2 * e + 3 * a3 * e ** 2


class RobotController:
    def __init__(self, dt: float, robot_data: RobotData):
        self._dt = float(dt)          # stored, never used by the reference either (robot_controller.cpp:7-19)
        self._robot_data = robot_data
        self._ctx = robot_data._ctx
        n = robot_data.get_dof()
        p = self._ctx.get_params()    # defaults: Kp 400 / Kv 40 joint, Kp 100 / Kv 20 task (robot_controller.cpp:12-15)
        self._kp_j, self._kv_j = np.array(p.Kp_joint[:n]), np.array(p.Kv_joint[:n])
        self._kp_t, self._kv_t = np.array(p.Kp_task[:]), np.array(p.Kv_task[:])

    # ---- gains (robot_controller.cpp:21-63); size mismatch raises like the reference's std::runtime_error
    def _check(self, v, n, what):
        v = np.asarray(v, np.float64).ravel()
        if v.size != n:
            raise RuntimeError(f"{what} size mismatch: expected {n}, got {v.size}")
        return v

    def _push(self):
        self._ctx.set_params(Kp_joint=self._kp_j, Kv_joint=self._kv_j, Kp_task=self._kp_t, Kv_task=self._kv_t)

    def set_joint_gain(self, kp, kv):
        n = self._robot_data.get_dof()
        self._kp_j, self._kv_j = self._check(kp, n, "Kp"), self._check(kv, n, "Kv")
        self._push()

    def set_joint_kp_gain(self, kp):
        self._kp_j = self._check(kp, self._robot_data.get_dof(), "Kp")
        self._push()

    def set_joint_kv_gain(self, kv):
        self._kv_j = self._check(kv, self._robot_data.get_dof(), "Kv")
        self._push()

    def set_task_gain(self, kp, kv):
        self._kp_t, self._kv_t = self._check(kp, 6, "Kp"), self._check(kv, 6, "Kv")
        self._push()

    def set_task_kp_gain(self, kp):
        self._kp_t = self._check(kp, 6, "Kp")
        self._push()

    def set_task_kv_gain(self, kv):
        # the reference's Python wrapper calls setTaskKpGain here (drc/manipulator/robot_controller.py:86, a bug);
        # the C++ class sets Kv (robot_controller.cpp:58-63) -- this mirror follows the C++ class
        self._kv_t = self._check(kv, 6, "Kv")
        self._push()

    # ---- plumbing
    def _sq(self, a):
        return a[0] if self._robot_data._single else a

    def _fid(self, link_name):
        fid = self._robot_data._model.frame_id(link_name)
        if fid < 0:
            raise KeyError(f"Link name {link_name} not found in URDF.")
        return fid

    # ---- joint space (robot_controller.cpp:65-154)
    def move_joint_position_cubic(self, q_target, qdot_target, q_init, qdot_init, current_time, init_time, duration):
        return _cubic(current_time, init_time, init_time + duration, q_init, q_target, qdot_init, qdot_target)[0]

    def move_joint_velocity_cubic(self, q_target, qdot_target, q_init, qdot_init, current_time, init_time, duration):
        return _cubic(current_time, init_time, init_time + duration, q_init, q_target, qdot_init, qdot_target)[1]

    def move_joint_torque_step(self, q_target=None, qdot_target=None, qddot_target=None):
        """the reference wrapper's three-keyword form (drc/manipulator/robot_controller.py:161-184): qddot_target -> M qddot + g
        (robot_controller.cpp:108-113); (q_target, qdot_target) -> PD acceleration first (:115-125).  A single positional
        argument is taken as qddot_target, like the C++ overload."""
        if qddot_target is None and qdot_target is None and q_target is not None:
            qddot_target, q_target = q_target, None
        if qddot_target is not None:
            qdd = np.asarray(qddot_target, np.float64)
            M, g = self._robot_data.get_mass_matrix(), self._robot_data.get_gravity()
            return np.einsum("...ij,...j->...i", M, qdd) + g
        if q_target is not None and qdot_target is not None:
            return self._sq(self._ctx.joint_torque_step(q_target, qdot_target))
        return None

    def move_joint_torque_cubic(self, q_target, qdot_target, q_init, qdot_init, current_time, init_time, duration):
        q_des, qd_des = _cubic(current_time, init_time, init_time + duration, q_init, q_target, qdot_init, qdot_target)
        return self.move_joint_torque_step(q_des, qd_des)

    # ---- task space without QP (robot_controller.cpp:156-247)
    def CLIK_step(self, x_target, xdot_target, link_name: str, null_qdot=None):
        return self._sq(self._ctx.clik_step(x_target, xdot_target, self._fid(link_name), null_qdot=null_qdot))

    def OSF(self, xddot_target, link_name: str, null_torque=None):
        return self._sq(self._ctx.osf(xddot_target, self._fid(link_name), null_torque=null_torque))

    def OSF_step(self, x_target, xdot_target, link_name: str, null_torque=None):
        return self._sq(self._ctx.osf_step(x_target, xdot_target, self._fid(link_name), null_torque=null_torque))

    def _cubic_task(self, x_target, xdot_target, x_init, xdot_init, current_time, init_time, duration):
        x_des, xd_des = self._ctx.task_space_cubic(x_target, xdot_target, x_init, xdot_init, current_time, init_time, duration)
        return x_des, xd_des

    def CLIK_cubic(self, x_target, xdot_target, x_init, xdot_init, current_time, init_time, duration, link_name: str,
                   null_qdot=None):
        x_des, xd_des = self._cubic_task(x_target, xdot_target, x_init, xdot_init, current_time, init_time, duration)
        return self.CLIK_step(x_des, xd_des, link_name, null_qdot)

    def OSF_cubic(self, x_target, xdot_target, x_init, xdot_init, current_time, init_time, duration, link_name: str,
                  null_torque=None):
        x_des, xd_des = self._cubic_task(x_target, xdot_target, x_init, xdot_init, current_time, init_time, duration)
        return self.OSF_step(x_des, xd_des, link_name, null_torque)

    # ---- QP controllers (robot_controller.cpp:277-360); failures print like the reference and return its fallback
    def _report(self, r, what):
        bad = int((r["status"] != 1).sum())
        if bad:
            print(f"QP {what} failed to compute optimal solution for {bad} robot(s)", file=sys.stderr)
        self.last_status, self.last_iters = r["status"], r["iters"]
        return self._sq(r["out"])

    def QPIK(self, xdot_target, link_name: str):
        return self._report(self._ctx.qpik(xdot_target, self._fid(link_name)), "IK")

    def QPIK_step(self, x_target, xdot_target, link_name: str):
        return self._report(self._ctx.qpik_step(x_target, xdot_target, self._fid(link_name)), "IK")

    def QPIK_cubic(self, x_target, xdot_target, x_init, xdot_init, current_time, init_time, duration, link_name: str):
        x_des, xd_des = self._cubic_task(x_target, xdot_target, x_init, xdot_init, current_time, init_time, duration)
        return self.QPIK_step(x_des, xd_des, link_name)

    def QPID(self, xddot_target, link_name: str):
        return self._report(self._ctx.qpid(xddot_target, self._fid(link_name)), "ID")

    def QPID_step(self, x_target, xdot_target, link_name: str):
        return self._report(self._ctx.qpid_step(x_target, xdot_target, self._fid(link_name)), "ID")

    def QPID_cubic(self, x_target, xdot_target, x_init, xdot_init, current_time, init_time, duration, link_name: str):
        x_des, xd_des = self._cubic_task(x_target, xdot_target, x_init, xdot_init, current_time, init_time, duration)
        return self.QPID_step(x_des, xd_des, link_name)
