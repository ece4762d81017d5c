"""MobileManipulator RobotController with the reference's interface (reference drc/mobile_manipulator/robot_controller.py
wrapping src/mobile_manipulator/robot_controller.cpp:7-250): manipulator joint-space helpers and the whole-body QPIK / QPID,
returning the reference's (mobile, manipulator) pairs.  Default gains are the reference's: task Kp 400 / Kv 40, manipulator
joint Kp 400 / Kv 40 (robot_controller.cpp:15-18)."""
from __future__ import annotations

import sys

import numpy as np

from ..manipulator.robot_controller import _cubic
from .robot_data import RobotData


class RobotController:
    def __init__(self, dt: float, robot_data: RobotData):
        self._dt = float(dt)          # stored, never used by the reference either (robot_controller.cpp:7-22)
        self._robot_data = robot_data
        self._ctx = robot_data._ctx
        k = robot_data.get_manipulator_dof()
        self._kp_t, self._kv_t = np.full(6, 400.0), np.full(6, 40.0)
        self._kp_j, self._kv_j = np.full(k, 400.0), np.full(k, 40.0)
        self._ctx.set_params(Kp_task=self._kp_t, Kv_task=self._kv_t)

    # ---- gains (robot_controller.cpp:24-76); size mismatch raises like the reference's std::runtime_error
    @staticmethod
    def _check(v, n, what):
        v = np.asarray(v, np.float64).ravel()
        if v.size != n:
            raise RuntimeError(f"{what} must be of size {n}.")
        return v

    def set_manipulator_joint_gain(self, kp, kv):
        k = self._robot_data.get_manipulator_dof()
        self._kp_j, self._kv_j = self._check(kp, k, "Kp and Kv"), self._check(kv, k, "Kp and Kv")

    def set_manipulator_joint_kp_gain(self, kp):
        self._kp_j = self._check(kp, self._robot_data.get_manipulator_dof(), "Kp")

    def set_manipulator_joint_kv_gain(self, kv):
        self._kv_j = self._check(kv, self._robot_data.get_manipulator_dof(), "Kv")

    def set_task_gain(self, kp, kv):
        self._kp_t, self._kv_t = self._check(kp, 6, "Kp and Kv"), self._check(kv, 6, "Kp and Kv")
        self._ctx.set_params(Kp_task=self._kp_t, Kv_task=self._kv_t)

    def set_task_kp_gain(self, kp):
        self._kp_t = self._check(kp, 6, "Kp")
        self._ctx.set_params(Kp_task=self._kp_t)

    def set_task_kv_gain(self, kv):
        # the reference's Python wrapper calls setTaskKpGain here (drc/mobile_manipulator/robot_controller.py:86, a bug); the C++
        # class sets Kv (robot_controller.cpp:69-76) -- this mirror follows the C++ class
        self._kv_t = self._check(kv, 6, "Kv")
        self._ctx.set_params(Kv_task=self._kv_t)

    # ---- manipulator joint space (robot_controller.cpp:78-145)
    def move_manipulator_joint_position_cubic(self, q_mani_target, qdot_mani_target, q_mani_init, qdot_mani_init, current_time,
                                              init_time, duration):
        return _cubic(current_time, init_time, init_time + duration, q_mani_init, q_mani_target, qdot_mani_init, qdot_mani_target)[0]

    def move_manipulator_joint_torque_step(self, q_mani_target=None, qdot_mani_target=None, qddot_mani_target=None):
        """the reference wrapper's three-keyword form (drc/mobile_manipulator/robot_controller.py:128-151): qddot_mani_target ->
        M_mani qddot + g_mani (robot_controller.cpp:104-110); (q_mani_target, qdot_mani_target) -> PD acceleration first (:112-119).
        A single positional argument is taken as qddot_mani_target, like the C++ overload."""
        rd = self._robot_data
        ms, k = rd._ji["mani_start"], rd.get_manipulator_dof()
        if qddot_mani_target is None and qdot_mani_target is None and q_mani_target is not None:
            qddot_mani_target, q_mani_target = q_mani_target, None
        if qddot_mani_target is not None:
            qdd = np.asarray(qddot_mani_target, np.float64)
        elif q_mani_target is not None and qdot_mani_target is not None:
            qdd = self._kp_j * (np.asarray(q_mani_target, np.float64) - rd.get_manipulator_joint_position()) + \
                  self._kv_j * (np.asarray(qdot_mani_target, np.float64) - rd.get_manipulator_joint_velocity())
        else:
            return None
        M, g = rd.get_mass_matrix(), rd.get_gravity()
        return np.einsum("...ij,...j->...i", M[..., ms:ms + k, ms:ms + k], qdd) + g[..., ms:ms + k]

    def move_manipulator_joint_torque_cubic(self, q_mani_target, qdot_mani_target, q_mani_init, qdot_mani_init, current_time,
                                            init_time, duration):
        q_des, qd_des = _cubic(current_time, init_time, init_time + duration, q_mani_init, q_mani_target, qdot_mani_init, qdot_mani_target)
        return self.move_manipulator_joint_torque_step(q_des, qd_des)

    # ---- whole-body QPs (robot_controller.cpp:147-250)
    def _fid(self, link_name):
        fid = self._robot_data._model.frame_id(link_name)
        if fid < 0:
            raise KeyError(f"Link name {link_name} not found in URDF.")
        return fid

    def _emit(self, r, what):
        bad = int((r["status"] != 1).sum())
        if bad:
            print(f"QP {what} failed to compute optimal solution for {bad} robot(s)", file=sys.stderr)
        self.last_status, self.last_iters = r["status"], r["iters"]
        sq = self._robot_data._sq
        if what == "IK":       # (opt_qdot_mobile, opt_qdot_manipulator)
            mob, mani = self._robot_data.split_actuated(r["out"])
            return sq(mob), sq(mani)
        mob, _ = self._robot_data.split_actuated(r["etadot"])   # (opt_qddot_mobile, opt_torque_manipulator)
        _, tau = self._robot_data.split_actuated(r["out"])
        return sq(mob), sq(tau)

    def QPIK(self, xdot_target, link_name: str):
        return self._emit(self._ctx.moma_qpik(xdot_target, self._fid(link_name)), "IK")

    def QPIK_step(self, x_target, xdot_target, link_name: str):
        return self._emit(self._ctx.moma_qpik_step(x_target, xdot_target, self._fid(link_name)), "IK")

    def QPIK_cubic(self, x_target, xdot_target, x_init, xdot_init, current_time, init_time, duration, link_name: str):
        x_des, xd_des = self._ctx.task_space_cubic(x_target, xdot_target, x_init, xdot_init, current_time, init_time, duration)
        return self.QPIK_step(x_des, xd_des, link_name)

    def QPID(self, xddot_target, link_name: str):
        return self._emit(self._ctx.moma_qpid(xddot_target, self._fid(link_name)), "ID")

    def QPID_step(self, x_target, xdot_target, link_name: str):
        return self._emit(self._ctx.moma_qpid_step(x_target, xdot_target, self._fid(link_name)), "ID")

    def QPID_cubic(self, x_target, xdot_target, x_init, xdot_init, current_time, init_time, duration, link_name: str):
        x_des, xd_des = self._ctx.task_space_cubic(x_target, xdot_target, x_init, xdot_init, current_time, init_time, duration)
        return self.QPID_step(x_des, xd_des, link_name)
