"""MobileManipulator RobotController with the reference's interface (reference drc/mobile_manipulator/robot_controller.py
wrapping src/mobile_manipulator/robot_controller.cpp:147-250): whole-body QPIK / QPID returning the reference's
(mobile, manipulator) pairs."""
from __future__ import annotations

import sys

import numpy as np

from .robot_data import RobotData


class RobotController:
    def __init__(self, dt: float, robot_data: RobotData):
        self._dt = float(dt)
        self._robot_data = robot_data
        self._ctx = robot_data._ctx

    def set_task_gain(self, kp, kv):
        kp, kv = np.asarray(kp, np.float64).ravel(), np.asarray(kv, np.float64).ravel()
        if kp.size != 6 or kv.size != 6:
            raise RuntimeError("task gain size mismatch: expected 6")
        self._ctx.set_params(Kp_task=kp, Kv_task=kv)

    def _fid(self, link_name):
        fid = self._robot_data._model.frame_id(link_name)
        if fid < 0:
            raise KeyError(f"Link name {link_name} not found in URDF.")
        return fid

    def _emit(self, r, what, first_key="out"):
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
