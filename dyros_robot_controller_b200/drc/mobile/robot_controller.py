"""Mobile RobotController with the reference's interface (reference drc/mobile/robot_controller.py wrapping
src/mobile/robot_controller.cpp): wheel velocities from a desired base velocity through the inverse-kinematics Jacobian
of the drive type.  Single-base arrays as in the reference (base_vel: (3,)), or a leading batch axis."""
from __future__ import annotations

import numpy as np

from .robot_data import RobotData


class RobotController:
    def __init__(self, dt: float, robot_data: RobotData):
        self._dt = float(dt)
        self._robot_data = robot_data     # kept alive like the reference wrapper does (drc/mobile/robot_controller.py:27-28)
        self._base = robot_data._base

    def _wheel_pos(self, B):
        wp = self._robot_data._wheel_pos  # caster: the steering angles of the last update_state (robot_controller.cpp:107)
        if wp.shape[0] == B:
            return wp
        if wp.shape[0] == 1:
            return np.repeat(wp, B, axis=0)
        raise ValueError(f"robot_data holds {wp.shape[0]} base states, the command has {B}")

    def compute_wheel_vel(self, base_vel) -> np.ndarray:            # robot_controller.cpp:43-47
        single = np.ndim(base_vel) == 1
        v = np.atleast_2d(np.asarray(base_vel, np.float64))
        _, wv = self._base.ik(self._wheel_pos(v.shape[0]), v, saturate=False, want_J=False)
        return wv[0] if single else wv

    def compute_IK_jacobian(self) -> np.ndarray:                   # robot_controller.cpp:50-63
        wp = self._robot_data._wheel_pos
        J, _ = self._base.ik(wp, None)
        return J[0] if self._robot_data._single else J

    def velocity_command(self, desired_base_vel) -> np.ndarray:     # robot_controller.cpp:14-41
        single = np.ndim(desired_base_vel) == 1
        v = np.atleast_2d(np.asarray(desired_base_vel, np.float64))
        _, wv = self._base.ik(self._wheel_pos(v.shape[0]), v, saturate=True, want_J=False)
        return wv[0] if single else wv
