"""Mobile RobotData with the reference's interface (reference drc/mobile/robot_data.py wrapping
src/mobile/robot_data.cpp) on top of the batched engine.  Single-base arrays as in the reference (wheel_pos: (w,)), or a
leading batch axis (wheel_pos: (B, w)); results carry the same leading axis."""
from __future__ import annotations

import numpy as np

from ... import engine
from ..type_define import KinematicParam


class RobotData:
    def __init__(self, param, device: int = 0):
        self._param = param
        self._kin = param.as_dict() if isinstance(param, KinematicParam) else dict(param)
        self._base = engine.MobileBase(self._kin, device)
        self._w = self._base.wheel_num
        self._single = True
        z = np.zeros((1, self._w))
        self._wheel_pos, self._wheel_vel = z.copy(), z.copy()
        # mobile/robot_data.cpp:21-29: the constructor evaluates the Jacobian at zero wheel angles
        self._J, self._base_vel = self._base.fk(self._wheel_pos, self._wheel_vel)

    def _sq(self, a):
        return a[0] if self._single else a

    def get_verbose(self) -> str:                    # mobile/robot_data.cpp:34-101
        k = self._kin
        t = ["Differential", "Mecanum", "Caster"][self._base.drive_type]
        return (f"type: {t}\nwheel_num: {self._w}\nwheel_radius: {k.get('wheel_radius', 0.0)}\n"
                f"max_lin_speed: {k.get('max_lin_speed', 0.0)}\nmax_ang_speed: {k.get('max_ang_speed', 0.0)}\n")

    def update_state(self, wheel_pos, wheel_vel) -> bool:   # mobile/robot_data.cpp:103-114
        self._single = np.ndim(wheel_pos) == 1
        self._wheel_pos = np.atleast_2d(np.asarray(wheel_pos, np.float64))
        self._wheel_vel = np.atleast_2d(np.asarray(wheel_vel, np.float64))
        if self._wheel_pos.shape[-1] != self._w or self._wheel_vel.shape != self._wheel_pos.shape:
            raise ValueError(f"wheel_pos / wheel_vel must hold {self._w} entries per base")
        self._J, self._base_vel = self._base.fk(self._wheel_pos, self._wheel_vel)
        return True

    def compute_base_vel(self, wheel_pos, wheel_vel) -> np.ndarray:   # :116-120 (stateless)
        single = np.ndim(wheel_pos) == 1
        _, bv = self._base.fk(np.atleast_2d(wheel_pos), np.atleast_2d(wheel_vel), want_J=False)
        return bv[0] if single else bv

    def compute_fk_jacobian(self, wheel_pos) -> np.ndarray:           # :122-136 (stateless)
        single = np.ndim(wheel_pos) == 1
        J, _ = self._base.fk(np.atleast_2d(wheel_pos), None)
        return J[0] if single else J

    def get_kine_param(self):
        return self._param

    def get_wheel_num(self) -> int:
        return self._w

    def get_wheel_pos(self) -> np.ndarray:
        return self._sq(self._wheel_pos)

    def get_wheel_vel(self) -> np.ndarray:
        return self._sq(self._wheel_vel)

    def get_base_vel(self) -> np.ndarray:
        return self._sq(self._base_vel)

    def get_FK_jacobian(self) -> np.ndarray:
        return self._sq(self._J)
