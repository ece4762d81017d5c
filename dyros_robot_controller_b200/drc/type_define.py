"""Result records with the reference's field names (include/dyros_robot_controller/type_define.h:13-171, exposed to
Python in src/bindings.cpp:235-278)."""
from dataclasses import dataclass
from enum import IntEnum
from typing import List, Optional

import numpy as np


@dataclass
class MinDistResult:
    distance: np.ndarray
    grad: np.ndarray
    grad_dot: np.ndarray


@dataclass
class ManipulabilityResult:
    manipulability: np.ndarray
    grad: np.ndarray
    grad_dot: np.ndarray


class DriveType(IntEnum):
    Differential = 0
    Mecanum = 1
    Caster = 2


class KinematicParam:
    """Mobile::KinematicParam (type_define.h:58-72) with the constructor of the reference's Python wrapper
    (drc/type_define.py:11-51): same argument names, defaults and per-type requirements."""

    def __init__(self, type, wheel_radius: float, max_lin_speed: float = 2.0, max_ang_speed: float = 2.0,
                 max_lin_acc: float = 2.0, max_ang_acc: float = 2.0, base_width: Optional[float] = None,
                 roller_angles: Optional[List] = None, base2wheel_positions: Optional[List] = None,
                 base2wheel_angles: Optional[List] = None, wheel_offset: Optional[float] = None) -> None:
        self.type = DriveType(int(type))
        self.wheel_radius = float(wheel_radius)
        self.max_lin_speed, self.max_ang_speed = float(max_lin_speed), float(max_ang_speed)
        self.max_lin_acc, self.max_ang_acc = float(max_lin_acc), float(max_ang_acc)
        self.base_width = self.wheel_offset = 0.0
        self.roller_angles, self.base2wheel_positions, self.base2wheel_angles = [], [], []
        if self.type == DriveType.Differential:
            assert base_width is not None
            self.base_width = float(base_width)
        elif self.type == DriveType.Mecanum:
            assert roller_angles is not None and base2wheel_angles is not None and base2wheel_positions is not None
            self.roller_angles = [float(a) for a in roller_angles]
            self.base2wheel_angles = [float(a) for a in base2wheel_angles]
            self.base2wheel_positions = [np.asarray(p, np.float64) for p in base2wheel_positions]
        elif self.type == DriveType.Caster:
            assert wheel_offset is not None and base2wheel_positions is not None
            self.wheel_offset = float(wheel_offset)
            self.base2wheel_positions = [np.asarray(p, np.float64) for p in base2wheel_positions]

    def cpp(self):
        """the reference returns the wrapped extension-module struct here (drc/type_define.py:53-54); this record IS that struct"""
        return self

    def as_dict(self) -> dict:
        return dict(type=int(self.type), wheel_radius=self.wheel_radius, base_width=self.base_width,
                    wheel_offset=self.wheel_offset, max_lin_speed=self.max_lin_speed, max_ang_speed=self.max_ang_speed,
                    max_lin_acc=self.max_lin_acc, max_ang_acc=self.max_ang_acc, roller_angles=self.roller_angles,
                    base2wheel_positions=[tuple(p[:2]) for p in self.base2wheel_positions],
                    base2wheel_angles=self.base2wheel_angles)


class JointIndex:
    """type_define.h:151-160"""

    def __init__(self, virtual_start: int, mani_start: int, mobi_start: int) -> None:
        self.virtual_start, self.mani_start, self.mobi_start = int(virtual_start), int(mani_start), int(mobi_start)

    def cpp(self):
        return self

    def as_dict(self) -> dict:
        return dict(virtual_start=self.virtual_start, mani_start=self.mani_start, mobi_start=self.mobi_start)


class ActuatorIndex:
    """type_define.h:162-171"""

    def __init__(self, mani_start: int, mobi_start: int) -> None:
        self.mani_start, self.mobi_start = int(mani_start), int(mobi_start)

    def cpp(self):
        return self

    def as_dict(self) -> dict:
        return dict(mani_start=self.mani_start, mobi_start=self.mobi_start)
