"""Result records with the reference's field names (include/dyros_robot_controller/type_define.h:13-171, exposed to
Python in src/bindings.cpp:235-278)."""
from dataclasses import dataclass

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
