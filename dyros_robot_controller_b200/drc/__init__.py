"""Host-side mirror of the reference's Python package `drc` (reference drc/__init__.py:1-4) for the hot path:
same class and method names, same argument order and return shapes, backed by the batched B200 engine through the
C ABI (include/drc_b200.h).  Every method accepts the reference's single-robot arrays (q: (n,), poses: (4,4)) and, as
the batched sibling, a leading batch axis (q: (B,n), poses: (B,4,4)); results carry the same leading axis.

    from dyros_robot_controller_b200.drc.manipulator import RobotData, RobotController
"""
from . import manipulator, mobile, mobile_manipulator  # noqa: F401
from .type_define import (ActuatorIndex, DriveType, JointIndex, KinematicParam, ManipulabilityResult,  # noqa: F401
                          MinDistResult)
