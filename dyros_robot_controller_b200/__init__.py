"""drc-b200: B200-native batched control-cycle engine (drop-in for dyros_robot_controller's hot path).

    from dyros_robot_controller_b200 import Model, Context
    model = Model(urdf, srdf); ctx = Context(model, max_batch=65536)
    ctx.cycle_qpik_step(q, qdot, x_target, xdot_target, "fr3_link8")

`dyros_robot_controller_b200.drc` mirrors the reference's Python package (same class / method names).
"""
from pathlib import Path

from .engine import Context, MobileBase, Model, device_count, fp64_peak_tflops, pose12, pose44  # noqa: F401

ROBOTS_DIR = Path(__file__).resolve().parent / "robots"
FR3_URDF = str(ROBOTS_DIR / "fr3" / "fr3.urdf")
FR3_SRDF = str(ROBOTS_DIR / "fr3" / "fr3.srdf")
# synthesized mobile manipulators of BASELINE configs 4-5 (tools/make_moma_urdf.py; NOT from the reference)
HUSKY_FR3_URDF, HUSKY_FR3_SRDF = str(ROBOTS_DIR / "husky_fr3" / "husky_fr3.urdf"), str(ROBOTS_DIR / "husky_fr3" / "husky_fr3.srdf")
XLS_FR3_URDF, XLS_FR3_SRDF = str(ROBOTS_DIR / "xls_fr3" / "xls_fr3.urdf"), str(ROBOTS_DIR / "xls_fr3" / "xls_fr3.srdf")
PCV_FR3_URDF, PCV_FR3_SRDF = str(ROBOTS_DIR / "pcv_fr3" / "pcv_fr3.urdf"), str(ROBOTS_DIR / "pcv_fr3" / "pcv_fr3.srdf")  # powered casters
__version__ = "0.1.0"
