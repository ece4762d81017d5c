from .robot_controller import RobotController  # noqa: F401
from .robot_data import RobotData  # noqa: F401
