"""MobileManipulator RobotData with the reference's interface (reference drc/mobile_manipulator/robot_data.py wrapping
src/mobile_manipulator/robot_data.cpp) on top of the batched engine: six state vectors in, actuated-space quantities
out.  Single-robot arrays as in the reference, or a leading batch axis."""
from __future__ import annotations

import sys

import numpy as np

from ... import engine
from ..type_define import ManipulabilityResult


class RobotData:
    def __init__(self, mobile_param: dict, joint_idx: dict, actuator_idx: dict, urdf_path: str, srdf_path: str = "",
                 packages_path: str = "", max_batch: int = 1, device: int = 0):
        as_dict = lambda a: a.as_dict() if hasattr(a, "as_dict") else a   # KinematicParam / JointIndex / ActuatorIndex or dicts
        mobile_param, joint_idx, actuator_idx = as_dict(mobile_param), as_dict(joint_idx), as_dict(actuator_idx)
        self._model = engine.Model(urdf_path, srdf_path, packages_path).attach_mobile_base(mobile_param, joint_idx, actuator_idx)
        self._ctx = engine.Context(self._model, max_batch, device)
        m = self._model
        self._ji, self._ai = m.joint_idx, m.actuator_idx
        self._w, self._k, self._act, self._dof = m.wheel_num, m.mani_dof, m.actuated_dof, m.dof
        self._single = True
        # base Jacobian: a constant for differential / mecanum drives, a function of the steering angles for casters
        self._caster = m.drive_type == 2
        self._base = engine.MobileBase(mobile_param, device) if self._caster else None
        self._J_mobile = self._base.fk(np.zeros((1, self._w)), None)[0][0] if self._caster else m.base_jacobian()

    # ---- vector assembly (robot_data.cpp:417-437)
    def get_joint_vector(self, q_virtual, q_mobile, q_mani) -> np.ndarray:
        qv, qm, qa = (np.atleast_2d(np.asarray(a, np.float64)) for a in (q_virtual, q_mobile, q_mani))
        q = np.zeros((qv.shape[0], self._dof))
        q[:, self._ji["virtual_start"]:self._ji["virtual_start"] + 3] = qv
        q[:, self._ji["mobi_start"]:self._ji["mobi_start"] + self._w] = qm
        q[:, self._ji["mani_start"]:self._ji["mani_start"] + self._k] = qa
        return q

    def get_actuator_vector(self, q_mobile, q_mani) -> np.ndarray:
        qm, qa = (np.atleast_2d(np.asarray(a, np.float64)) for a in (q_mobile, q_mani))
        q = np.zeros((qm.shape[0], self._act))
        q[:, self._ai["mobi_start"]:self._ai["mobi_start"] + self._w] = qm
        q[:, self._ai["mani_start"]:self._ai["mani_start"] + self._k] = qa
        return q

    def split_actuated(self, v):
        """(mobile, manipulator) segments of an actuated vector (ActuatorIndex split, robot_controller.cpp:162-165)."""
        v = np.asarray(v)
        return (v[..., self._ai["mobi_start"]:self._ai["mobi_start"] + self._w],
                v[..., self._ai["mani_start"]:self._ai["mani_start"] + self._k])

    def _sq(self, a):
        return a[0] if self._single else a

    # ---- state update (robot_data.cpp:83-144)
    def update_state(self, q_virtual, q_mobile, q_mani, qdot_virtual, qdot_mobile, qdot_mani) -> bool:
        self._single = np.ndim(q_mani) == 1
        self._q = self.get_joint_vector(q_virtual, q_mobile, q_mani)
        self._qdot = self.get_joint_vector(qdot_virtual, qdot_mobile, qdot_mani)
        self._q_act = self.get_actuator_vector(q_mobile, q_mani)
        self._qdot_act = self.get_actuator_vector(qdot_mobile, qdot_mani)
        self._wheel_vel = np.atleast_2d(np.asarray(qdot_mobile, np.float64))
        if self._caster:   # Mobile::RobotData::updateState (mobile/robot_data.cpp:110)
            self._J_mobile = self._base.fk(np.atleast_2d(np.asarray(q_mobile, np.float64)), None)[0]
        return bool(self._ctx.moma_update_state(self._q, self._qdot))

    def get_dof(self) -> int:
        return self._dof

    def get_actuator_dof(self) -> int:
        return self._act

    def get_manipulator_dof(self) -> int:
        return self._k

    def get_mobile_dof(self) -> int:
        return self._w

    def get_joint_index(self) -> dict:
        return dict(self._ji)

    def get_actuator_index(self) -> dict:
        return dict(self._ai)

    def get_joint_position(self):
        return self._sq(self._q)

    def get_joint_velocity(self):
        return self._sq(self._qdot)

    def get_joint_position_actuated(self):
        return self._sq(self._q_act)

    def get_joint_velocity_actuated(self):
        return self._sq(self._qdot_act)

    def get_FK_jacobian(self) -> np.ndarray:          # Mobile::RobotData::getFKJacobian
        return self._sq(self._J_mobile).copy() if self._J_mobile.ndim == 3 else self._J_mobile.copy()

    def get_base_vel(self) -> np.ndarray:             # Mobile::RobotData::getBaseVel = J_mobile * wheel_vel
        if self._J_mobile.ndim == 3:
            return self._sq(np.einsum("brk,bk->br", self._J_mobile, self._wheel_vel))
        return self._sq(self._wheel_vel @ self._J_mobile.T)

    def _get(self, link_name, key, neutral=None):
        fid = self._model.frame_id(link_name) if link_name is not None else 0
        if fid < 0:
            print(f"\033[1;31mError: Link name {link_name} not found in URDF.\033[0m", file=sys.stderr)
            return neutral
        v = self._ctx.moma_get_state(fid, want=(key,))[key]
        if key == "pose":
            v = engine.pose44(v)
        return self._sq(v)

    def get_pose(self, link_name: str):
        return self._get(link_name, "pose", np.eye(4))

    def get_velocity(self, link_name: str):
        return self._get(link_name, "vel", np.zeros(6))

    def get_jacobian_actuated(self, link_name: str):
        return self._get(link_name, "J", np.zeros((6, self._act)))

    def get_jacobian_actuated_time_variation(self, link_name: str):
        return self._get(link_name, "Jdot", np.zeros((6, self._act)))

    def get_mass_matrix_actuated(self):
        return self._get(None, "M")

    def get_mass_matrix_actuated_inv(self):
        return self._get(None, "Minv")

    def get_gravity_actuated(self):
        return self._get(None, "g")

    def get_nonlinear_effects_actuated(self):
        return self._get(None, "nle")

    def get_coriolis_actuated(self):
        return self.get_nonlinear_effects_actuated() - self.get_gravity_actuated()

    def get_manipulability(self, with_grad: bool, with_graddot: bool, link_name: str) -> ManipulabilityResult:
        fid = self._model.frame_id(link_name)
        if fid < 0:
            print(f"\033[1;31mError: Link name {link_name} not found in URDF.\033[0m", file=sys.stderr)
            return ManipulabilityResult(0.0, np.zeros(self._k), np.zeros(self._k))
        r = self._ctx.moma_get_state(fid, want=("mani", "mani_grad", "mani_graddot"))
        z = np.zeros_like(r["mani_grad"])
        return ManipulabilityResult(self._sq(r["mani"]), self._sq(r["mani_grad"] if (with_grad or with_graddot) else z),
                                    self._sq(r["mani_graddot"] if with_graddot else z))
