"""MobileManipulator RobotData with the reference's interface (reference drc/mobile_manipulator/robot_data.py:7-833 wrapping
src/mobile_manipulator/robot_data.cpp) on top of the batched engine: six state vectors in, full-dof and actuated-space
quantities out.  Single-robot arrays as in the reference, or a leading batch axis.

Every method of the reference class exists here with the same name, argument order and return shape.  Cached getters read the
device state cache written by update_state; the stateless compute_* twins (robot_data.cpp:146-405) run on a second, private
context so that they do not disturb it.  Reference quirks, decided as in SURVEY.md Appendix A:
  * compute_*_actuated evaluate the full-dof quantity at q_virtual = 0 and multiply by S(q_virtual) -- reproduced as written
    (robot_data.cpp:185-232, 382-405);
  * computeSelectionMatrix sizes the wheel identity block with mani_dof (robot_data.cpp:372, quirk Q8) -- the correct S is built;
  * computeManipulability(with_graddot) passes an unsized matrix to Pinocchio (quirk Q4) -- implemented correctly."""
from __future__ import annotations

import sys
from typing import Tuple

import numpy as np

from ... import engine
from ..type_define import ManipulabilityResult, MinDistResult


class RobotData:
    def __init__(self, mobile_param: dict, joint_idx: dict, actuator_idx: dict, urdf_path: str, srdf_path: str = "",
                 packages_path: str = "", max_batch: int = 1, device: int = 0):
        as_dict = lambda a: a.as_dict() if hasattr(a, "as_dict") else a   # KinematicParam / JointIndex / ActuatorIndex or dicts
        self._param_in, self._ji_in, self._ai_in = mobile_param, joint_idx, actuator_idx
        mobile_param, joint_idx, actuator_idx = as_dict(mobile_param), as_dict(joint_idx), as_dict(actuator_idx)
        self._model = engine.Model(urdf_path, srdf_path, packages_path).attach_mobile_base(mobile_param, joint_idx, actuator_idx)
        self._ctx = engine.Context(self._model, max_batch, device)
        self._scratch = None
        self._max_batch, self._device = max_batch, device
        m = self._model
        self._ji, self._ai = m.joint_idx, m.actuator_idx
        self._w, self._k, self._act, self._dof = m.wheel_num, m.mani_dof, m.actuated_dof, m.dof
        self._single = True
        # base Jacobian: a constant for differential / mecanum drives, a function of the steering angles for casters
        self._caster = m.drive_type == 2
        self._base = engine.MobileBase(mobile_param, device) if self._caster else None
        self._J_const = None if self._caster else m.base_jacobian()
        self._J_mobile = self._base.fk(np.zeros((1, self._w)), None)[0][0] if self._caster else self._J_const
        z = np.zeros((1, self._dof))
        self._q, self._qdot = z.copy(), z.copy()
        self._q_act, self._qdot_act = np.zeros((1, self._act)), np.zeros((1, self._act))
        self._wheel_vel = np.zeros((1, self._w))

    # ---- vector assembly (robot_data.cpp:417-437)
    def get_joint_vector(self, q_virtual, q_mobile, q_mani) -> np.ndarray:
        qv, qm, qa = (np.atleast_2d(np.asarray(a, np.float64)) for a in (q_virtual, q_mobile, q_mani))
        q = np.zeros((qa.shape[0], self._dof))
        q[:, self._ji["virtual_start"]:self._ji["virtual_start"] + 3] = qv
        q[:, self._ji["mobi_start"]:self._ji["mobi_start"] + self._w] = qm
        q[:, self._ji["mani_start"]:self._ji["mani_start"] + self._k] = qa
        return q

    def get_actuator_vector(self, q_mobile, q_mani) -> np.ndarray:
        qm, qa = (np.atleast_2d(np.asarray(a, np.float64)) for a in (q_mobile, q_mani))
        q = np.zeros((qa.shape[0], self._act))
        q[:, self._ai["mobi_start"]:self._ai["mobi_start"] + self._w] = qm
        q[:, self._ai["mani_start"]:self._ai["mani_start"] + self._k] = qa
        return q

    def split_actuated(self, v):
        """(mobile, manipulator) segments of an actuated vector (ActuatorIndex split, robot_controller.cpp:162-165)."""
        v = np.asarray(v)
        return (v[..., self._ai["mobi_start"]:self._ai["mobi_start"] + self._w],
                v[..., self._ai["mani_start"]:self._ai["mani_start"] + self._k])

    def _sq(self, a, single=None):
        return a[0] if (self._single if single is None else single) else a

    def _seg(self, v, start, n):
        return self._sq(v[:, start:start + n])

    def get_verbose(self) -> str:
        return self._model.verbose()

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

    # ---- sizes and indices
    def get_dof(self) -> int:
        return self._dof

    def get_actuator_dof(self) -> int:
        return self._act

    def get_manipulator_dof(self) -> int:
        return self._k

    def get_mobile_dof(self) -> int:
        return self._w

    def get_joint_index(self):
        return self._ji_in if not isinstance(self._ji_in, dict) else dict(self._ji)

    def get_actuator_index(self):
        return self._ai_in if not isinstance(self._ai_in, dict) else dict(self._ai)

    # ---- joint state (robot_data.h:449-520)
    def get_joint_position(self):
        return self._sq(self._q)

    def get_joint_velocity(self):
        return self._sq(self._qdot)

    def get_joint_position_limit(self) -> Tuple[np.ndarray, np.ndarray]:
        return self._model.q_lower.copy(), self._model.q_upper.copy()

    def get_joint_velocity_limit(self) -> Tuple[np.ndarray, np.ndarray]:
        return -self._model.v_limit, self._model.v_limit.copy()

    def get_virtual_joint_position(self):
        return self._seg(self._q, self._ji["virtual_start"], 3)

    def get_mobile_joint_position(self):
        return self._seg(self._q, self._ji["mobi_start"], self._w)

    def get_manipulator_joint_position(self):
        return self._seg(self._q, self._ji["mani_start"], self._k)

    def get_virtual_joint_velocity(self):
        return self._seg(self._qdot, self._ji["virtual_start"], 3)

    def get_mobile_joint_velocity(self):
        return self._seg(self._qdot, self._ji["mobi_start"], self._w)

    def get_manipulator_joint_velocity(self):
        return self._seg(self._qdot, self._ji["mani_start"], self._k)

    def get_joint_position_actuated(self):
        return self._sq(self._q_act)

    def get_joint_velocity_actuated(self):
        return self._sq(self._qdot_act)

    # ---- mobile base (Mobile::RobotData getters, mobile/robot_data.h:96-107)
    def _J_of(self, q_mobile):
        """(B, 3, w) base Jacobian at the wheel angles q_mobile."""
        qm = np.atleast_2d(np.asarray(q_mobile, np.float64))
        if self._caster:
            return self._base.fk(qm, None)[0]
        return np.broadcast_to(self._J_const, (qm.shape[0],) + self._J_const.shape)

    def get_mobile_FK_jacobian(self) -> np.ndarray:
        return self._sq(self._J_mobile).copy() if np.ndim(self._J_mobile) == 3 else np.array(self._J_mobile)

    get_FK_jacobian = get_mobile_FK_jacobian          # name used by earlier versions of this mirror

    def get_mobile_base_vel(self) -> np.ndarray:       # J_mobile * wheel_vel
        if np.ndim(self._J_mobile) == 3:
            return self._sq(np.einsum("brk,bk->br", self._J_mobile, self._wheel_vel))
        return self._sq(self._wheel_vel @ np.asarray(self._J_mobile).T)

    get_base_vel = get_mobile_base_vel

    def _selection(self, q_virtual, q_mobile):
        """(B, dof, act) selection matrix (robot_data.cpp:22-25, 115-120): identity blocks + Rz(yaw) J_mobile."""
        qv = np.atleast_2d(np.asarray(q_virtual, np.float64))
        J = self._J_of(q_mobile)
        B = max(qv.shape[0], J.shape[0])
        S = np.zeros((B, self._dof, self._act))
        S[:, self._ji["mani_start"] + np.arange(self._k), self._ai["mani_start"] + np.arange(self._k)] = 1.0
        S[:, self._ji["mobi_start"] + np.arange(self._w), self._ai["mobi_start"] + np.arange(self._w)] = 1.0
        c, s = np.cos(qv[:, 2]), np.sin(qv[:, 2])
        Rz = np.zeros((qv.shape[0], 3, 3))
        Rz[:, 0, 0], Rz[:, 0, 1], Rz[:, 1, 0], Rz[:, 1, 1], Rz[:, 2, 2] = c, -s, s, c, 1.0
        v0, m0 = self._ji["virtual_start"], self._ai["mobi_start"]
        S[:, v0:v0 + 3, m0:m0 + self._w] = Rz @ J
        return S

    def get_selection_matrix(self) -> np.ndarray:
        v0 = self._ji["virtual_start"]
        return self._sq(self._selection(self._q[:, v0:v0 + 3], self._q[:, self._ji["mobi_start"]:self._ji["mobi_start"] + self._w]))

    # ---- cached getters: full-dof quantities inherited from Manipulator::RobotData
    def _dyn(self, ctx, key, single=None):
        return self._sq(ctx.get_dynamics(want=(key,))[key], single)

    def get_mass_matrix(self):
        return self._dyn(self._ctx, "M")

    def get_mass_matrix_inv(self):
        return self._dyn(self._ctx, "Minv")

    def get_coriolis(self):
        return self._dyn(self._ctx, "c")

    def get_gravity(self):
        return self._dyn(self._ctx, "g")

    def get_nonlinear_effects(self):
        return self._dyn(self._ctx, "nle")

    def _fid(self, link_name):
        fid = self._model.frame_id(link_name)
        if fid < 0:   # reference: stderr + neutral value (manipulator/robot_data.cpp:380-384)
            print(f"\033[1;31mError: Link name {link_name} not found in URDF.\033[0m", file=sys.stderr)
        return fid

    def _frame(self, ctx, link_name, key, neutral, single=None):
        fid = self._fid(link_name)
        if fid < 0:
            return neutral
        v = ctx.get_frame(fid, want=(key,))[key]
        if key == "pose":
            v = engine.pose44(v)
        return self._sq(v, single)

    def get_pose(self, link_name: str):
        return self._frame(self._ctx, link_name, "pose", np.eye(4))

    def get_jacobian(self, link_name: str):
        return self._frame(self._ctx, link_name, "J", np.zeros((6, self._dof)))

    def get_jacobian_time_variation(self, link_name: str):
        return self._frame(self._ctx, link_name, "Jdot", np.zeros((6, self._dof)))

    def get_velocity(self, link_name: str):
        return self._frame(self._ctx, link_name, "vel", np.zeros(6))

    def _min_distance(self, ctx, with_grad, with_graddot, single=None) -> MinDistResult:
        d, g, gd, _ = ctx.get_min_distance(with_graddot=bool(with_graddot))
        z = np.zeros_like(g)
        return MinDistResult(self._sq(d, single), self._sq(g if (with_grad or with_graddot) else z, single),
                             self._sq(gd if with_graddot else z, single))

    def get_min_distance(self, with_grad: bool, with_graddot: bool, verbose: bool = False) -> MinDistResult:
        return self._min_distance(self._ctx, with_grad, with_graddot)

    # ---- cached getters: actuated space (robot_data.cpp:126-144, 407-415)
    def _get(self, ctx, link_name, key, neutral=None, single=None):
        fid = self._fid(link_name) if link_name is not None else 0
        if fid < 0:
            return neutral
        v = ctx.moma_get_state(fid, want=(key,))[key]
        if key == "pose":
            v = engine.pose44(v)
        return self._sq(v, single)

    def get_jacobian_actuated(self, link_name: str):
        return self._get(self._ctx, link_name, "J", np.zeros((6, self._act)))

    def get_jacobian_actuated_time_variation(self, link_name: str):
        return self._get(self._ctx, link_name, "Jdot", np.zeros((6, self._act)))

    def get_mass_matrix_actuated(self):
        return self._get(self._ctx, None, "M")

    def get_mass_matrix_actuated_inv(self):
        return self._get(self._ctx, None, "Minv")

    def get_gravity_actuated(self):
        return self._get(self._ctx, None, "g")

    def get_nonlinear_effects_actuated(self):
        return self._get(self._ctx, None, "nle")

    def get_coriolis_actuated(self):
        return self.get_nonlinear_effects_actuated() - self.get_gravity_actuated()

    def _manipulability(self, ctx, with_grad, with_graddot, link_name, single=None) -> ManipulabilityResult:
        fid = self._fid(link_name)
        if fid < 0:
            return ManipulabilityResult(0.0, np.zeros(self._k), np.zeros(self._k))
        r = ctx.moma_get_state(fid, want=("mani", "mani_grad", "mani_graddot"))
        z = np.zeros_like(r["mani_grad"])
        return ManipulabilityResult(self._sq(r["mani"], single), self._sq(r["mani_grad"] if (with_grad or with_graddot) else z, single),
                                    self._sq(r["mani_graddot"] if with_graddot else z, single))

    def get_manipulability(self, with_grad: bool, with_graddot: bool, link_name: str) -> ManipulabilityResult:
        return self._manipulability(self._ctx, with_grad, with_graddot, link_name)

    # ---- stateless twins (robot_data.cpp:146-405): evaluated on a private context, the cache stays untouched
    def _at(self, q_virtual, q_mobile, q_mani, qdot_virtual=None, qdot_mobile=None, qdot_mani=None):
        if self._scratch is None:
            self._scratch = engine.Context(self._model, self._max_batch, self._device)
        single = np.ndim(q_mani) == 1
        q = self.get_joint_vector(q_virtual, q_mobile, q_mani)
        z3, zw, zk = np.zeros((q.shape[0], 3)), np.zeros((q.shape[0], self._w)), np.zeros((q.shape[0], self._k))
        qd = self.get_joint_vector(z3 if qdot_virtual is None else qdot_virtual, zw if qdot_mobile is None else qdot_mobile,
                                   zk if qdot_mani is None else qdot_mani)
        self._scratch.moma_update_state(q, qd)
        return self._scratch, single

    def compute_mass_matrix(self, q_virtual, q_mobile, q_mani):
        c, s = self._at(q_virtual, q_mobile, q_mani)
        return self._dyn(c, "M", s)

    def compute_gravity(self, q_virtual, q_mobile, q_mani):
        c, s = self._at(q_virtual, q_mobile, q_mani)
        return self._dyn(c, "g", s)

    def compute_coriolis(self, q_virtual, q_mobile, q_mani, qdot_virtual, qdot_mobile, qdot_mani):
        c, s = self._at(q_virtual, q_mobile, q_mani, qdot_virtual, qdot_mobile, qdot_mani)
        return self._dyn(c, "c", s)

    def compute_nonlinear_effects(self, q_virtual, q_mobile, q_mani, qdot_virtual, qdot_mobile, qdot_mani):
        c, s = self._at(q_virtual, q_mobile, q_mani, qdot_virtual, qdot_mobile, qdot_mani)
        return self._dyn(c, "nle", s)

    def _act_twin(self, key, q_virtual, q_mobile, q_mani, qdot_mobile=None, qdot_mani=None, link_name=None):
        """S(q_virtual)' X(q with q_virtual = 0) (S), the reference's actuated twins (robot_data.cpp:185-232, 382-405)."""
        zv = np.zeros_like(np.atleast_2d(np.asarray(q_virtual, np.float64)))
        c, single = self._at(zv, q_mobile, q_mani, None, qdot_mobile, qdot_mani)
        S = self._selection(q_virtual, q_mobile)
        St = np.swapaxes(S, 1, 2)
        d = lambda k: c.get_dynamics(want=(k,))[k]
        if key == "M":
            out = St @ d("M") @ S
        elif key == "g":
            out = np.einsum("bij,bj->bi", St, d("g"))
        elif key == "c":
            out = np.einsum("bij,bj->bi", St, d("nle") - d("g"))
        elif key == "nle":
            out = np.einsum("bij,bj->bi", St, d("nle"))
        else:   # "J" / "Jdot"
            fid = self._fid(link_name)
            if fid < 0:
                return np.zeros((6, self._act))
            out = c.get_frame(fid, want=(key,))[key] @ S
        return self._sq(out, single)

    def compute_mass_matrix_actuated(self, q_virtual, q_mobile, q_mani):
        return self._act_twin("M", q_virtual, q_mobile, q_mani)

    def compute_gravity_actuated(self, q_virtual, q_mobile, q_mani):
        return self._act_twin("g", q_virtual, q_mobile, q_mani)

    def compute_coriolis_actuated(self, q_virtual, q_mobile, q_mani, qdot_mobile, qdot_mani):
        return self._act_twin("c", q_virtual, q_mobile, q_mani, qdot_mobile, qdot_mani)

    def compute_nonlinear_effects_actuated(self, q_virtual, q_mobile, q_mani, qdot_mobile, qdot_mani):
        return self._act_twin("nle", q_virtual, q_mobile, q_mani, qdot_mobile, qdot_mani)

    def compute_pose(self, q_virtual, q_mobile, q_mani, link_name: str):
        c, s = self._at(q_virtual, q_mobile, q_mani)
        return self._frame(c, link_name, "pose", np.eye(4), s)

    def compute_jacobian(self, q_virtual, q_mobile, q_mani, link_name: str):
        c, s = self._at(q_virtual, q_mobile, q_mani)
        return self._frame(c, link_name, "J", np.zeros((6, self._dof)), s)

    def compute_jacobian_time_variation(self, q_virtual, q_mobile, q_mani, qdot_virtual, qdot_mobile, qdot_mani, link_name: str):
        c, s = self._at(q_virtual, q_mobile, q_mani, qdot_virtual, qdot_mobile, qdot_mani)
        return self._frame(c, link_name, "Jdot", np.zeros((6, self._dof)), s)

    def compute_velocity(self, q_virtual, q_mobile, q_mani, qdot_virtual, qdot_mobile, qdot_mani, link_name: str):
        c, s = self._at(q_virtual, q_mobile, q_mani, qdot_virtual, qdot_mobile, qdot_mani)
        return self._frame(c, link_name, "vel", np.zeros(6), s)

    def compute_min_distance(self, q_virtual, q_mobile, q_mani, qdot_virtual, qdot_mobile, qdot_mani, with_grad: bool,
                             with_graddot: bool, verbose: bool = False) -> MinDistResult:
        c, s = self._at(q_virtual, q_mobile, q_mani, qdot_virtual, qdot_mobile, qdot_mani)
        return self._min_distance(c, with_grad, with_graddot, s)

    def compute_selection_matrix(self, q_virtual, q_mobile):
        return self._sq(self._selection(q_virtual, q_mobile), np.ndim(q_virtual) == 1)

    def compute_jacobian_actuated(self, q_virtual, q_mobile, q_mani, link_name: str):
        return self._act_twin("J", q_virtual, q_mobile, q_mani, link_name=link_name)

    def compute_jacobian_time_variation_actuated(self, q_virtual, q_mobile, q_mani, qdot_virtual, qdot_mobile, qdot_mani,
                                                 link_name: str):
        return self._act_twin("Jdot", q_virtual, q_mobile, q_mani, qdot_mobile, qdot_mani, link_name=link_name)

    def compute_manipulability(self, q_mani, qdot_mani, with_grad: bool, with_graddot: bool, link_name: str) -> ManipulabilityResult:
        """manipulator-only manipulability: base at the origin (robot_data.cpp:287-351; it does not depend on the base frame)."""
        qa = np.atleast_2d(np.asarray(q_mani, np.float64))
        z3, zw = np.zeros((qa.shape[0], 3)), np.zeros((qa.shape[0], self._w))
        c, s = self._at(z3, zw, q_mani, z3, zw, qdot_mani)
        return self._manipulability(c, with_grad, with_graddot, link_name, s)

    def compute_mobile_FK_jacobian(self, q_mobile):
        return self._sq(np.array(self._J_of(q_mobile)), np.ndim(q_mobile) == 1)

    def compute_mobile_base_vel(self, q_mobile, qdot_mobile):
        J = self._J_of(q_mobile)
        return self._sq(np.einsum("brk,bk->br", J, np.atleast_2d(np.asarray(qdot_mobile, np.float64))), np.ndim(q_mobile) == 1)
