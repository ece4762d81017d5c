"""Manipulator RobotData with the reference's interface (reference drc/manipulator/robot_data.py:6-355 wrapping
src/manipulator/robot_data.cpp) on top of the batched engine.  Cached getters read the device state cache written by
update_state; the stateless compute_* twins (robot_data.cpp:128-374) run on a second, private context so that they do
not disturb the cache -- exactly the reference's semantics."""
from __future__ import annotations

import sys
from typing import Tuple

import numpy as np

from ... import engine
from ..type_define import ManipulabilityResult, MinDistResult


class RobotData:
    def __init__(self, urdf_path: str, srdf_path: str = "", packages_path: str = "", max_batch: int = 1, device: int = 0):
        # the reference exits the process on a missing URDF (robot_data.cpp:15-19); here it is an exception
        self._model = engine.Model(urdf_path, srdf_path, packages_path)
        self._ctx = engine.Context(self._model, max_batch, device)
        self._scratch = None
        self._max_batch, self._device = max_batch, device
        self._single = True
        self._q = np.zeros(self._model.dof)
        self._qdot = np.zeros(self._model.dof)

    # ---- plumbing
    def _stateless(self) -> engine.Context:
        if self._scratch is None:
            self._scratch = engine.Context(self._model, self._max_batch, self._device)
        return self._scratch

    def _squeeze(self, a, single=None):
        return a[0] if (self._single if single is None else single) else a

    def _link(self, ctx, link_name: str):
        fid = self._model.frame_id(link_name)
        if fid < 0:  # reference: stderr + neutral value (robot_data.cpp:380-384)
            print(f"\033[1;31mError: Link name {link_name} not found in URDF.\033[0m", file=sys.stderr)
        return fid

    def get_verbose(self) -> str:
        return self._model.verbose()

    # ---- state update (robot_data.cpp:91-124)
    def update_state(self, q: np.ndarray, qdot: np.ndarray) -> bool:
        q, qdot = np.asarray(q, np.float64), np.asarray(qdot, np.float64)
        self._single = q.ndim == 1
        self._q, self._qdot = q.copy(), qdot.copy()
        return bool(self._ctx.update_state(q, qdot))

    # ---- cached getters
    def get_dof(self) -> int:
        return self._model.dof

    def get_joint_position(self) -> np.ndarray:
        return self._q

    def get_joint_velocity(self) -> np.ndarray:
        return self._qdot

    def get_joint_position_limit(self) -> Tuple[np.ndarray, np.ndarray]:
        return self._model.q_lower.copy(), self._model.q_upper.copy()

    def get_joint_velocity_limit(self) -> Tuple[np.ndarray, np.ndarray]:
        return -self._model.v_limit, self._model.v_limit.copy()

    def _dyn(self, ctx, key, single=None):
        if isinstance(ctx, tuple):      # (scratch context, single) from _at(): stateless twins
            ctx, single = ctx
        return self._squeeze(ctx.get_dynamics(want=(key,))[key], single)

    def get_mass_matrix(self) -> np.ndarray:
        return self._dyn(self._ctx, "M")

    def get_mass_matrix_inv(self) -> np.ndarray:
        return self._dyn(self._ctx, "Minv")

    def get_coriolis(self) -> np.ndarray:
        return self._dyn(self._ctx, "c")

    def get_gravity(self) -> np.ndarray:
        return self._dyn(self._ctx, "g")

    def get_nonlinear_effects(self) -> np.ndarray:
        return self._dyn(self._ctx, "nle")

    def _frame(self, ctx, link_name, key, neutral, single=None):
        if isinstance(ctx, tuple):
            ctx, single = ctx
        fid = self._link(ctx, link_name)
        if fid < 0:
            return neutral
        v = ctx.get_frame(fid, want=(key,))[key]
        if key == "pose":
            v = engine.pose44(v)
        return self._squeeze(v, single)

    def get_pose(self, link_name: str) -> np.ndarray:
        return self._frame(self._ctx, link_name, "pose", np.eye(4))

    def get_jacobian(self, link_name: str) -> np.ndarray:
        return self._frame(self._ctx, link_name, "J", np.zeros((6, self._model.dof)))

    def get_jacobian_time_variation(self, link_name: str) -> np.ndarray:
        return self._frame(self._ctx, link_name, "Jdot", np.zeros((6, self._model.dof)))

    def get_velocity(self, link_name: str) -> np.ndarray:
        return self._frame(self._ctx, link_name, "vel", np.zeros(6))

    def _min_distance(self, ctx, with_grad, with_graddot, single=None) -> MinDistResult:
        if isinstance(ctx, tuple):
            ctx, single = ctx
        d, g, gd, _ = ctx.get_min_distance(with_graddot=bool(with_graddot))
        z = np.zeros_like(g)
        r = MinDistResult(self._squeeze(d, single), self._squeeze(g if (with_grad or with_graddot) else z, single),
                          self._squeeze(gd if with_graddot else z, single))
        return r

    def get_min_distance(self, with_grad: bool, with_graddot: bool, verbose: bool = False) -> MinDistResult:
        return self._min_distance(self._ctx, with_grad, with_graddot)

    def _manipulability(self, ctx, with_grad, with_graddot, link_name, single=None) -> ManipulabilityResult:
        if isinstance(ctx, tuple):
            ctx, single = ctx
        fid = self._link(ctx, link_name)
        n = self._model.dof
        if fid < 0:
            return ManipulabilityResult(0.0, np.zeros(n), np.zeros(n))
        m, g, gd = ctx.get_manipulability(fid, with_graddot=bool(with_graddot))
        z = np.zeros_like(g)
        return ManipulabilityResult(self._squeeze(m, single), self._squeeze(g if (with_grad or with_graddot) else z, single),
                                    self._squeeze(gd if with_graddot else z, single))

    def get_manipulability(self, with_grad: bool, with_graddot: bool, link_name: str) -> ManipulabilityResult:
        return self._manipulability(self._ctx, with_grad, with_graddot, link_name)

    # ---- stateless twins (robot_data.cpp:128-374): evaluate at (q, qdot) without touching the cache
    def _at(self, q, qdot=None):
        """(scratch context holding the state (q, qdot), single-robot flag of THIS call): the cached getters' flag, set by
        update_state only, is left alone."""
        q = np.asarray(q, np.float64)
        s = self._stateless()
        s.update_state(q, np.zeros_like(q) if qdot is None else qdot)
        return s, q.ndim == 1

    def compute_mass_matrix(self, q) -> np.ndarray:
        return self._dyn(self._at(q), "M")

    def compute_gravity(self, q) -> np.ndarray:
        return self._dyn(self._at(q), "g")

    def compute_coriolis(self, q, qdot) -> np.ndarray:
        return self._dyn(self._at(q, qdot), "c")

    def compute_nonlinear_effects(self, q, qdot) -> np.ndarray:
        return self._dyn(self._at(q, qdot), "nle")

    def compute_pose(self, q, link_name: str) -> np.ndarray:
        return self._frame(self._at(q), link_name, "pose", np.eye(4))

    def compute_jacobian(self, q, link_name: str) -> np.ndarray:
        return self._frame(self._at(q), link_name, "J", np.zeros((6, self._model.dof)))

    def compute_jacobian_time_variation(self, q, qdot, link_name: str) -> np.ndarray:
        return self._frame(self._at(q, qdot), link_name, "Jdot", np.zeros((6, self._model.dof)))

    def compute_velocity(self, q, qdot, link_name: str) -> np.ndarray:
        return self._frame(self._at(q, qdot), link_name, "vel", np.zeros(6))

    def compute_min_distance(self, q, qdot, with_grad: bool, with_graddot: bool, verbose: bool = False) -> MinDistResult:
        return self._min_distance(self._at(q, qdot), with_grad, with_graddot)

    def compute_manipulability(self, q, qdot, with_grad: bool, with_graddot: bool, link_name: str) -> ManipulabilityResult:
        return self._manipulability(self._at(q, qdot), with_grad, with_graddot, link_name)
