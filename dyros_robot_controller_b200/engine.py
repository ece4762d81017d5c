"""Batched control-cycle engine: thin Python front-end over the C ABI (include/drc_b200.h).

* numpy arrays  -> drc_host_* entry points (host buffers, H2D/D2H inside the call, synchronous)
* torch CUDA tensors -> drc_batch_* entry points (device pointers, asynchronous on the current
  torch stream); torch is only plumbing here (device memory, streams).

Shapes are batch-major: q (B, n), poses (B, 4, 4) / (B, 12) [top three rows of the homogeneous
matrix], task vectors (B, 6), matrices (B, n, n) / (B, 6, n).  A leading batch axis may be omitted
for a single robot.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _capi
from ._capi import DrcParams, check, lib

_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int)


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def pose12(T) -> np.ndarray:
    """(...,4,4) / (...,3,4) / (...,12) -> (...,12)."""
    T = np.asarray(T, np.float64)
    if T.shape[-1] == 12 and (T.ndim == 1 or T.shape[-2:] != (3, 4)):
        return np.ascontiguousarray(T)
    return np.ascontiguousarray(T[..., :3, :].reshape(T.shape[:-2] + (12,)))


def pose44(p) -> np.ndarray:
    p = np.asarray(p, np.float64)
    out = np.zeros(p.shape[:-1] + (4, 4))
    out[..., :3, :] = p.reshape(p.shape[:-1] + (3, 4))
    out[..., 3, 3] = 1.0
    return out


class Model:
    """Compiled robot model (replaces the Pinocchio model the reference builds in RobotData's ctor)."""

    def __init__(self, urdf_path: str, srdf_path: str = "", packages_path: str = ""):
        h = C.c_void_p()
        check(lib().drc_model_create_from_urdf(str(urdf_path).encode(), str(srdf_path).encode(),
                                               str(packages_path).encode(), C.byref(h)), "drc_model_create_from_urdf")
        self._h = h
        self.dof = lib().drc_model_dof(h)
        s = (C.c_int * 6)()
        check(lib().drc_model_info(h, s), "drc_model_info")
        self.info = dict(dof=s[0], geoms=s[1], pairs=s[2], groups=s[3], frames=s[4], skipped_geoms=s[5], skipped_meshes=0)
        s2 = (C.c_int * 2)()
        check(lib().drc_model_mesh_info(h, s2), "drc_model_mesh_info")
        self.info.update(mesh_geoms=s2[0], hull_vertices=s2[1])
        n = self.dof
        lo, hi, vl, ef = (np.zeros(n) for _ in range(4))
        check(lib().drc_model_limits(h, lo.ctypes.data_as(_D), hi.ctypes.data_as(_D), vl.ctypes.data_as(_D),
                                     ef.ctypes.data_as(_D)), "drc_model_limits")
        self.q_lower, self.q_upper, self.v_limit, self.effort_limit = lo, hi, vl, ef
        self.frame_names = [lib().drc_model_frame_name(h, i).decode() for i in range(self.info["frames"])]
        self.joint_names = [lib().drc_model_joint_name(h, i).decode() for i in range(n)]

    def frame_id(self, link_name: str) -> int:
        return lib().drc_model_frame_id(self._h, link_name.encode())

    # ---- mobile base (reference Mobile::KinematicParam, JointIndex, ActuatorIndex: type_define.h:13-72)
    DRIVE_TYPES = dict(Differential=0, Mecanum=1, Caster=2)

    def attach_mobile_base(self, kin: dict, joint_idx: dict, actuator_idx: dict):
        """kin: type ("Differential" | "Mecanum" | "Caster" or 0/1/2), wheel_radius, base_width, wheel_offset,
        roller_angles, base2wheel_positions [(x, y)...], base2wheel_angles.  Call before creating contexts."""
        t = kin["type"] if isinstance(kin["type"], (int, np.integer)) else self.DRIVE_TYPES[kin["type"]]
        pos = np.asarray(kin.get("base2wheel_positions", np.zeros((0, 2))), np.float64).reshape(-1, 2)
        w = 2 if t == 0 else (len(kin["roller_angles"]) if t == 1 else 2 * len(pos))
        arr = lambda a: np.ascontiguousarray(np.asarray(a, np.float64))
        ra = arr(kin.get("roller_angles", np.zeros(w)))
        ba = arr(kin.get("base2wheel_angles", np.zeros(w)))
        bx, by = arr(pos[:, 0]) if len(pos) else np.zeros(w), arr(pos[:, 1]) if len(pos) else np.zeros(w)
        check(lib().drc_model_attach_mobile_base(
            self._h, int(t), C.c_double(kin.get("wheel_radius", 0.0)), C.c_double(kin.get("base_width", 0.0)),
            C.c_double(kin.get("wheel_offset", 0.0)), int(w), ra.ctypes.data_as(_D), bx.ctypes.data_as(_D), by.ctypes.data_as(_D),
            ba.ctypes.data_as(_D), int(joint_idx["virtual_start"]), int(joint_idx["mani_start"]), int(joint_idx["mobi_start"]),
            int(actuator_idx["mani_start"]), int(actuator_idx["mobi_start"])), "drc_model_attach_mobile_base")
        s = (C.c_int * 4)()
        check(lib().drc_model_moma_info(self._h, s), "drc_model_moma_info")
        self.drive_type, self.wheel_num, self.mani_dof, self.actuated_dof = s[0], s[1], s[2], s[3]
        self.joint_idx, self.actuator_idx = dict(joint_idx), dict(actuator_idx)
        return self

    def base_jacobian(self) -> np.ndarray:
        J = np.zeros((3, self.wheel_num))
        check(lib().drc_model_base_jacobian(self._h, J.ctypes.data_as(_D)), "drc_model_base_jacobian")
        return J

    def verbose(self) -> str:
        return lib().drc_model_verbose(self._h).decode()

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().drc_model_destroy(self._h)
                self._h = None
        except Exception:
            pass


class MobileBase:
    """Batched mobile base (reference Mobile::RobotData + Mobile::RobotController; no URDF, only the KinematicParam).

    kin: type ("Differential" | "Mecanum" | "Caster" or 0/1/2), wheel_radius, base_width, wheel_offset, max_lin_speed,
    max_ang_speed, max_lin_acc, max_ang_acc, roller_angles, base2wheel_positions [(x, y)...], base2wheel_angles
    (type_define.h:58-72).  numpy arrays go through the drc_host_mobile_* entry points, float64 CUDA tensors through
    drc_batch_mobile_* on the current torch stream."""

    def __init__(self, kin: dict, device: int = 0):
        t = kin["type"] if isinstance(kin["type"], (int, np.integer)) else Model.DRIVE_TYPES[kin["type"]]
        pos = np.asarray(kin.get("base2wheel_positions", np.zeros((0, 2))), np.float64).reshape(-1, 2)
        w = 2 if t == 0 else (len(kin.get("roller_angles", [])) if t == 1 else 2 * len(pos))
        arr = lambda a: np.ascontiguousarray(np.asarray(a, np.float64))
        ra = arr(kin.get("roller_angles", np.zeros(w)))
        ba = arr(kin.get("base2wheel_angles", np.zeros(w)))
        bx, by = (arr(pos[:, 0]), arr(pos[:, 1])) if len(pos) else (np.zeros(max(w, 1)), np.zeros(max(w, 1)))
        h = C.c_void_p()
        g = lambda k: C.c_double(float(kin.get(k, 0.0)))
        check(lib().drc_mobile_create(int(t), g("wheel_radius"), g("base_width"), g("wheel_offset"), g("max_lin_speed"),
                                      g("max_ang_speed"), g("max_lin_acc"), g("max_ang_acc"), int(w), ra.ctypes.data_as(_D),
                                      bx.ctypes.data_as(_D), by.ctypes.data_as(_D), ba.ctypes.data_as(_D), int(device),
                                      C.byref(h)), "drc_mobile_create")
        self._h = h
        self.drive_type, self.wheel_num, self.kin = int(t), int(w), dict(kin)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().drc_mobile_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def launch_count(self) -> int:
        return int(lib().drc_mobile_launch_count(self._h))

    def _run(self, fk, wheel_pos, vec, want_J, saturate=False):
        w = self.wheel_num
        kin_in, kout = (w, 3) if fk else (3, w)
        name = "mobile_fk" if fk else "mobile_ik"
        if _is_torch(wheel_pos) or _is_torch(vec):
            import torch
            ref = wheel_pos if _is_torch(wheel_pos) else vec
            wp, B = Context._t_in(None, wheel_pos, w)
            v, B = Context._t_in(None, vec, kin_in, B)
            J = torch.empty((B, 3, w) if fk else (B, w, 3), dtype=torch.float64, device=ref.device) if want_J else None
            out = torch.empty((B, kout), dtype=torch.float64, device=ref.device) if v is not None else None
            tp, st = Context._tp, Context._stream()
            if fk:
                rc = lib().drc_batch_mobile_fk(self._h, B, tp(wp), tp(v), tp(J), tp(out), _capi.LAYOUT_AOS, st)
            else:
                rc = lib().drc_batch_mobile_ik(self._h, B, tp(wp), tp(v), int(bool(saturate)), tp(J), tp(out), _capi.LAYOUT_AOS, st)
            check(rc, "drc_batch_" + name)
            return J, out
        wp, B = Context._np_in(wheel_pos, w)
        v, B = Context._np_in(vec, kin_in, B)
        J = np.zeros((B, 3, w) if fk else (B, w, 3)) if want_J else None
        out = np.zeros((B, kout)) if v is not None else None
        p = Context._p
        if fk:
            rc = lib().drc_host_mobile_fk(self._h, B, p(wp), p(v), p(J), p(out))
        else:
            rc = lib().drc_host_mobile_ik(self._h, B, p(wp), p(v), int(bool(saturate)), p(J), p(out))
        check(rc, "drc_host_" + name)
        return J, out

    def fk(self, wheel_pos, wheel_vel=None, want_J=True):
        """Mobile::RobotData::updateState: (J_fk (B,3,w), base_vel (B,3)).  wheel_pos may be None for differential /
        mecanum bases when wheel_vel is given."""
        return self._run(True, wheel_pos, wheel_vel, want_J)

    def ik(self, wheel_pos, base_vel=None, saturate=False, want_J=True):
        """Mobile::RobotController: (J_ik (B,w,3), wheel_vel (B,w)); saturate=True is VelocityCommand."""
        return self._run(False, wheel_pos, base_vel, want_J, saturate)


class Context:
    """Device state cache + scratch for up to `max_batch` robots on one GPU."""

    def __init__(self, model: Model, max_batch: int, device: int = 0):
        h = C.c_void_p()
        check(lib().drc_ctx_create(model._h, int(device), int(max_batch), C.byref(h)), "drc_ctx_create")
        self._h = h
        self.model = model
        self.n = model.dof
        self.device = device
        self.max_batch = max_batch

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().drc_ctx_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ parameters
    def get_params(self) -> DrcParams:
        p = DrcParams()
        check(lib().drc_ctx_get_params(self._h, C.byref(p)), "drc_ctx_get_params")
        return p

    def set_params(self, **kw):
        p = self.get_params()
        for k, v in kw.items():
            cur = getattr(p, k)
            if hasattr(cur, "__len__"):
                v = np.asarray(v, np.float64).ravel()
                for i in range(len(v)):
                    cur[i] = float(v[i])
            else:
                setattr(p, k, v)
        check(lib().drc_ctx_set_params(self._h, C.byref(p)), "drc_ctx_set_params")

    def synchronize(self):
        check(lib().drc_ctx_synchronize(self._h), "drc_ctx_synchronize")

    def enable_timing(self, on=True):
        check(lib().drc_ctx_enable_timing(self._h, int(on)), "drc_ctx_enable_timing")

    def last_timing(self):
        ms = (C.c_float * 4)()
        check(lib().drc_ctx_last_timing(self._h, ms), "drc_ctx_last_timing")
        return dict(collision_ms=ms[0], build_ms=ms[1], admm_ms=ms[2], total_ms=ms[3])

    def last_trace(self):
        """[(name, ms since the start of the call)] of the last fused cycle call (main and priority pipeline marks)."""
        ms = (C.c_float * 32)()
        names = C.create_string_buffer(1024)
        n = lib().drc_ctx_last_trace(self._h, 32, ms, names, 1024)
        if n < 0:
            check(n, "drc_ctx_last_trace")
        return list(zip(names.value.decode().split(";")[:n], [float(ms[i]) for i in range(n)]))

    # QP shapes (NC core variables, KU unit bundles per variable, NR dense + equality rows) of the four formulations
    def _qp_shape(self, kind: str):
        moma = kind.startswith("moma")
        nc = self.model.actuated_dof if moma else self.n
        ident = kind.endswith("id")
        return nc, (4 if ident else 2), (2 + nc if ident else 2)

    def enable_qp_debug(self, on=True):
        """keep the primal / dual vectors of every QP solve on the device (test instrumentation, see qp_debug)."""
        check(lib().drc_ctx_enable_qp_debug(self._h, int(on)), "drc_ctx_enable_qp_debug")

    def qp_debug(self, kind: str, B: Optional[int] = None):
        """primal x and unscaled dual y of the last QP call of `kind` ("ik", "id", "moma_ik", "moma_id") in STRUCTURED order:
        x = [core (NC) | unit slacks (KU, NC) | row singletons (NR)],
        y = [core bound rows (NC) | unit rows (KU, NC) | unit-slack bound rows (KU, NC) | rows (NR) | row-singleton bound rows (NR)]."""
        B = self._B if B is None else B
        nc, ku, nr = self._qp_shape(kind)
        nx, ny = nc * (1 + ku) + nr, nc * (1 + 2 * ku) + 2 * nr
        x, y = np.zeros((B, nx)), np.zeros((B, ny))
        check(lib().drc_host_get_qp_debug(self._h, B, nx, ny, self._p(x), self._p(y)), "drc_host_get_qp_debug")
        return dict(x=x, y=y, nc=nc, ku=ku, nr=nr)

    @property
    def launch_count(self) -> int:
        return int(lib().drc_ctx_launch_count(self._h))

    # ------------------------------------------------------------------ argument plumbing
    def _frame(self, link) -> int:
        fid = link if isinstance(link, (int, np.integer)) else self.model.frame_id(link)
        if fid < 0:
            raise KeyError(f"link '{link}' not found in the URDF")
        return int(fid)

    @staticmethod
    def _np_in(a, K, B=None):
        if a is None:
            return None, B
        a = np.ascontiguousarray(np.asarray(a, np.float64).reshape(-1, K))
        if B is not None and a.shape[0] != B:
            raise ValueError(f"expected batch {B}, got {a.shape[0]}")
        return a, a.shape[0]

    @staticmethod
    def _p(a):
        return None if a is None else a.ctypes.data_as(_D)

    @staticmethod
    def _pi(a):
        return None if a is None else a.ctypes.data_as(_I)

    # torch helpers (device path)
    def _t_in(self, t, K, B=None):
        import torch
        if t is None:
            return None, B
        if t.dtype != torch.float64 or not t.is_cuda:
            raise TypeError("device entry points need float64 CUDA tensors")
        t = t.reshape(-1, K).contiguous()
        if B is not None and t.shape[0] != B:
            raise ValueError(f"expected batch {B}, got {t.shape[0]}")
        return t, t.shape[0]

    @staticmethod
    def _tp(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    @staticmethod
    def _stream():
        import torch
        h = torch.cuda.current_stream().cuda_stream
        return C.c_void_p(h if h else 1)  # handle 0 is torch's default stream = cudaStreamLegacy (0x1); NULL means "ctx stream"

    # ------------------------------------------------------------------ RobotData
    def update_state(self, q, qdot):
        if _is_torch(q):
            q, B = self._t_in(q, self.n)
            qd, _ = self._t_in(qdot, self.n, B)
            check(lib().drc_batch_update_state(self._h, B, self._tp(q), self._tp(qd), _capi.LAYOUT_AOS, self._stream()),
                  "drc_batch_update_state")
            self._B = B
            return True
        q, B = self._np_in(q, self.n)
        qd, _ = self._np_in(qdot, self.n, B)
        check(lib().drc_host_update_state(self._h, B, self._p(q), self._p(qd)), "drc_host_update_state")
        self._B = B
        return True

    def get_frame(self, link, want=("pose", "J", "Jdot", "vel")):
        B, n, f = self._B, self.n, self._frame(link)
        o = dict(pose=np.zeros((B, 12)) if "pose" in want else None, J=np.zeros((B, 6, n)) if "J" in want else None,
                 Jdot=np.zeros((B, 6, n)) if "Jdot" in want else None, vel=np.zeros((B, 6)) if "vel" in want else None)
        check(lib().drc_host_get_frame(self._h, B, f, self._p(o["pose"]), self._p(o["J"]), self._p(o["Jdot"]),
                                       self._p(o["vel"])), "drc_host_get_frame")
        return {k: v for k, v in o.items() if v is not None}

    def get_dynamics(self, want=("M", "Minv", "g", "c", "nle")):
        B, n = self._B, self.n
        o = dict(M=np.zeros((B, n, n)) if "M" in want else None, Minv=np.zeros((B, n, n)) if "Minv" in want else None,
                 g=np.zeros((B, n)) if "g" in want else None, c=np.zeros((B, n)) if "c" in want else None,
                 nle=np.zeros((B, n)) if "nle" in want else None)
        check(lib().drc_host_get_dynamics(self._h, B, self._p(o["M"]), self._p(o["Minv"]), self._p(o["g"]), self._p(o["c"]),
                                          self._p(o["nle"])), "drc_host_get_dynamics")
        return {k: v for k, v in o.items() if v is not None}

    def get_manipulability(self, link, with_graddot=False):
        B, n, f = self._B, self.n, self._frame(link)
        m, g, gd = np.zeros(B), np.zeros((B, n)), np.zeros((B, n))
        check(lib().drc_host_get_manipulability(self._h, B, f, int(with_graddot), self._p(m), self._p(g), self._p(gd)),
              "drc_host_get_manipulability")
        return m, g, gd

    def get_min_distance(self, with_graddot=False):
        B, n = self._B, self.n
        d, g, gd, pr = np.zeros(B), np.zeros((B, n)), np.zeros((B, n)), np.zeros(B, np.int32)
        check(lib().drc_host_get_min_distance(self._h, B, int(with_graddot), self._p(d), self._p(g), self._p(gd), self._pi(pr)),
              "drc_host_get_min_distance")
        return d, g, gd, pr

    # ------------------------------------------------------------------ RobotController (cached state)
    def _qp_host(self, fn, name, B, args, n_out2=False):
        out, st, it = np.zeros((B, self.n)), np.zeros(B, np.int32), np.zeros(B, np.int32)
        if n_out2:
            out2 = np.zeros((B, self.n))
            check(fn(self._h, B, *args, self._p(out), self._p(out2), self._pi(st), self._pi(it)), name)
            return dict(out=out, qddot=out2, status=st, iters=it)
        check(fn(self._h, B, *args, self._p(out), self._pi(st), self._pi(it)), name)
        return dict(out=out, status=st, iters=it)

    def qpik(self, xdot_des, link):
        x, B = self._np_in(xdot_des, 6, self._B)
        return self._qp_host(lib().drc_host_qpik, "drc_host_qpik", B, (self._p(x), self._frame(link)))

    def qpik_step(self, x_target, xdot_target, link):
        xt, B = self._np_in(pose12(x_target), 12, self._B)
        xd, _ = self._np_in(xdot_target, 6, B)
        return self._qp_host(lib().drc_host_qpik_step, "drc_host_qpik_step", B, (self._p(xt), self._p(xd), self._frame(link)))

    def qpid(self, xddot_des, link):
        x, B = self._np_in(xddot_des, 6, self._B)
        return self._qp_host(lib().drc_host_qpid, "drc_host_qpid", B, (self._p(x), self._frame(link)), n_out2=True)

    def qpid_step(self, x_target, xdot_target, link):
        xt, B = self._np_in(pose12(x_target), 12, self._B)
        xd, _ = self._np_in(xdot_target, 6, B)
        return self._qp_host(lib().drc_host_qpid_step, "drc_host_qpid_step", B, (self._p(xt), self._p(xd), self._frame(link)),
                             n_out2=True)

    def _taskspace_torch(self, fn_name, x_target, xdot_target, null_vec, link, out):
        """device path of CLIKStep / OSFStep: torch CUDA tensors in, torch tensor out (asynchronous)."""
        import torch
        xt, B = self._t_in(x_target, 12)
        xd, _ = self._t_in(xdot_target, 6, B)
        nv = None
        if null_vec is not None:
            nv, _ = self._t_in(null_vec, self.n, B)
        out = torch.empty((B, self.n), dtype=torch.float64, device=xt.device) if out is None else out
        check(getattr(lib(), fn_name)(self._h, B, self._tp(xt), self._tp(xd), self._tp(nv), self._frame(link), self._tp(out),
                                      _capi.LAYOUT_AOS, self._stream()), fn_name)
        return out

    def clik_step(self, x_target, xdot_target, link, null_qdot=None, out=None):
        if _is_torch(x_target):
            return self._taskspace_torch("drc_batch_clik_step", x_target, xdot_target, null_qdot, link, out)
        xt, B = self._np_in(pose12(x_target), 12, self._B)
        xd, _ = self._np_in(xdot_target, 6, B)
        nq, _ = self._np_in(null_qdot, self.n, B)
        out = np.zeros((B, self.n)) if out is None else out
        check(lib().drc_host_clik_step(self._h, B, self._p(xt), self._p(xd), self._p(nq), self._frame(link), self._p(out)),
              "drc_host_clik_step")
        return out

    def osf(self, xddot_target, link, null_torque=None):
        x, B = self._np_in(xddot_target, 6, self._B)
        nt, _ = self._np_in(null_torque, self.n, B)
        out = np.zeros((B, self.n))
        check(lib().drc_host_osf(self._h, B, self._p(x), self._p(nt), self._frame(link), self._p(out)), "drc_host_osf")
        return out

    def osf_step(self, x_target, xdot_target, link, null_torque=None, out=None):
        if _is_torch(x_target):
            return self._taskspace_torch("drc_batch_osf_step", x_target, xdot_target, null_torque, link, out)
        xt, B = self._np_in(pose12(x_target), 12, self._B)
        xd, _ = self._np_in(xdot_target, 6, B)
        nt, _ = self._np_in(null_torque, self.n, B)
        out = np.zeros((B, self.n)) if out is None else out
        check(lib().drc_host_osf_step(self._h, B, self._p(xt), self._p(xd), self._p(nt), self._frame(link), self._p(out)),
              "drc_host_osf_step")
        return out

    def joint_torque_step(self, q_target, qdot_target):
        qt, B = self._np_in(q_target, self.n, self._B)
        qdt, _ = self._np_in(qdot_target, self.n, B)
        out = np.zeros((B, self.n))
        check(lib().drc_host_joint_torque_step(self._h, B, self._p(qt), self._p(qdt), self._p(out)),
              "drc_host_joint_torque_step")
        return out

    def task_space_cubic(self, x_target, xdot_target, x_init, xdot_init, t, t0, duration):
        xt, B = self._np_in(pose12(x_target), 12)
        xd, _ = self._np_in(xdot_target, 6, B)
        xi, _ = self._np_in(pose12(x_init), 12, B)
        xdi, _ = self._np_in(xdot_init, 6, B)
        o1, o2 = np.zeros((B, 12)), np.zeros((B, 6))
        check(lib().drc_host_task_space_cubic(self._h, B, self._p(xt), self._p(xd), self._p(xi), self._p(xdi), C.c_double(t),
                                              C.c_double(t0), C.c_double(duration), self._p(o1), self._p(o2)),
              "drc_host_task_space_cubic")
        return o1, o2

    # ------------------------------------------------------------------ fused control cycle
    def cycle_qpik_step(self, q, qdot, x_target, xdot_target, link, out=None, status=None, iters=None):
        """updateState + QPIKStep.  numpy in -> numpy out (host path); torch CUDA in -> torch out (async)."""
        f = self._frame(link)
        if _is_torch(q):
            import torch
            q, B = self._t_in(q, self.n)
            qd, _ = self._t_in(qdot, self.n, B)
            xt, _ = self._t_in(x_target, 12, B)
            xd, _ = self._t_in(xdot_target, 6, B)
            out = torch.empty((B, self.n), dtype=torch.float64, device=q.device) if out is None else out
            status = torch.empty(B, dtype=torch.int32, device=q.device) if status is None else status
            iters = torch.empty(B, dtype=torch.int32, device=q.device) if iters is None else iters
            check(lib().drc_batch_cycle_qpik_step(self._h, B, self._tp(q), self._tp(qd), self._tp(xt), self._tp(xd), f,
                                                  self._tp(out), self._tp(status), self._tp(iters), _capi.LAYOUT_AOS,
                                                  self._stream()), "drc_batch_cycle_qpik_step")
            self._B = B
            return dict(out=out, status=status, iters=iters)
        q, B = self._np_in(q, self.n)
        qd, _ = self._np_in(qdot, self.n, B)
        xt, _ = self._np_in(pose12(x_target), 12, B)
        xd, _ = self._np_in(xdot_target, 6, B)
        out = np.zeros((B, self.n)) if out is None else out
        status = np.zeros(B, np.int32) if status is None else status
        iters = np.zeros(B, np.int32) if iters is None else iters
        check(lib().drc_host_cycle_qpik_step(self._h, B, self._p(q), self._p(qd), self._p(xt), self._p(xd), f, self._p(out),
                                             self._pi(status), self._pi(iters)), "drc_host_cycle_qpik_step")
        self._B = B
        return dict(out=out, status=status, iters=iters)

    def cycle_clik_osf_step(self, q, qdot, x_target, xdot_target, link, out=None, out2=None):
        """updateState + CLIKStep (qdot*, `out`) + OSFStep (tau*, `out2`) as ONE launch (BASELINE config 2), no null-space vectors.
        numpy in -> numpy out (host path); torch CUDA in -> torch out (async).  Leaves the full state cache like update_state."""
        f = self._frame(link)
        if _is_torch(q):
            import torch
            q, B = self._t_in(q, self.n)
            qd, _ = self._t_in(qdot, self.n, B)
            xt, _ = self._t_in(x_target, 12, B)
            xd, _ = self._t_in(xdot_target, 6, B)
            out = torch.empty((B, self.n), dtype=torch.float64, device=q.device) if out is None else out
            out2 = torch.empty((B, self.n), dtype=torch.float64, device=q.device) if out2 is None else out2
            check(lib().drc_batch_cycle_clik_osf_step(self._h, B, self._tp(q), self._tp(qd), self._tp(xt), self._tp(xd), f,
                                                      self._tp(out), self._tp(out2), _capi.LAYOUT_AOS, self._stream()),
                  "drc_batch_cycle_clik_osf_step")
            self._B = B
            return dict(qdot=out, tau=out2)
        q, B = self._np_in(q, self.n)
        qd, _ = self._np_in(qdot, self.n, B)
        xt, _ = self._np_in(pose12(x_target), 12, B)
        xd, _ = self._np_in(xdot_target, 6, B)
        out = np.zeros((B, self.n)) if out is None else out
        out2 = np.zeros((B, self.n)) if out2 is None else out2
        check(lib().drc_host_cycle_clik_osf_step(self._h, B, self._p(q), self._p(qd), self._p(xt), self._p(xd), f, self._p(out),
                                                 self._p(out2)), "drc_host_cycle_clik_osf_step")
        self._B = B
        return dict(qdot=out, tau=out2)

    def rollout_qpik(self, q, qdot, x_target, xdot_target, link, ticks, dt, x_init=None, xdot_init=None, t_start=0.0, t0=0.0,
                     duration=0.0):
        """Closed-loop rollout: `ticks` control cycles of updateState + QPIKCubic (duration > 0) / QPIKStep, each followed by
        the example's integrate step q += qdot* dt, qdot = qdot* (examples/C++/src/fr3_controller.cpp:116-131); no host round
        trip between ticks.  torch tensors are updated in place; numpy inputs are copied.  Returns q, qdot, fail_ticks,
        iters_total."""
        f = self._frame(link)
        if _is_torch(q):
            import torch
            q, B = self._t_in(q, self.n)
            qd, _ = self._t_in(qdot, self.n, B)
            xt, _ = self._t_in(x_target if x_target.shape[-1] == 12 else x_target[..., :3, :].reshape(-1, 12), 12, B)
            xd, _ = self._t_in(xdot_target, 6, B)
            xi = xdi = None
            if duration > 0:
                xi, _ = self._t_in(x_init if x_init.shape[-1] == 12 else x_init[..., :3, :].reshape(-1, 12), 12, B)
                xdi, _ = self._t_in(xdot_init, 6, B)
            fail = torch.empty(B, dtype=torch.int32, device=q.device)
            its = torch.empty(B, dtype=torch.int32, device=q.device)
            check(lib().drc_batch_rollout_qpik(self._h, B, int(ticks), C.c_double(dt), self._tp(q), self._tp(qd), self._tp(xt), self._tp(xd),
                                               self._tp(xi), self._tp(xdi), C.c_double(t_start), C.c_double(t0), C.c_double(duration), f,
                                               self._tp(fail), self._tp(its), _capi.LAYOUT_AOS, self._stream()), "drc_batch_rollout_qpik")
            return dict(q=q, qdot=qd, fail_ticks=fail, iters_total=its)
        q, B = self._np_in(np.array(q, np.float64), self.n)
        qd, _ = self._np_in(np.array(qdot, np.float64), self.n, B)
        xt, _ = self._np_in(pose12(x_target), 12, B)
        xd, _ = self._np_in(xdot_target, 6, B)
        xi = xdi = None
        if duration > 0:
            xi, _ = self._np_in(pose12(x_init), 12, B)
            xdi, _ = self._np_in(xdot_init, 6, B)
        fail, its = np.zeros(B, np.int32), np.zeros(B, np.int32)
        check(lib().drc_host_rollout_qpik(self._h, B, int(ticks), C.c_double(dt), self._p(q), self._p(qd), self._p(xt), self._p(xd),
                                          self._p(xi), self._p(xdi), C.c_double(t_start), C.c_double(t0), C.c_double(duration), f,
                                          self._pi(fail), self._pi(its)), "drc_host_rollout_qpik")
        return dict(q=q, qdot=qd, fail_ticks=fail, iters_total=its)

    def cycle_qpid_step(self, q, qdot, x_target, xdot_target, link, out=None, status=None, iters=None):
        """updateState + QPIDStep -> torque.  numpy in -> numpy out (host path); torch CUDA in -> torch out (async)."""
        f = self._frame(link)
        if _is_torch(q):
            import torch
            q, B = self._t_in(q, self.n)
            qd, _ = self._t_in(qdot, self.n, B)
            xt, _ = self._t_in(x_target, 12, B)
            xd, _ = self._t_in(xdot_target, 6, B)
            out = torch.empty((B, self.n), dtype=torch.float64, device=q.device) if out is None else out
            status = torch.empty(B, dtype=torch.int32, device=q.device) if status is None else status
            iters = torch.empty(B, dtype=torch.int32, device=q.device) if iters is None else iters
            check(lib().drc_batch_cycle_qpid_step(self._h, B, self._tp(q), self._tp(qd), self._tp(xt), self._tp(xd), f,
                                                  self._tp(out), self._tp(status), self._tp(iters), _capi.LAYOUT_AOS,
                                                  self._stream()), "drc_batch_cycle_qpid_step")
            self._B = B
            return dict(out=out, status=status, iters=iters)
        q, B = self._np_in(q, self.n)
        qd, _ = self._np_in(qdot, self.n, B)
        xt, _ = self._np_in(pose12(x_target), 12, B)
        xd, _ = self._np_in(xdot_target, 6, B)
        out = np.zeros((B, self.n)) if out is None else out
        status = np.zeros(B, np.int32) if status is None else status
        iters = np.zeros(B, np.int32) if iters is None else iters
        check(lib().drc_host_cycle_qpid_step(self._h, B, self._p(q), self._p(qd), self._p(xt), self._p(xd), f, self._p(out),
                                             self._pi(status), self._pi(iters)), "drc_host_cycle_qpid_step")
        self._B = B
        return dict(out=out, status=status, iters=iters)

    # ------------------------------------------------------------------ mobile manipulator (whole-body)
    def moma_update_state(self, q, qdot):
        """q, qdot: (B, dof) joint-ordered vectors (getJointVector of the reference)."""
        if _is_torch(q):
            q, B = self._t_in(q, self.n)
            qd, _ = self._t_in(qdot, self.n, B)
            check(lib().drc_batch_moma_update_state(self._h, B, self._tp(q), self._tp(qd), _capi.LAYOUT_AOS, self._stream()),
                  "drc_batch_moma_update_state")
        else:
            q, B = self._np_in(q, self.n)
            qd, _ = self._np_in(qdot, self.n, B)
            check(lib().drc_host_moma_update_state(self._h, B, self._p(q), self._p(qd)), "drc_host_moma_update_state")
        self._B = B
        return True

    def moma_get_state(self, link, want=("pose", "J", "Jdot", "vel", "M", "Minv", "g", "nle", "mani", "mani_grad", "mani_graddot")):
        B, a, k, f = self._B, self.model.actuated_dof, self.model.mani_dof, self._frame(link)
        shp = dict(pose=(B, 12), J=(B, 6, a), Jdot=(B, 6, a), vel=(B, 6), M=(B, a, a), Minv=(B, a, a), g=(B, a), nle=(B, a),
                   mani=(B,), mani_grad=(B, k), mani_graddot=(B, k))
        o = {key: (np.zeros(shp[key]) if key in want else None) for key in shp}
        if o["mani"] is None and (o["mani_grad"] is not None or o["mani_graddot"] is not None):
            o["mani"] = np.zeros(B)
        check(lib().drc_host_moma_get_state(self._h, B, f, *(self._p(o[key]) for key in ("pose", "J", "Jdot", "vel", "M", "Minv", "g",
                                                                                          "nle", "mani", "mani_grad", "mani_graddot"))),
              "drc_host_moma_get_state")
        return {key: v for key, v in o.items() if v is not None}

    def _moma_qp(self, name, B, args, two):
        a = self.model.actuated_dof
        out, st, it = np.zeros((B, a)), np.zeros(B, np.int32), np.zeros(B, np.int32)
        fn = getattr(lib(), name)
        if two:
            out2 = np.zeros((B, a))
            check(fn(self._h, B, *args, self._p(out), self._p(out2), self._pi(st), self._pi(it)), name)
            return dict(out=out, etadot=out2, status=st, iters=it)
        check(fn(self._h, B, *args, self._p(out), self._pi(st), self._pi(it)), name)
        return dict(out=out, status=st, iters=it)

    def moma_qpik(self, xdot_des, link):
        x, B = self._np_in(xdot_des, 6, self._B)
        return self._moma_qp("drc_host_moma_qpik", B, (self._p(x), self._frame(link)), False)

    def moma_qpik_step(self, x_target, xdot_target, link):
        xt, B = self._np_in(pose12(x_target), 12, self._B)
        xd, _ = self._np_in(xdot_target, 6, B)
        return self._moma_qp("drc_host_moma_qpik_step", B, (self._p(xt), self._p(xd), self._frame(link)), False)

    def moma_qpid(self, xddot_des, link):
        x, B = self._np_in(xddot_des, 6, self._B)
        return self._moma_qp("drc_host_moma_qpid", B, (self._p(x), self._frame(link)), True)

    def moma_qpid_step(self, x_target, xdot_target, link):
        xt, B = self._np_in(pose12(x_target), 12, self._B)
        xd, _ = self._np_in(xdot_target, 6, B)
        return self._moma_qp("drc_host_moma_qpid_step", B, (self._p(xt), self._p(xd), self._frame(link)), True)

    def moma_cycle(self, kind, q, qdot, x_target, xdot_target, link, out=None, out2=None, status=None, iters=None):
        """Fused updateState + whole-body QPIKStep (kind "ik") / QPIDStep (kind "id").  numpy in -> numpy out (host path);
        torch CUDA tensors in -> torch out (device path, asynchronous)."""
        f, a, ident = self._frame(link), self.model.actuated_dof, kind == "id"
        if _is_torch(q):
            import torch
            q, B = self._t_in(q, self.n)
            qd, _ = self._t_in(qdot, self.n, B)
            xt, _ = self._t_in(x_target, 12, B)
            xd, _ = self._t_in(xdot_target, 6, B)
            out = torch.empty((B, a), dtype=torch.float64, device=q.device) if out is None else out
            status = torch.empty(B, dtype=torch.int32, device=q.device) if status is None else status
            iters = torch.empty(B, dtype=torch.int32, device=q.device) if iters is None else iters
            if ident:
                out2 = torch.empty((B, a), dtype=torch.float64, device=q.device) if out2 is None else out2
                check(lib().drc_batch_moma_cycle_qpid_step(self._h, B, self._tp(q), self._tp(qd), self._tp(xt), self._tp(xd), f,
                                                           self._tp(out), self._tp(out2), self._tp(status), self._tp(iters),
                                                           _capi.LAYOUT_AOS, self._stream()), "drc_batch_moma_cycle_qpid_step")
            else:
                check(lib().drc_batch_moma_cycle_qpik_step(self._h, B, self._tp(q), self._tp(qd), self._tp(xt), self._tp(xd), f,
                                                           self._tp(out), self._tp(status), self._tp(iters), _capi.LAYOUT_AOS,
                                                           self._stream()), "drc_batch_moma_cycle_qpik_step")
            self._B = B
            return dict(out=out, etadot=out2, status=status, iters=iters)
        q, B = self._np_in(q, self.n)
        qd, _ = self._np_in(qdot, self.n, B)
        xt, _ = self._np_in(pose12(x_target), 12, B)
        xd, _ = self._np_in(xdot_target, 6, B)
        out = np.zeros((B, a)) if out is None else out
        status = np.zeros(B, np.int32) if status is None else status
        iters = np.zeros(B, np.int32) if iters is None else iters
        if ident:
            out2 = np.zeros((B, a)) if out2 is None else out2
            check(lib().drc_host_moma_cycle_qpid_step(self._h, B, self._p(q), self._p(qd), self._p(xt), self._p(xd), f, self._p(out),
                                                      self._p(out2), self._pi(status), self._pi(iters)), "drc_host_moma_cycle_qpid_step")
        else:
            check(lib().drc_host_moma_cycle_qpik_step(self._h, B, self._p(q), self._p(qd), self._p(xt), self._p(xd), f, self._p(out),
                                                      self._pi(status), self._pi(iters)), "drc_host_moma_cycle_qpik_step")
        self._B = B
        return dict(out=out, etadot=out2, status=status, iters=iters)


def fp64_peak_tflops(device: int = 0) -> float:
    v = C.c_double()
    check(lib().drc_bench_fp64_peak(int(device), C.byref(v)), "drc_bench_fp64_peak")
    return float(v.value)


def device_count() -> int:
    return int(lib().drc_device_count())
