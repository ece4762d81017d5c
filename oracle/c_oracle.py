"""ORACLE (test infrastructure) -- ctypes front-end of oracle/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  PARITY UNPINNED: see oracle/README.md.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from . import urdf_model

_HERE = Path(__file__).resolve().parent
_LIB = None

QP_SOLVED, QP_MAX_ITER, QP_PRIMAL_INFEASIBLE, QP_DUAL_INFEASIBLE, QP_NON_CVX, QP_SOLVED_INACCURATE = 1, 2, 3, 4, 5, 6


def build(force: bool = False) -> Path:
    so = _HERE / "liboracle.so"
    srcs = list((_HERE / "src").glob("*"))
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = _HERE / "liboracle.so"
        if not so.exists():
            build()
        _LIB = C.CDLL(str(so))
        _LIB.orc_model_create.restype = C.c_void_p
        _LIB.orc_shape_distance.restype = C.c_double
    return _LIB


def _d(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int))


def _c(a, dtype=np.float64):
    return np.ascontiguousarray(a, dtype=dtype)


def pose12(T) -> np.ndarray:
    """(…,4,4) or (…,3,4) homogeneous -> (…,12) top three rows, row-major."""
    T = np.asarray(T, np.float64)
    return np.ascontiguousarray(T[..., :3, :].reshape(T.shape[:-2] + (12,)))


def pose44(p12) -> np.ndarray:
    p12 = np.asarray(p12, np.float64)
    out = np.zeros(p12.shape[:-1] + (4, 4))
    out[..., :3, :] = p12.reshape(p12.shape[:-1] + (3, 4))
    out[..., 3, 3] = 1.0
    return out


class Oracle:
    """CPU restatement of the reference hot path for one robot model."""

    def __init__(self, urdf_path: str, srdf_path: str = "", threads: int = 1, packages_path: str = ""):
        self.model = m = urdf_model.load(urdf_path, srdf_path, packages_path)
        self.nv = m.nv
        L = lib()
        self._keep = [_c(m.parent, np.int32), _c(m.jtype, np.int32), _c(m.axis), _c(m.jR), _c(m.jp), _c(m.mass),
                      _c(m.com), _c(m.inertia), _c(m.q_lo), _c(m.q_hi), _c(m.v_lim), _c(m.frame_parent, np.int32),
                      _c(m.frame_R), _c(m.frame_p), _c(m.geom_type, np.int32), _c(m.geom_param),
                      _c(m.geom_parent, np.int32), _c(m.geom_R), _c(m.geom_p), _c(m.pairs, np.int32), _c(m.gravity)]
        k = self._keep
        self.h = C.c_void_p(L.orc_model_create(
            C.c_int(m.nv), _i(k[0]), _i(k[1]), _d(k[2]), _d(k[3]), _d(k[4]), _d(k[5]), _d(k[6]), _d(k[7]), _d(k[8]),
            _d(k[9]), _d(k[10]), C.c_int(len(m.frame_names)), _i(k[11]), _d(k[12]), _d(k[13]),
            C.c_int(len(m.geom_names)), _i(k[14]), _d(k[15]), _i(k[16]), _d(k[17]), _d(k[18]),
            C.c_int(len(m.pairs)), _i(k[19]), _d(k[20])))
        if not self.h:
            raise RuntimeError("orc_model_create failed")
        if len(m.hull):  # mesh collision geometry: hull vertices per geometry
            hv, ho, hn = _c(m.hull), _c(m.hull_off, np.int32), _c(m.hull_n, np.int32)
            L.orc_model_set_hulls(self.h, C.c_int(len(hv)), _d(hv), _i(ho), _i(hn))
        self.set_threads(threads)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().orc_model_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # -- configuration
    def set_threads(self, t: int):
        lib().orc_set_threads(self.h, C.c_int(int(t)))

    def set_fresh_workspace(self, on: bool):
        """CPU-baseline mode: allocate and release the whole workspace every control cycle, as the reference does (a new
        OsqpEigen::Solver per solveQP, QP_base.h:143-177; a new pinocchio::Data per getManipulability, robot_data.cpp:542)."""
        lib().orc_set_fresh_workspace(self.h, C.c_int(int(on)))

    def set_task_gains(self, kp, kv):
        lib().orc_set_task_gains(self.h, _d(_c(kp)), _d(_c(kv)))

    def set_joint_gains(self, kp, kv):
        lib().orc_set_joint_gains(self.h, _d(_c(kp)), _d(_c(kv)))

    def set_qp_settings(self, rho=0.1, sigma=1e-6, alpha=1.6, eps_abs=1e-3, eps_rel=1e-3, eps_prim_inf=1e-4,
                        eps_dual_inf=1e-4, max_iter=4000, check_termination=25, scaling=10, adaptive_rho=1,
                        adaptive_rho_interval=50, adaptive_rho_tolerance=5.0):
        lib().orc_set_qp_settings(self.h, C.c_double(rho), C.c_double(sigma), C.c_double(alpha), C.c_double(eps_abs),
                                  C.c_double(eps_rel), C.c_double(eps_prim_inf), C.c_double(eps_dual_inf),
                                  C.c_int(max_iter), C.c_int(check_termination), C.c_int(scaling), C.c_int(adaptive_rho),
                                  C.c_int(adaptive_rho_interval), C.c_double(adaptive_rho_tolerance))

    def set_geom_params(self, gjk_tol=1e-10, gjk_max_iter=128, epa_tol=1e-6, epa_max_iter=96):
        lib().orc_set_geom_params(self.h, C.c_double(gjk_tol), C.c_int(gjk_max_iter), C.c_double(epa_tol),
                                  C.c_int(epa_max_iter))

    def frame_id(self, name: str) -> int:
        return self.model.frame_id(name)

    # -- hot path pieces (batch-major arrays)
    def update_state(self, q, qd, frame: int):
        q, qd = _c(q).reshape(-1, self.nv), _c(qd).reshape(-1, self.nv)
        B, n = q.shape
        out = dict(pose=np.zeros((B, 12)), J=np.zeros((B, 6, n)), Jdot=np.zeros((B, 6, n)), M=np.zeros((B, n, n)),
                   Minv=np.zeros((B, n, n)), g=np.zeros((B, n)), nle=np.zeros((B, n)), oMi=np.zeros((B, n, 12)))
        lib().orc_update_state(self.h, C.c_int(B), _d(q), _d(qd), C.c_int(frame), _d(out["pose"]), _d(out["J"]),
                               _d(out["Jdot"]), _d(out["M"]), _d(out["Minv"]), _d(out["g"]), _d(out["nle"]),
                               _d(out["oMi"]))
        return out

    def manipulability(self, q, qd, frame: int, with_graddot=True):
        q, qd = _c(q).reshape(-1, self.nv), _c(qd).reshape(-1, self.nv)
        B, n = q.shape
        m, g, gd = np.zeros(B), np.zeros((B, n)), np.zeros((B, n))
        lib().orc_manipulability(self.h, C.c_int(B), _d(q), _d(qd), C.c_int(frame), C.c_int(int(with_graddot)), _d(m),
                                 _d(g), _d(gd))
        return m, g, gd

    def min_distance(self, q, qd, with_graddot=True):
        q, qd = _c(q).reshape(-1, self.nv), _c(qd).reshape(-1, self.nv)
        B, n = q.shape
        d, g, gd = np.zeros(B), np.zeros((B, n)), np.zeros((B, n))
        pair, pa, pb = np.zeros(B, np.int32), np.zeros((B, 3)), np.zeros((B, 3))
        lib().orc_min_distance(self.h, C.c_int(B), _d(q), _d(qd), C.c_int(int(with_graddot)), _d(d), _d(g), _d(gd),
                               _i(pair), _d(pa), _d(pb))
        return dict(d=d, grad=g, grad_dot=gd, pair=pair, pa=pa, pb=pb)

    def pair_distances(self, q):
        q = _c(q).reshape(self.nv)
        npair = len(self.model.pairs)
        d, pa, pb, it = np.zeros(npair), np.zeros((npair, 3)), np.zeros((npair, 3)), np.zeros(npair, np.int32)
        lib().orc_pair_distances(self.h, _d(q), _d(d), _d(pa), _d(pb), _i(it))
        return d, pa, pb, it

    def shape_distance(self, ta, prm_a, Ta, tb, prm_b, Tb):
        wa, wb, it = np.zeros(3), np.zeros(3), np.zeros(2, np.int32)
        d = lib().orc_shape_distance(self.h, C.c_int(ta), _d(_c(prm_a)), _d(pose12(Ta)), C.c_int(tb), _d(_c(prm_b)),
                                     _d(pose12(Tb)), _d(wa), _d(wb), _i(it))
        return float(d), wa, wb, it

    def qp_sizes(self, kind: int):
        nx, nc = C.c_int(), C.c_int()
        lib().orc_qp_sizes(self.h, C.c_int(kind), C.byref(nx), C.byref(nc))
        return nx.value, nc.value

    def build_qp(self, kind: int, q, qd, des, frame: int):
        nx, nc = self.qp_sizes(kind)
        P, qv, A, l, u = np.zeros((nx, nx)), np.zeros(nx), np.zeros((nc, nx)), np.zeros(nc), np.zeros(nc)
        lib().orc_build_qp(self.h, C.c_int(kind), _d(_c(q)), _d(_c(qd)), _d(_c(des)), C.c_int(frame), _d(P), _d(qv),
                           _d(A), _d(l), _d(u))
        return P, qv, A, l, u

    def solve_qp(self, P, qv, A, l, u):
        P, qv, A, l, u = _c(P), _c(qv), _c(A), _c(l), _c(u)
        n, m = qv.size, l.size
        x, y, it, info = np.zeros(n), np.zeros(m), C.c_int(), np.zeros(4)
        st = lib().orc_solve_qp(self.h, C.c_int(n), C.c_int(m), _d(P), _d(qv), _d(A), _d(l), _d(u), _d(x), _d(y),
                                C.byref(it), _d(info))
        return dict(status=st, x=x, y=y, iters=it.value, pri_res=info[0], dua_res=info[1], rho=info[2],
                    rho_updates=int(info[3]))

    def cycle(self, mode: int, q, qd, x_target, xdot_target, frame: int, want_x=False, want_y=False):
        """mode 0 QPIK(xdot_des) 1 QPIKStep 2 QPID(xddot_des) 3 QPIDStep. x_target: (B,12) or None.
        want_x / want_y: primal / dual vector of the QP in the reference's order (QP_base.h:204-226:
        rows = [bounds; inequalities; equalities])."""
        q, qd = _c(q).reshape(-1, self.nv), _c(qd).reshape(-1, self.nv)
        B, n = q.shape
        xt = None if x_target is None else _c(x_target).reshape(B, 12)
        xd = _c(xdot_target).reshape(B, 6)
        out, st, it = np.zeros((B, n)), np.zeros(B, np.int32), np.zeros(B, np.int32)
        nx, nc = self.qp_sizes(0 if mode <= 1 else 1)
        qx = np.zeros((B, nx)) if want_x else None
        qy = np.zeros((B, nc)) if want_y else None
        lib().orc_cycle_xy(self.h, C.c_int(mode), C.c_int(B), _d(q), _d(qd), _d(xt), _d(xd), C.c_int(frame), _d(out),
                           _i(st), _i(it), _d(qx), _d(qy))
        r = dict(out=out, status=st, iters=it)
        if want_x:
            r["x"] = qx
        if want_y:
            r["y"] = qy
        return r

    def cycle_warm(self, mode: int, q, qd, x_target, xdot_target, frame: int, warm_x, warm_y):
        """Step cycle (mode 1 / 3) warm started from (warm_x, warm_y) = the previous tick's primal / dual solution (oracle order);
        both arrays are updated in place with this tick's.  Zeros = cold start.  An extension (QP_base.h:146 never warm starts)."""
        q, qd = _c(q).reshape(-1, self.nv), _c(qd).reshape(-1, self.nv)
        B, n = q.shape
        xt, xd = _c(x_target).reshape(B, 12), _c(xdot_target).reshape(B, 6)
        out, st, it = np.zeros((B, n)), np.zeros(B, np.int32), np.zeros(B, np.int32)
        nx, nc = self.qp_sizes(0 if mode <= 1 else 1)
        assert warm_x.shape == (B, nx) and warm_y.shape == (B, nc) and warm_x.flags.c_contiguous and warm_y.flags.c_contiguous
        lib().orc_cycle_warm(self.h, C.c_int(mode), C.c_int(B), _d(q), _d(qd), _d(xt), _d(xd), C.c_int(frame), _d(out),
                             _i(st), _i(it), C.c_int(nx), C.c_int(nc), _d(warm_x), _d(warm_y))
        return dict(out=out, status=st, iters=it)

    def desired_task(self, mode: int, q, qd, x_target, xdot_target, frame: int):
        """desired task-space signal the Step controllers hand to the QP (robot_controller.cpp:292-300, 335-345):
        Kp e + Kv edot (modes 1, 3)."""
        q, qd = _c(q).reshape(-1, self.nv), _c(qd).reshape(-1, self.nv)
        B = q.shape[0]
        des = np.zeros((B, 6))
        lib().orc_desired_task(self.h, C.c_int(mode), C.c_int(B), _d(q), _d(qd), _d(_c(x_target).reshape(B, 12)),
                               _d(_c(xdot_target).reshape(B, 6)), C.c_int(frame), _d(des))
        return des

    def taskspace(self, mode: int, q, qd, x_target, xdot_target, frame: int, null_vec=None):
        """mode 0 CLIKStep, 1 OSFStep, 2 OSF(xddot given in xdot_target)."""
        q, qd = _c(q).reshape(-1, self.nv), _c(qd).reshape(-1, self.nv)
        B, n = q.shape
        xt = None if x_target is None else _c(x_target).reshape(B, 12)
        xd = _c(xdot_target).reshape(B, 6)
        nvv = None if null_vec is None else _c(null_vec).reshape(B, n)
        out = np.zeros((B, n))
        lib().orc_taskspace(self.h, C.c_int(mode), C.c_int(B), _d(q), _d(qd), _d(xt), _d(xd), _d(nvv), C.c_int(frame),
                            _d(out))
        return out

    def joint_torque_step(self, q, qd, q_t, qd_t):
        q, qd = _c(q).reshape(-1, self.nv), _c(qd).reshape(-1, self.nv)
        q_t, qd_t = _c(q_t).reshape(q.shape), _c(qd_t).reshape(q.shape)
        tau = np.zeros_like(q)
        lib().orc_joint_torque_step(self.h, C.c_int(q.shape[0]), _d(q), _d(qd), _d(q_t), _d(qd_t), _d(tau))
        return tau


def mobile_base(kin: dict, fk: bool, wheel_pos, vec, saturate: bool = False):
    """Mobile::RobotData (fk: J (B,3,w), base velocity from wheel velocities) / Mobile::RobotController (not fk: J (B,w,3),
    wheel velocities from a base velocity; VelocityCommand's saturation when `saturate`) restated in oracle/src/omoma.h.
    kin: dict like Mobile::KinematicParam (type, wheel_radius, base_width, wheel_offset, max_lin_speed, max_ang_speed,
    roller_angles, base2wheel_positions, base2wheel_angles)."""
    t = kin["type"] if isinstance(kin["type"], int) else dict(Differential=0, Mecanum=1, Caster=2)[kin["type"]]
    pos = np.asarray(kin.get("base2wheel_positions", np.zeros((0, 2))), np.float64).reshape(-1, 2)
    w = 2 if t == 0 else (len(kin["roller_angles"]) if t == 1 else 2 * len(pos))
    wp = _c(wheel_pos).reshape(-1, w)
    B = wp.shape[0]
    v = None if vec is None else _c(vec).reshape(B, w if fk else 3)
    J = np.zeros((B, 3, w) if fk else (B, w, 3))
    out = np.zeros((B, 3) if fk else (B, w))
    ra = _c(kin.get("roller_angles", np.zeros(w))) if t == 1 else None
    ba = _c(kin.get("base2wheel_angles", np.zeros(w))) if t == 1 else None
    bx = _c(pos[:, 0]) if len(pos) else None
    by = _c(pos[:, 1]) if len(pos) else None
    lib().orc_mobile_base(C.c_int(t), C.c_double(kin.get("wheel_radius", 0.0)), C.c_double(kin.get("base_width", 0.0)),
                          C.c_double(kin.get("wheel_offset", 0.0)), C.c_double(kin.get("max_lin_speed", 0.0)),
                          C.c_double(kin.get("max_ang_speed", 0.0)), C.c_int(w), _d(ra), _d(bx), _d(by), _d(ba),
                          C.c_int(1 if fk else 0), C.c_int(1 if saturate else 0), C.c_int(B), _d(wp), _d(v), _d(J), _d(out))
    return J, out


class MomaOracle(Oracle):
    """Mobile-manipulator restatement (oracle/src/omoma.h).  kin: dict(type, wheel_radius, base_width, wheel_offset,
    roller_angles, base2wheel_positions [(x, y)...], base2wheel_angles) like the reference's Mobile::KinematicParam
    (type_define.h:58-72); joint_idx / actuator_idx like JointIndex / ActuatorIndex."""

    DRIVE = dict(Differential=0, Mecanum=1, Caster=2)

    def __init__(self, urdf_path, srdf_path, kin: dict, joint_idx: dict, actuator_idx: dict, threads: int = 1):
        super().__init__(urdf_path, srdf_path, threads)
        t = kin["type"] if isinstance(kin["type"], int) else self.DRIVE[kin["type"]]
        pos = np.asarray(kin.get("base2wheel_positions", np.zeros((0, 2))), np.float64).reshape(-1, 2)
        if t == 0:
            w = 2
        elif t == 1:
            w = len(kin["roller_angles"])
        else:
            w = 2 * len(pos)
        self.w, self.act, self.mani = w, self.nv - 3, self.nv - 3 - w
        self.joint_idx, self.actuator_idx = dict(joint_idx), dict(actuator_idx)
        ra = _c(kin.get("roller_angles", np.zeros(w))) if t == 1 else None
        ba = _c(kin.get("base2wheel_angles", np.zeros(w))) if t == 1 else None
        bx = _c(pos[:, 0]) if len(pos) else None
        by = _c(pos[:, 1]) if len(pos) else None
        lib().orc_model_set_moma(self.h, C.c_int(t), C.c_double(kin.get("wheel_radius", 0.0)), C.c_double(kin.get("base_width", 0.0)),
                                 C.c_double(kin.get("wheel_offset", 0.0)), C.c_int(w), _d(ra), _d(bx), _d(by), _d(ba),
                                 C.c_int(joint_idx["virtual_start"]), C.c_int(joint_idx["mani_start"]),
                                 C.c_int(joint_idx["mobi_start"]), C.c_int(actuator_idx["mani_start"]),
                                 C.c_int(actuator_idx["mobi_start"]))

    def mobile_state(self, wheel_pos, wheel_vel):
        wp, wv = _c(wheel_pos).reshape(-1, self.w), _c(wheel_vel).reshape(-1, self.w)
        B = wp.shape[0]
        J, bv = np.zeros((B, 3, self.w)), np.zeros((B, 3))
        lib().orc_mobile_state(self.h, C.c_int(B), _d(wp), _d(wv), _d(J), _d(bv))
        return J, bv

    def moma_update_state(self, q, qd, frame: int):
        q, qd = _c(q).reshape(-1, self.nv), _c(qd).reshape(-1, self.nv)
        B, n, a, k = q.shape[0], self.nv, self.act, self.mani
        o = dict(S=np.zeros((B, n, a)), M=np.zeros((B, a, a)), Minv=np.zeros((B, a, a)), g=np.zeros((B, a)), nle=np.zeros((B, a)),
                 J=np.zeros((B, 6, a)), Jdot=np.zeros((B, 6, a)), mani=np.zeros(B), mani_grad=np.zeros((B, k)),
                 mani_graddot=np.zeros((B, k)))
        lib().orc_moma_update_state(self.h, C.c_int(B), _d(q), _d(qd), C.c_int(frame), _d(o["S"]), _d(o["M"]), _d(o["Minv"]),
                                    _d(o["g"]), _d(o["nle"]), _d(o["J"]), _d(o["Jdot"]), _d(o["mani"]), _d(o["mani_grad"]),
                                    _d(o["mani_graddot"]))
        return o

    def moma_qp_sizes(self, kind: int):
        nx, nc = C.c_int(), C.c_int()
        lib().orc_moma_qp_sizes(self.h, C.c_int(kind), C.byref(nx), C.byref(nc))
        return nx.value, nc.value

    def moma_build_qp(self, kind: int, q, qd, des, frame: int):
        nx, nc = self.moma_qp_sizes(kind)
        P, qv, A, l, u = np.zeros((nx, nx)), np.zeros(nx), np.zeros((nc, nx)), np.zeros(nc), np.zeros(nc)
        lib().orc_moma_build_qp(self.h, C.c_int(kind), _d(_c(q)), _d(_c(qd)), _d(_c(des)), C.c_int(frame), _d(P), _d(qv), _d(A),
                                _d(l), _d(u))
        return P, qv, A, l, u

    def moma_cycle(self, mode: int, q, qd, x_target, xdot_target, frame: int, want_xy=False):
        """mode 0 QPIK 1 QPIKStep 2 QPID 3 QPIDStep -> out (B, act): eta* | tau*; out2: eta_dot* (modes 2/3).
        want_xy: also the primal / dual vectors of the QP in the reference's order."""
        q, qd = _c(q).reshape(-1, self.nv), _c(qd).reshape(-1, self.nv)
        B = q.shape[0]
        xt = None if x_target is None else _c(x_target).reshape(B, 12)
        xd = _c(xdot_target).reshape(B, 6)
        out, out2 = np.zeros((B, self.act)), np.zeros((B, self.act))
        st, it = np.zeros(B, np.int32), np.zeros(B, np.int32)
        nx, nc = self.moma_qp_sizes(0 if mode <= 1 else 1)
        qx = np.zeros((B, nx)) if want_xy else None
        qy = np.zeros((B, nc)) if want_xy else None
        lib().orc_moma_cycle_xy(self.h, C.c_int(mode), C.c_int(B), _d(q), _d(qd), _d(xt), _d(xd), C.c_int(frame), _d(out),
                                _d(out2), _i(st), _i(it), _d(qx), _d(qy))
        r = dict(out=out, out2=out2, status=st, iters=it)
        if want_xy:
            r["x"], r["y"] = qx, qy
        return r


def task_space_cubic(x_target, xdot_target, x_init, xdot_init, t, t0, dur):
    xd, xdd = np.zeros(12), np.zeros(6)
    lib().orc_task_space_cubic(_d(pose12(x_target)), _d(_c(xdot_target)), _d(pose12(x_init)), _d(_c(xdot_init)),
                               C.c_double(t), C.c_double(t0), C.c_double(dur), _d(xd), _d(xdd))
    return pose44(xd), xdd


def pinv(A):
    A = _c(A)
    m, n = A.shape
    out = np.zeros((n, m))
    lib().orc_pinv(_d(A), C.c_int(m), C.c_int(n), _d(out))
    return out
