"""TEST INFRASTRUCTURE: ctypes front-end of tests/kernel_emu/libdrc_emu.so (host emulation of the
CUDA kernel bodies; see tests/kernel_emu/emu.cpp).  Never imported by the product package."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent / "kernel_emu"
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
        _LIB = C.CDLL(str(_HERE / "libdrc_emu.so"))
        _LIB.emu_create.restype = C.c_void_p
        _LIB.emu_create2.restype = C.c_void_p
        _LIB.emu_shape_distance.restype = C.c_double
        _LIB.emu_pair_lower_bound.restype = C.c_double
    return _LIB


def _d(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int))


def _c(a, dt=np.float64):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


class Emu:
    def __init__(self, urdf_path: str, srdf_path: str = "", packages_path: str = ""):
        L = lib()
        urdf = Path(urdf_path).read_text()
        srdf = Path(srdf_path).read_text() if srdf_path else ""
        err = C.create_string_buffer(512)
        h = L.emu_create2(urdf.encode(), srdf.encode(), str(Path(urdf_path).resolve().parent).encode(), str(packages_path).encode(), err, 512)
        if not h:
            raise RuntimeError(err.value.decode())
        self.h = C.c_void_p(h)
        self.nv = L.emu_nv(self.h)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().emu_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def frame_id(self, name):
        return lib().emu_frame_id(self.h, name.encode())

    def sizes(self):
        out = np.zeros(6, np.int32)
        nbytes = lib().emu_model_sizes(self.h, _i(out))
        return dict(nv=int(out[0]), ngeom=int(out[1]), npair=int(out[2]), ngroup=int(out[3]), nframes=int(out[4]),
                    skipped=int(out[5]), dev_bytes=int(nbytes))

    def mesh_info(self):
        ng = self.sizes()["ngeom"]
        out, vn = np.zeros(2, np.int32), np.zeros(ng, np.int32)
        lib().emu_mesh_info(self.h, _i(out), _i(vn), None)
        hull = np.zeros((int(out[1]), 3))
        if out[1]:
            lib().emu_mesh_info(self.h, _i(out), _i(vn), _d(hull))
        return dict(mesh_geoms=int(out[0]), hull_vertices=int(out[1]), vert_n=vn, hull=hull)

    def model_arrays(self):
        s = self.sizes()
        n, ng, npair, nf = s["nv"], s["ngeom"], s["npair"], s["nframes"]
        a = dict(parent=np.zeros(n, np.int32), jtype=np.zeros(n, np.int32), axis=np.zeros((n, 3)), jR=np.zeros((n, 3, 3)),
                 jp=np.zeros((n, 3)), mass=np.zeros(n), com=np.zeros((n, 3)), inertia6=np.zeros((n, 6)), q_lo=np.zeros(n),
                 q_hi=np.zeros(n), v_lim=np.zeros(n), geom_type=np.zeros(ng, np.int32), geom_parent=np.zeros(ng, np.int32),
                 geom_prm=np.zeros((ng, 3)), geom_R=np.zeros((ng, 3, 3)), geom_p=np.zeros((ng, 3)),
                 pairs=np.zeros((npair, 2), np.int32), frame_parent=np.zeros(nf, np.int32), frame_R=np.zeros((nf, 3, 3)),
                 frame_p=np.zeros((nf, 3)))
        lib().emu_model_arrays(self.h, _i(a["parent"]), _i(a["jtype"]), _d(a["axis"]), _d(a["jR"]), _d(a["jp"]), _d(a["mass"]),
                               _d(a["com"]), _d(a["inertia6"]), _d(a["q_lo"]), _d(a["q_hi"]), _d(a["v_lim"]),
                               _i(a["geom_type"]), _i(a["geom_parent"]), _d(a["geom_prm"]), _d(a["geom_R"]), _d(a["geom_p"]),
                               _i(a["pairs"]), _i(a["frame_parent"]), _d(a["frame_R"]), _d(a["frame_p"]))
        return a

    def set_params(self, kp_task=None, kv_task=None, kp_joint=None, kv_joint=None, adaptive_rho_interval=-1, max_iter=0,
                   gjk_tol=0.0):
        lib().emu_set_params(self.h, _d(_c(kp_task)), _d(_c(kv_task)), _d(_c(kp_joint)), _d(_c(kv_joint)),
                             C.c_int(adaptive_rho_interval), C.c_int(max_iter), C.c_double(gjk_tol))

    def update_and_get(self, q, qd, frame):
        q, qd = _c(q).reshape(-1, self.nv), _c(qd).reshape(-1, self.nv)
        B, n = q.shape
        o = dict(pose=np.zeros((B, 12)), J=np.zeros((B, 6, n)), Jdot=np.zeros((B, 6, n)), vel=np.zeros((B, 6)),
                 M=np.zeros((B, n, n)), Minv=np.zeros((B, n, n)), g=np.zeros((B, n)), nle=np.zeros((B, n)),
                 oMi=np.zeros((B, 12 * n)), mani=np.zeros(B), mgrad=np.zeros((B, n)), mgraddot=np.zeros((B, n)))
        rc = lib().emu_update_and_get(self.h, C.c_int(frame), C.c_int(B), _d(q), _d(qd), _d(o["pose"]), _d(o["J"]),
                                      _d(o["Jdot"]), _d(o["vel"]), _d(o["M"]), _d(o["Minv"]), _d(o["g"]), _d(o["nle"]),
                                      _d(o["oMi"]), _d(o["mani"]), _d(o["mgrad"]), _d(o["mgraddot"]))
        assert rc == 0
        return o

    def min_distance(self, q, qd):
        q, qd = _c(q).reshape(-1, self.nv), _c(qd).reshape(-1, self.nv)
        B, n = q.shape
        o = dict(d=np.zeros(B), grad=np.zeros((B, n)), grad_dot=np.zeros((B, n)), pair=np.zeros(B, np.int32),
                 witness=np.zeros((B, 6)))
        o["nepa"] = lib().emu_min_distance(self.h, C.c_int(B), _d(q), _d(qd), _d(o["d"]), _d(o["grad"]), _d(o["grad_dot"]),
                                           _i(o["pair"]), _d(o["witness"]))
        return o

    def cycle(self, mode, q, qd, x_target, xdot_target, frame, want_records=False):
        q, qd = _c(q).reshape(-1, self.nv), _c(qd).reshape(-1, self.nv)
        B, n = q.shape
        xt = None if x_target is None else _c(x_target).reshape(B, 12)
        xd = _c(xdot_target).reshape(B, 6)
        ku, nr = (2, 2) if mode < 2 else (4, 2 + n)
        o = dict(out=np.zeros((B, n)), status=np.zeros(B, np.int32), iters=np.zeros(B, np.int32),
                 x=np.zeros((B, n * (1 + ku) + nr)))
        rec = np.zeros((B, lib().emu_qp_stride(C.c_int(mode)))) if want_records else None
        rc = lib().emu_cycle(self.h, C.c_int(mode), C.c_int(frame), C.c_int(B), _d(q), _d(qd), _d(xt), _d(xd), _d(o["out"]),
                             _i(o["status"]), _i(o["iters"]), _d(o["x"]), _d(rec))
        assert rc == 0
        if want_records:
            o["records"] = rec
        return o

    def cycle_warm(self, mode, q, qd, x_target, xdot_target, frame, warm_x, warm_y):
        """QPIKStep warm started from (warm_x, warm_y) (structured order, updated in place; zeros = cold start)"""
        q, qd = _c(q).reshape(-1, self.nv), _c(qd).reshape(-1, self.nv)
        B, n = q.shape
        xt, xd = _c(x_target).reshape(B, 12), _c(xdot_target).reshape(B, 6)
        o = dict(out=np.zeros((B, n)), status=np.zeros(B, np.int32), iters=np.zeros(B, np.int32))
        assert warm_x.shape == (B, 3 * n + 2) and warm_y.shape == (B, 5 * n + 4)
        rc = lib().emu_cycle_warm(self.h, C.c_int(mode), C.c_int(frame), C.c_int(B), _d(q), _d(qd), _d(xt), _d(xd), _d(o["out"]),
                                  _i(o["status"]), _i(o["iters"]), _d(warm_x), _d(warm_y))
        assert rc == 0
        return o

    def cycle_xy(self, mode, q, qd, x_target, xdot_target, frame):
        """cycle() plus the dual vector: returns the same dictionary as engine.Context.qp_debug next to out / status / iters."""
        q, qd = _c(q).reshape(-1, self.nv), _c(qd).reshape(-1, self.nv)
        B, n = q.shape
        xt = None if x_target is None else _c(x_target).reshape(B, 12)
        xd = _c(xdot_target).reshape(B, 6)
        ku, nr = (2, 2) if mode < 2 else (4, 2 + n)
        o = dict(out=np.zeros((B, n)), status=np.zeros(B, np.int32), iters=np.zeros(B, np.int32),
                 x=np.zeros((B, n * (1 + ku) + nr)), y=np.zeros((B, n * (1 + 2 * ku) + 2 * nr)), nc=n, ku=ku, nr=nr)
        rc = lib().emu_cycle_xy(self.h, C.c_int(mode), C.c_int(frame), C.c_int(B), _d(q), _d(qd), _d(xt), _d(xd), _d(o["out"]),
                                _i(o["status"]), _i(o["iters"]), _d(o["x"]), _d(o["y"]))
        assert rc == 0
        return o

    def taskspace(self, mode, q, qd, x_target, xdot_target, frame, aux=None, aux2=None):
        q, qd = _c(q).reshape(-1, self.nv), _c(qd).reshape(-1, self.nv)
        B, n = q.shape
        xt = None if x_target is None else _c(x_target).reshape(B, 12)
        xd = None if xdot_target is None else _c(xdot_target).reshape(B, 6)
        a1 = None if aux is None else _c(aux).reshape(B, n)
        a2 = None if aux2 is None else _c(aux2).reshape(B, n)
        out = np.zeros((B, n))
        rc = lib().emu_taskspace(self.h, C.c_int(mode), C.c_int(frame), C.c_int(B), _d(q), _d(qd), _d(xt), _d(xd), _d(a1),
                                 _d(a2), _d(out))
        assert rc == 0
        return out

    def shape_distance(self, ta, prm_a, Ta, tb, prm_b, Tb):
        """Product narrow phase (closed form / GJK / EPA) on one pair; poses are 4x4 or 12-vectors."""
        def p12(T):
            T = np.asarray(T, np.float64)
            return np.ascontiguousarray(T.reshape(-1)[:12] if T.size == 12 else T[:3, :].reshape(12))
        wa, wb, info = np.zeros(3), np.zeros(3), np.zeros(2, np.int32)
        d = lib().emu_shape_distance(self.h, C.c_int(ta), _d(_c(prm_a)), _d(p12(Ta)), C.c_int(tb), _d(_c(prm_b)), _d(p12(Tb)),
                                     _d(wa), _d(wb), _i(info))
        return float(d), wa, wb, info


class MomaEmu(Emu):
    """Kernel bodies of the mobile-manipulator path (rows a15-a18) on the CPU."""

    def __init__(self, urdf_path, srdf_path, kin: dict, joint_idx: dict, actuator_idx: dict):
        super().__init__(urdf_path, srdf_path)
        t = kin["type"] if isinstance(kin["type"], int) else dict(Differential=0, Mecanum=1, Caster=2)[kin["type"]]
        pos = np.asarray(kin.get("base2wheel_positions", np.zeros((0, 2))), np.float64).reshape(-1, 2)
        w = 2 if t == 0 else (len(kin["roller_angles"]) if t == 1 else 2 * len(pos))
        self.w, self.act, self.mani = w, self.nv - 3, self.nv - 3 - w
        ra = _c(kin.get("roller_angles", np.zeros(w))) if t == 1 else None
        ba = _c(kin.get("base2wheel_angles", np.zeros(w))) if t == 1 else None
        bx = _c(pos[:, 0]) if len(pos) else None
        by = _c(pos[:, 1]) if len(pos) else None
        err = C.create_string_buffer(512)
        rc = lib().emu_moma_attach(self.h, C.c_int(t), C.c_double(kin.get("wheel_radius", 0.0)), C.c_double(kin.get("base_width", 0.0)),
                                   C.c_double(kin.get("wheel_offset", 0.0)), C.c_int(w), _d(ra), _d(bx), _d(by), _d(ba),
                                   C.c_int(joint_idx["virtual_start"]), C.c_int(joint_idx["mani_start"]), C.c_int(joint_idx["mobi_start"]),
                                   C.c_int(actuator_idx["mani_start"]), C.c_int(actuator_idx["mobi_start"]), err, 512)
        if rc:
            raise RuntimeError(err.value.decode())

    def base_jacobian(self):
        J = np.zeros((3, self.w))
        lib().emu_moma_base_jacobian(self.h, _d(J))
        return J

    def moma_state(self, q, qd, frame):
        q, qd = _c(q).reshape(-1, self.nv), _c(qd).reshape(-1, self.nv)
        B, a, k = q.shape[0], self.act, self.mani
        o = dict(pose=np.zeros((B, 12)), J=np.zeros((B, 6, a)), Jdot=np.zeros((B, 6, a)), vel=np.zeros((B, 6)), M=np.zeros((B, a, a)),
                 Minv=np.zeros((B, a, a)), g=np.zeros((B, a)), nle=np.zeros((B, a)), mani=np.zeros(B), mani_grad=np.zeros((B, k)),
                 mani_graddot=np.zeros((B, k)))
        rc = lib().emu_moma_state(self.h, C.c_int(frame), C.c_int(B), _d(q), _d(qd), _d(o["pose"]), _d(o["J"]), _d(o["Jdot"]),
                                  _d(o["vel"]), _d(o["M"]), _d(o["Minv"]), _d(o["g"]), _d(o["nle"]), _d(o["mani"]), _d(o["mani_grad"]),
                                  _d(o["mani_graddot"]))
        assert rc == 0, rc
        return o

    def moma_cycle(self, mode, q, qd, x_target, xdot_target, frame):
        q, qd = _c(q).reshape(-1, self.nv), _c(qd).reshape(-1, self.nv)
        B = q.shape[0]
        xt = None if x_target is None else _c(x_target).reshape(B, 12)
        xd = _c(xdot_target).reshape(B, 6)
        o = dict(out=np.zeros((B, self.act)), out2=np.zeros((B, self.act)), status=np.zeros(B, np.int32), iters=np.zeros(B, np.int32))
        rc = lib().emu_moma_cycle(self.h, C.c_int(mode), C.c_int(frame), C.c_int(B), _d(q), _d(qd), _d(xt), _d(xd), _d(o["out"]),
                                  _d(o["out2"]), _i(o["status"]), _i(o["iters"]))
        assert rc == 0, rc
        return o


def pair_lower_bound(ta, prm_a, Ta, tb, prm_b, Tb):
    def p12(T):
        T = np.asarray(T, np.float64)
        return np.ascontiguousarray(T.reshape(-1)[:12] if T.size == 12 else T[:3, :].reshape(12))
    return float(lib().emu_pair_lower_bound(C.c_int(ta), _d(_c(prm_a)), _d(p12(Ta)), C.c_int(tb), _d(_c(prm_b)), _d(p12(Tb))))


def mobile_base(kin: dict, fk: bool, wheel_pos, vec, saturate: bool = False):
    """Bodies of k_mobile_fk / k_mobile_ik (csrc/drc_mobile.h) on the CPU; same interface as oracle.c_oracle.mobile_base."""
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
    rc = lib().emu_mobile(C.c_int(t), C.c_double(kin.get("wheel_radius", 0.0)), C.c_double(kin.get("base_width", 0.0)),
                          C.c_double(kin.get("wheel_offset", 0.0)), C.c_double(kin.get("max_lin_speed", 0.0)),
                          C.c_double(kin.get("max_ang_speed", 0.0)), C.c_int(w), _d(ra), _d(bx), _d(by), _d(ba),
                          C.c_int(1 if fk else 0), C.c_int(1 if saturate else 0), C.c_int(B), _d(wp), _d(v), _d(J), _d(out))
    assert rc == 0, rc
    return J, out
