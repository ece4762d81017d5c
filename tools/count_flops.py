#!/usr/bin/env python3
"""Algorithmic fp64 flop model of every benchmark workload -> bench/flops_model.json (loaded by bench.py for the roofline).

The product's kernel bodies (the DRC_HD routines nvcc compiles into the kernels) are compiled by tools/flopcount/flopcount.cpp
with a flop-counting scalar and run on a sample of the benchmark inputs (bench.make_workload / make_moma_workload, seed 0).
Every +, -, *, /, sqrt, sin, cos counts 1 (an FMA therefore 2); operations on exact structural zeros (the branch-free kernels
multiply absent rows by 0) and multiplications by +-1 are not counted.  Phases: kinematics, dynamics, manipulability, QP build,
self-collision; ADMM load / equilibration / factorisation / iteration / termination check / emit.

    python tools/count_flops.py [--sample 512]
"""
import argparse
import ctypes as C
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import LINK, MOMA_DESC, WORKLOADS, make_moma_workload, make_workload, robot_paths  # noqa: E402

PH = ["other", "kin", "dyn", "mani", "build", "collision", "qp_load", "qp_scale", "qp_factor", "qp_iter", "qp_check", "qp_emit"]
_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int)


def lib():
    d = ROOT / "tools" / "flopcount"
    so = d / "libflopcount.so"
    srcs = [d / "flopcount.cpp", ROOT / "dyros_robot_controller_b200" / "csrc" / "model.cpp"]
    newest = max(p.stat().st_mtime for p in list((ROOT / "dyros_robot_controller_b200" / "csrc").glob("*.h")) + srcs)
    if not so.exists() or so.stat().st_mtime < newest:
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", str(so), *map(str, srcs)], check=True)
    L = C.CDLL(str(so))
    L.fc_create.restype = C.c_void_p
    L.fc_rho_updates.restype = C.c_longlong
    return L


def p(a, t=_D):
    return None if a is None else a.ctypes.data_as(t)


def count(L, name, B):
    wl = WORKLOADS[name]
    urdf, srdf = robot_paths(wl["robot"])
    h = C.c_void_p(L.fc_create(Path(urdf).read_text().encode(), Path(srdf).read_text().encode()))
    moma = wl["robot"] not in ("fr3", "ur5e")
    link = wl.get("link", LINK)
    import dyros_robot_controller_b200.engine as eng   # model limits without a GPU: the host model compiler only
    model = eng.Model(urdf, srdf)
    if moma:
        md = MOMA_DESC[wl["robot"]]
        kin = md["kin"]
        t = dict(Differential=0, Mecanum=1, Caster=2)[kin["type"]]
        pos = np.asarray(kin.get("base2wheel_positions", np.zeros((0, 2))), np.float64).reshape(-1, 2)
        w = md["w"]
        arr = lambda a: np.ascontiguousarray(np.asarray(a, np.float64))
        ra, ba = arr(kin.get("roller_angles", np.zeros(w))), arr(kin.get("base2wheel_angles", np.zeros(w)))
        bx, by = (arr(pos[:, 0]), arr(pos[:, 1])) if len(pos) else (np.zeros(w), np.zeros(w))
        assert L.fc_attach_base(h, t, C.c_double(kin.get("wheel_radius", 0.0)), C.c_double(kin.get("base_width", 0.0)), C.c_double(kin.get("wheel_offset", 0.0)),
                                w, p(ra), p(bx), p(by), p(ba), md["joint_idx"]["virtual_start"], md["joint_idx"]["mani_start"], md["joint_idx"]["mobi_start"],
                                md["actuator_idx"]["mani_start"], md["actuator_idx"]["mobi_start"]) == 0
        q, qd, q_t, xd = make_moma_workload(model.q_lower, model.q_upper, model.v_limit, w, B, seed=0)
        # base twist consistent with the wheels is not needed for a flop count
    else:
        q, qd, q_t, xd = make_workload(model, B, seed=0)
    f = L.fc_frame_id(h, link.encode())
    x_t = np.zeros((B, 12))
    assert L.fc_pose(h, f, B, p(np.ascontiguousarray(q_t)), p(x_t)) == 0
    kind = dict(ik=0, id=1, taskspace=2)[wl["kind"]]
    iters, status, fl = np.zeros(B, np.int32), np.zeros(B, np.int32), np.zeros(len(PH), np.int64)
    assert L.fc_run(h, kind, f, B, p(np.ascontiguousarray(q)), p(np.ascontiguousarray(qd)), p(x_t), p(np.ascontiguousarray(xd)), p(iters, _I), p(status, _I),
                    fl.ctypes.data_as(C.POINTER(C.c_longlong))) == 0
    rho_updates = int(L.fc_rho_updates())
    L.fc_destroy(h)
    per = dict(zip(PH, (fl / B).tolist()))
    front = {k: per[k] for k in ("kin", "dyn", "mani", "build", "collision")}
    front["total"] = sum(front.values())
    out = dict(front=front, sample=dict(robots=B, seed=0), source="tools/count_flops.py: product kernel bodies with a flop-counting scalar")
    if kind != 2:
        n_it, n_chk, n_fac = float(iters.sum()), float(np.ceil(iters / 25.0).sum()), float(B + rho_updates)
        out["admm"] = dict(iter=float(fl[PH.index("qp_iter")]) / n_it, check=float(fl[PH.index("qp_check")]) / n_chk,
                           factor=float(fl[PH.index("qp_factor")]) / n_fac,
                           scale=(per["qp_load"] + per["qp_scale"] + per["qp_emit"]), refactor_rate=rho_updates / n_it)
        out["sample"].update(mean_iters=float(iters.mean()), solved=float((status == 1).mean()), mean_rho_updates=rho_updates / B,
                             admm_flops_per_robot=float(sum(fl[PH.index(k)] for k in ("qp_load", "qp_scale", "qp_factor", "qp_iter", "qp_check", "qp_emit"))) / B)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sample", type=int, default=512)
    a = ap.parse_args()
    L = lib()
    res = {}
    for name in WORKLOADS:
        res[name] = count(L, name, a.sample)
        ad = res[name].get("admm")
        print(name, {k: round(v, 1) for k, v in res[name]["front"].items()}, {k: round(v, 3) for k, v in ad.items()} if ad else "", res[name]["sample"])
    out = ROOT / "bench" / "flops_model.json"
    out.write_text(json.dumps(dict(note="fp64 flops per robot; FMA = 2; structural zeros and +-1 factors not counted (tools/count_flops.py)", workloads=res), indent=1) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
