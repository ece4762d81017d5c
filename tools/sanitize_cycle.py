#!/usr/bin/env python3
"""Every pipeline once at small batches (incl. the priority pipeline): a quick end-to-end pass; also the driver for
compute-sanitizer where that tool is available (it is closed on the graft GPU pool).

    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_cycle.py
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import dyros_robot_controller_b200 as drc
from bench import MOMA_DESC, make_moma_workload, make_workload, robot_paths

LINK = "fr3_link8"
model = drc.Model(drc.FR3_URDF, drc.FR3_SRDF)
B = 9000                                   # >= 8192: the priority pipeline is live on the second tick
ctx = drc.Context(model, B)
q, qd, q_t, xd = make_workload(model, B, 0)
ctx.update_state(q_t, qd)
x_t = ctx.get_frame(LINK, want=("pose",))["pose"]
for k in range(3):
    r = ctx.cycle_qpik_step(q + k * 1e-3 * qd, qd, x_t, xd, LINK)
print("qpik ticks ok", int((r["status"] == 1).sum()), int(r["iters"].max()))
r = ctx.cycle_qpid_step(q[:9000], qd[:9000], x_t[:9000], xd[:9000], LINK)
r = ctx.cycle_qpid_step(q[:9000], qd[:9000], x_t[:9000], xd[:9000], LINK)
print("qpid ok", int((r["status"] == 1).sum()))
ctx.update_state(q[:500], qd[:500])
ctx.get_frame(LINK); ctx.get_dynamics(); ctx.get_manipulability(LINK, True); ctx.get_min_distance(True)
ctx.clik_step(x_t[:500], xd[:500], LINK); ctx.osf_step(x_t[:500], xd[:500], LINK); ctx.qpik(xd[:500], LINK)
print("getters ok")
urdf, srdf = robot_paths("xls_fr3")
md = MOMA_DESC["xls_fr3"]
mm = drc.Model(urdf, srdf).attach_mobile_base(md["kin"], md["joint_idx"], md["actuator_idx"])
mc = drc.Context(mm, 600)
q, qd, q_t, xd = make_moma_workload(mm.q_lower, mm.q_upper, mm.v_limit, md["w"], 600, 1)
mc.moma_update_state(q_t, qd)
x_t = mc.moma_get_state(LINK, want=("pose",))["pose"]
for kind in ("ik", "id"):
    r = mc.moma_cycle(kind, q, qd, x_t, xd, LINK)
    r = mc.moma_cycle(kind, q, qd, x_t, xd, LINK)
    print("moma", kind, "ok", int((r["status"] == 1).sum()))
# ---- mobile base alone, powered-caster whole-body cycle, closed-loop rollout
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tests.conftest import MOMA  # noqa: E402
from tests.test_mobile_cpu import KINS, wheels_of  # noqa: E402

rng = np.random.default_rng(3)
for name, kin in KINS.items():
    w = wheels_of(kin)
    base = drc.MobileBase(kin)
    for nb in (1, 127, 129, 5000):       # ragged tails of the 128-thread blocks
        J, v = base.fk(rng.uniform(-3, 3, (nb, w)), rng.uniform(-2, 2, (nb, w)))
        Ji, wv = base.ik(rng.uniform(-3, 3, (nb, w)), rng.normal(size=(nb, 3)), saturate=True)
        assert np.isfinite(J).all() and np.isfinite(v).all() and np.isfinite(Ji).all() and np.isfinite(wv).all()
print("mobile base ok")
d = MOMA["pcv_fr3"]
pm = drc.Model(d["urdf"], d["srdf"]).attach_mobile_base(d["kin"], d["joint_idx"], d["actuator_idx"])
pc = drc.Context(pm, 300)
q, qd, q_t, xd = make_moma_workload(pm.q_lower, pm.q_upper, pm.v_limit, 4, 300, 2)
pc.moma_update_state(q_t, qd)
x_t = pc.moma_get_state(LINK, want=("pose",))["pose"]
r = pc.moma_cycle("ik", q, qd, x_t, xd, LINK)
print("caster moma ok", int((r["status"] == 1).sum()))
q, qd, q_t, xd = make_workload(model, 9000, 4)
ctx.update_state(q_t, qd)
x_t = ctx.get_frame(LINK, want=("pose",))["pose"]
r = ctx.rollout_qpik(q, qd, x_t, xd, LINK, 4, 1e-3)
ctx.update_state(q, qd)
x_i = ctx.get_frame(LINK, want=("pose", "vel"))
r2 = ctx.rollout_qpik(q, qd, x_t, 0 * xd, LINK, 4, 1e-3, x_init=x_i["pose"], xdot_init=x_i["vel"], t_start=0.0, t0=0.0, duration=0.2)
print("rollout ok", int(r["fail_ticks"].sum()), int(r2["fail_ticks"].sum()), float(np.abs(r2["q"] - q).max()))
