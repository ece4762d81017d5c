"""GPU path against the COMMITTED golden fixtures (tests/golden/*.npz; generators tools/make_golden.py,
tools/make_golden_mobile.py): the CUDA kernels through the C ABI must reproduce the stored vectors."""
from pathlib import Path

import numpy as np
import pytest

from tests.conftest import LINK, MOMA
from tests.test_mobile_cpu import KINS

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def test_fr3_golden_vectors(gpu_ctx):
    model, ctx = gpu_ctx
    G = np.load(GOLD / "fr3_golden.npz")
    q, qd = G["q"], G["qd"]
    ctx.update_state(q, qd)
    fr, dy = ctx.get_frame(LINK), ctx.get_dynamics()
    assert rel(fr["pose"], G["pose"]) < 1e-12 and rel(fr["J"], G["J"]) < 1e-12 and rel(fr["Jdot"], G["Jdot"]) < 1e-11
    assert rel(dy["M"], G["M"]) < 1e-9 and rel(dy["g"], G["g"]) < 1e-9 and rel(dy["nle"], G["nle"]) < 1e-9
    assert rel(dy["Minv"], G["Minv"]) < 1e-7
    m = ctx.get_manipulability(LINK, with_graddot=True)
    assert np.abs(m[0] - G["mani"]).max() < 1e-11 and np.abs(m[1] - G["mani_grad"]).max() < 1e-9
    d = ctx.get_min_distance(with_graddot=True)
    assert np.abs(d[0] - G["dist"]).max() < 1e-6      # GJK gap tolerance of curved pairs
    for name, fn in (("qpik_step", ctx.cycle_qpik_step), ("qpid_step", ctx.cycle_qpid_step)):
        r = fn(q, qd, G["x_target"], G["xdot_target"], LINK)
        same = (r["iters"] == G[name + "_iters"]) & (r["status"] == G[name + "_status"])
        assert same.mean() >= 0.9, (name, same.mean())
        scale = max(1.0, np.abs(G[name + "_out"]).max())
        assert np.abs(r["out"] - G[name + "_out"])[same].max() < 1e-4 * scale
    ctx.update_state(q, qd)
    r = ctx.qpik(G["des"], LINK)
    same = r["iters"] == G["qpik_iters"]
    assert same.mean() >= 0.9 and np.abs(r["out"] - G["qpik_out"])[same].max() < 1e-4
    assert rel(ctx.clik_step(G["x_target"], G["xdot_target"], LINK, null_qdot=G["null"]), G["clik"]) < 1e-8
    assert rel(ctx.osf_step(G["x_target"], G["xdot_target"], LINK, null_torque=G["null"]), G["osf_step"]) < 1e-7


def test_mobile_golden_vectors():
    import dyros_robot_controller_b200 as drc
    G = np.load(GOLD / "mobile_golden.npz")
    for name, kin in KINS.items():
        base = drc.MobileBase(kin, device=0)
        g = lambda k: G[f"{name}_{k}"]
        J, vel = base.fk(g("wheel_pos"), g("wheel_vel"))
        Ji, wheel = base.ik(g("wheel_pos"), g("base_vel_des"))
        _, wsat = base.ik(g("wheel_pos"), g("base_vel_des"), saturate=True)
        for got, key in ((J, "J_fk"), (vel, "base_vel"), (Ji, "J_ik"), (wheel, "wheel_cmd"), (wsat, "wheel_cmd_saturated")):
            assert np.abs(got - g(key)).max() < 1e-11 * max(1.0, np.abs(g(key)).max()), (name, key)
    d = MOMA["pcv_fr3"]
    model = drc.Model(d["urdf"], d["srdf"]).attach_mobile_base(d["kin"], d["joint_idx"], d["actuator_idx"])
    ctx = drc.Context(model, 64, device=0)
    ctx.moma_update_state(G["pcv_q"], G["pcv_qd"])
    s = ctx.moma_get_state(LINK)
    assert rel(s["M"], G["pcv_M"]) < 1e-9 and rel(s["J"], G["pcv_J"]) < 1e-12 and rel(s["Jdot"], G["pcv_Jdot"]) < 1e-11
    assert np.abs(s["mani"] - G["pcv_mani"]).max() < 1e-11
    r = ctx.moma_cycle("ik", G["pcv_q"], G["pcv_qd"], G["pcv_x_target"], G["pcv_xdot_target"], LINK)
    same = (r["iters"] == G["pcv_qpik_step_iters"]) & (r["status"] == G["pcv_qpik_step_status"])
    assert same.mean() >= 0.9
    assert np.abs(r["out"] - G["pcv_qpik_step_out"])[same].max() < 1e-4 * max(1.0, np.abs(G["pcv_qpik_step_out"]).max())
