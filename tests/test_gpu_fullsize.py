"""GPU parity at the BASELINE batch sizes (BASELINE.json configs 3-5; config 1 is tests/test_gpu_parity.py::test_full_size_properties):

  config 3  FR3 QPIDStep                       B = 65 536     (src/manipulator/QP_ID.cpp:92-193)
  config 4  Husky-FR3 whole-body QPIK + QPID   B = 262 144    (src/mobile_manipulator/QP_IK.cpp:59-128, QP_ID.cpp:75-184)
  config 5  XLS-FR3 whole-body QPIK + QPID     B = 131 072    (the per-GPU shard of the 1 M batch on 8 GPUs)

Each case: size-independent properties over the FULL batch (determinism, permutation equivariance, OSQP's check cadence,
reference fallbacks) and a strided >= 4096-robot comparison with the oracle that includes
  * the ACTIVE-SET gate of north_star: the dual vector of OSQP's final iterate is exported by the solver
    (drc_ctx_enable_qp_debug) and compared row by row with the oracle's (rows in the reference's order); a row is active
    iff its multiplier is non-zero (ADMM: y = rho (v - Proj(v)) vanishes exactly off the bound); disagreements are allowed
    only inside the stated band |y| < 1e-6 (1 + |y|_max);
  * the per-outlier justification: every robot on the same ADMM path (status, iteration count) whose command differs by more
    than 1e-4 must have an ACTIVE self-collision row whose argmin pair is a GJK-type pair (cylinder / box against
    cylinder / box; DESIGN.md "GJK witness precision") -- anything else fails the test."""
import numpy as np
import pytest

from tests.conftest import LINK, MOMA, moma_workload, struct_to_ref, workload

pytestmark = pytest.mark.gpu

Y_BAND = 1e-6      # multipliers below Y_BAND * (1 + max |y|) may be active in one implementation and not in the other
CMD_TOL = 1e-4     # north_star: commanded qdot / torque (torques: relative to the torque scale of the batch)
OUTLIER_MAX = 2e-2  # GJK-justified outliers stay below this (times the scale)


def gate(kind, r, dbg, ref, gjk_pair, col_row, scale):
    """active-set gate + outlier justification on one strided sample.  r / ref: dicts with out, status, iters; dbg: product
    x / y in reference order; ref["x"], ref["y"]: oracle; gjk_pair: argmin pair is GJK-type; col_row: index of the
    self-collision row."""
    same = (r["iters"] == ref["iters"]) & (r["status"] == ref["status"])
    assert same.mean() > 0.99, f"{kind}: {(~same).sum()} of {same.size} robots took a different ADMM path"
    xg, yg = dbg
    ymax = 1.0 + np.abs(ref["y"]).max(axis=1, keepdims=True)
    act_g, act_r = np.abs(yg) > Y_BAND * ymax, np.abs(ref["y"]) > Y_BAND * ymax
    # outside the band the sets must agree: a disagreement means one side is above the band and the other EXACTLY inactive
    # or below a tenth of the band
    hard = (act_g & (np.abs(ref["y"]) < 0.1 * Y_BAND * ymax)) | (act_r & (np.abs(yg) < 0.1 * Y_BAND * ymax))
    assert not hard[same].any(), f"{kind}: active-set disagreement outside the band on {hard[same].any(axis=1).sum()} robots"
    agree = (act_g == act_r)[same]
    assert agree.mean() > 0.9999
    err = np.abs(r["out"] - ref["out"]).max(axis=1)
    out = same & (err > CMD_TOL * scale)
    col_active = (np.abs(ref["y"][:, col_row]) > 0) | (np.abs(yg[:, col_row]) > 0)
    unjustified = out & ~(col_active & gjk_pair)
    assert not unjustified.any(), (f"{kind}: {unjustified.sum()} robots differ by more than {CMD_TOL} without an active GJK-type "
                                   f"self-collision row (errors {err[unjustified][:5]})")
    assert out.mean() < 5e-3 and err[same].max() < OUTLIER_MAX * scale
    # primal vectors (slacks, torques) agree wherever the command does
    okx = same & ~out
    # (relative to each robot's own magnitude: whole-body QPID accelerations / torques reach 1e5 on a few ill-conditioned states)
    # and for 99.9 % of them; the hard-constraint whole-body QPID (KKT condition ~1e8, no slacks) amplifies rounding on a handful of
    # states -- those still agree to 2 x CMD_TOL of the batch's scale)
    relx = np.abs(xg - ref["x"]).max(axis=1) / (1.0 + np.abs(ref["x"]).max(axis=1))
    assert np.quantile(relx[okx], 0.999) < 10 * CMD_TOL
    assert np.abs(xg - ref["x"])[okx].max() < 2 * CMD_TOL * max(1.0, np.abs(ref["x"][okx]).max())
    if (~same).any():   # different iteration count: still inside OSQP's own tolerance band
        assert err[~same].max() < 5e-2 * scale
    return dict(same=float(same.mean()), outliers=int(out.sum()), agree=float(agree.mean()))


def gjk_type(o, q, qd):
    md = o.min_distance(q, qd, with_graddot=False)
    pairs, gt = np.asarray(o.model.pairs), o.model.geom_type
    pa = pairs[md["pair"]]
    return (gt[pa[:, 0]] != 0) & (gt[pa[:, 1]] != 0)


def full_batch_properties(run, B, status_ok_min, rng_seed=0):
    r = run(None)
    assert (r["status"] == 1).mean() > status_ok_min
    assert (r["iters"] % 25 == 0).all() and r["iters"].min() >= 25
    r2 = run(None)
    assert np.array_equal(r2["out"], r["out"]) and np.array_equal(r2["iters"], r["iters"])          # determinism
    perm = np.random.default_rng(rng_seed).permutation(B)
    r3 = run(perm)
    assert np.array_equal(r3["out"], r["out"][perm]) and np.array_equal(r3["iters"], r["iters"][perm])  # robots are independent
    return r


def test_config1_fr3_qpik_65536_gate(gpu_ctx, oracle):
    """the headline configuration: active-set gate and outlier justification on 4096 robots of the 65 536 batch (the full-batch
    properties of this configuration are in tests/test_gpu_parity.py::test_full_size_properties)."""
    model, ctx = gpu_ctx
    B = 65536
    q, qd, q_t, xd = workload(oracle.model, B, 12)
    f = oracle.frame_id(LINK)
    ctx.update_state(q_t, qd)
    x_t = ctx.get_frame(LINK, want=("pose",))["pose"]
    ctx.enable_qp_debug(True)
    try:
        r = ctx.cycle_qpik_step(q, qd, x_t, xd, LINK)
        dbg = ctx.qp_debug("ik", B)
    finally:
        ctx.enable_qp_debug(False)
    idx = np.arange(0, B, 16)
    ref = oracle.cycle(1, q[idx], qd[idx], x_t[idx], xd[idx], f, want_x=True, want_y=True)
    xs, ys = struct_to_ref("ik", {k: (v[idx] if isinstance(v, np.ndarray) else v) for k, v in dbg.items()})
    res = gate("fr3 qpik", {k: v[idx] for k, v in r.items()}, (xs, ys), ref, gjk_type(oracle, q[idx], qd[idx]), 3 * 7 + 2 + 2 * 7 + 1, 1.0)
    print("config 1:", res)


def test_config3_fr3_qpid_65536(gpu_ctx, oracle):
    model, ctx = gpu_ctx
    B = 65536
    q, qd, q_t, xd = workload(oracle.model, B, 112)
    f = oracle.frame_id(LINK)
    ctx.update_state(q_t, qd)
    x_t = ctx.get_frame(LINK, want=("pose",))["pose"]

    def run(perm):
        a = (q, qd, x_t, xd) if perm is None else (q[perm], qd[perm], x_t[perm], xd[perm])
        return ctx.cycle_qpid_step(*a, LINK)

    ctx.enable_qp_debug(True)
    try:
        full_batch_properties(run, B, 0.995)
        r = run(None)                      # the debug vectors belong to the LAST solve
        dbg = ctx.qp_debug("id", B)
    finally:
        ctx.enable_qp_debug(False)
    # failures follow the reference fallback: tau = g (robot_controller.cpp:326-330)
    bad = r["status"] != 1
    if bad.any():
        g = oracle.update_state(q[bad], qd[bad], f)["g"]
        assert np.abs(r["out"][bad] - g).max() < 1e-9 * max(1.0, np.abs(g).max())
    idx = np.arange(0, B, 16)                                   # 4096 robots
    ref = oracle.cycle(3, q[idx], qd[idx], x_t[idx], xd[idx], f, want_x=True, want_y=True)
    xs, ys = struct_to_ref("id", {k: (v[idx] if isinstance(v, np.ndarray) else v) for k, v in dbg.items()})
    sub = {k: v[idx] for k, v in r.items()}
    n = 7
    res = gate("fr3 qpid", sub, (xs, ys), ref, gjk_type(oracle, q[idx], qd[idx]), 6 * n + 2 + 4 * n + 1, max(1.0, np.abs(ref["out"]).max()))
    print("config 3:", res)


@pytest.mark.parametrize("name,B", [("husky_fr3", 262144), ("xls_fr3", 131072)])
def test_config45_whole_body(name, B):
    import dyros_robot_controller_b200 as drc
    from oracle.c_oracle import MomaOracle
    d = MOMA[name]
    o = MomaOracle(d["urdf"], d["srdf"], d["kin"], d["joint_idx"], d["actuator_idx"], threads=16)
    model = drc.Model(d["urdf"], d["srdf"]).attach_mobile_base(d["kin"], d["joint_idx"], d["actuator_idx"])
    ctx = drc.Context(model, B, device=0)
    f, w, act, m = o.frame_id(LINK), o.w, o.act, 7
    am = d["actuator_idx"]["mani_start"]
    q, qd, q_t, xd = moma_workload(o.model, w, B, 70)
    # base twist consistent with the wheel speeds (the reference derives the virtual joint velocities from the wheels)
    J, bv = o.mobile_state(q[:, 3:3 + w], qd[:, 3:3 + w])
    c, s = np.cos(q[:, 2]), np.sin(q[:, 2])
    qd[:, 0], qd[:, 1], qd[:, 2] = c * bv[:, 0] - s * bv[:, 1], s * bv[:, 0] + c * bv[:, 1], bv[:, 2]
    ctx.moma_update_state(q_t, qd)
    x_t = ctx.moma_get_state(LINK, want=("pose",))["pose"]
    idx = np.arange(0, B, B // 4096)
    gjk = gjk_type(o, q[idx], qd[idx])
    ctx.enable_qp_debug(True)
    for kind, mode, col_row in (("ik", 1, act + 2 * m + 1), ("id", 3, 4 * m + 1)):
        def run(perm):
            a = (q, qd, x_t, xd) if perm is None else (q[perm], qd[perm], x_t[perm], xd[perm])
            return ctx.moma_cycle(kind, *a, LINK)
        full_batch_properties(run, B, 0.9)
        r = run(None)                      # the debug vectors belong to the LAST solve
        dbg = ctx.qp_debug("moma_" + kind, B)
        bad = r["status"] != 1
        if bad.any():   # fallbacks: zeros (QPIK) / zero eta_dot (QPID), mobile_manipulator/robot_controller.cpp:156-160, 208-213
            assert np.abs((r["out"] if kind == "ik" else r["etadot"])[bad]).max() == 0.0
        ref = o.moma_cycle(mode, q[idx], qd[idx], x_t[idx], xd[idx], f, want_xy=True)
        xs, ys = struct_to_ref("moma_" + kind, {k: (v[idx] if isinstance(v, np.ndarray) else v) for k, v in dbg.items()}, n_mani=m, am=am)
        sub = dict(out=r["out"][idx], status=r["status"][idx], iters=r["iters"][idx])
        res = gate(f"{name} {kind}", sub, (xs, ys), ref, gjk, col_row, max(1.0, np.abs(ref["out"]).max()))
        print(f"{name} whole-body QP{kind.upper()} at {B}:", res)
    ctx.enable_qp_debug(False)
