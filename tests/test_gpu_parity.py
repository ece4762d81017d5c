"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs.  Tolerances: dynamics 1e-9 relative (north_star), QP command 1e-4 (north_star),
identical OSQP status / iteration counts (same algorithm, same schedule)."""
import numpy as np
import pytest

from tests.conftest import LINK, workload

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def test_native_library_loaded(gpu_ctx):
    from dyros_robot_controller_b200 import _capi
    assert _capi.lib_path().exists()
    assert _capi.lib().drc_device_count() >= 1


@pytest.mark.parametrize("B,seed,stress", [(1, 0, False), (257, 1, False), (4096, 2, True)])
def test_update_state_and_getters(gpu_ctx, oracle, B, seed, stress):
    model, ctx = gpu_ctx
    q, qd, _, _ = workload(oracle.model, B, seed, stress)
    f = oracle.frame_id(LINK)
    ref = oracle.update_state(q, qd, f)
    ctx.update_state(q, qd)
    fr = ctx.get_frame(LINK)
    dy = ctx.get_dynamics()
    assert rel(fr["pose"], ref["pose"]) < 1e-12
    assert rel(fr["J"], ref["J"]) < 1e-12
    assert rel(fr["Jdot"], ref["Jdot"]) < 1e-11
    assert rel(fr["vel"], np.einsum("bij,bj->bi", ref["J"], qd)) < 1e-11
    assert rel(dy["M"], ref["M"]) < 1e-9
    assert rel(dy["g"], ref["g"]) < 1e-9
    assert rel(dy["nle"], ref["nle"]) < 1e-9
    assert rel(dy["c"], ref["nle"] - ref["g"]) < 1e-8
    assert rel(dy["Minv"], ref["Minv"]) < 1e-8  # cond(M) ~ 2e4
    # M symmetric positive definite, M Minv = I
    assert np.abs(dy["M"] - np.swapaxes(dy["M"], 1, 2)).max() == 0.0
    assert np.abs(np.einsum("bij,bjk->bik", dy["M"], dy["Minv"]) - np.eye(7)).max() < 1e-8


def test_manipulability(gpu_ctx, oracle):
    model, ctx = gpu_ctx
    q, qd, _, _ = workload(oracle.model, 2048, 3, stress=True)
    f = oracle.frame_id(LINK)
    m_ref, g_ref, gd_ref = oracle.manipulability(q, qd, f, with_graddot=True)
    ctx.update_state(q, qd)
    m, g, gd = ctx.get_manipulability(LINK, with_graddot=True)
    assert np.abs(m - m_ref).max() < 1e-11
    assert np.abs(g - g_ref).max() < 1e-9 * max(1.0, np.abs(g_ref).max())
    assert np.abs(gd - gd_ref).max() < 1e-8 * max(1.0, np.abs(gd_ref).max())


def test_min_distance(gpu_ctx, oracle):
    model, ctx = gpu_ctx
    q, qd, _, _ = workload(oracle.model, 4096, 4)
    ref = oracle.min_distance(q, qd, with_graddot=True)
    ctx.update_state(q, qd)
    d, g, gd, pair = ctx.get_min_distance(with_graddot=True)
    assert (pair == ref["pair"]).mean() > 0.999
    same = pair == ref["pair"]
    assert np.abs(d - ref["d"])[same].max() < 1e-8          # GJK pairs stop at a 1e-10 duality gap
    assert np.abs(g - ref["grad"])[same].max() < 1e-4        # witness points of curved shapes ~ sqrt(gap)
    assert np.abs(gd - ref["grad_dot"])[same].max() < 1e-4
    # penetrating configurations exist in the sample and agree in sign
    assert ((d < 0) == (ref["d"] < 0)).all()


@pytest.mark.parametrize("mode,B,seed,stress", [(1, 1, 0, False), (1, 1000, 5, False), (1, 8192, 6, True), (3, 2048, 7, False)])
def test_control_cycle_matches_oracle(gpu_ctx, oracle, mode, B, seed, stress):
    """updateState + QPIKStep (mode 1) / QPIDStep (mode 3), fused device path via the host C-ABI."""
    model, ctx = gpu_ctx
    q, qd, q_t, xdot_t = workload(oracle.model, B, seed, stress)
    f = oracle.frame_id(LINK)
    x_t = oracle.update_state(q_t, qd, f)["pose"]
    ref = oracle.cycle(mode, q, qd, x_t, xdot_t, f)
    r = ctx.cycle_qpik_step(q, qd, x_t, xdot_t, LINK) if mode == 1 else ctx.cycle_qpid_step(q, qd, x_t, xdot_t, LINK)
    assert (r["status"] == ref["status"]).mean() > 0.999
    same = (r["iters"] == ref["iters"]) & (r["status"] == ref["status"])
    assert same.mean() > 0.99, f"{(~same).sum()} of {B} robots took a different ADMM path"
    scale = np.abs(ref["out"]).max()
    tol = 1e-4 if mode == 1 else 1e-4 * max(1.0, scale)   # qdot [rad/s] | torque [Nm], QPID KKT cond ~ 1e8
    err = np.abs(r["out"] - ref["out"]).max(axis=1)
    # 1e-4 for (nearly) every robot; the exceptions are robots whose ACTIVE self-collision row comes from a
    # cylinder/box pair: GJK witness points of curved shapes are only good to ~sqrt(gap * radius) ~ 3e-6 on either
    # side, and an active row amplifies that ("GJK witness precision")
    assert (err[same] < tol).mean() > 0.995
    assert err[same].max() < 100 * tol
    # robots whose iteration count differs still agree within OSQP's own tolerance band
    if (~same).any():
        assert np.abs(r["out"] - ref["out"])[~same].max() < 5e-2 * max(1.0, scale)


def test_unfused_equals_fused(gpu_ctx, oracle):
    model, ctx = gpu_ctx
    q, qd, q_t, xdot_t = workload(oracle.model, 777, 8)
    f = oracle.frame_id(LINK)
    x_t = oracle.update_state(q_t, qd, f)["pose"]
    a = ctx.cycle_qpik_step(q, qd, x_t, xdot_t, LINK)
    ctx.update_state(q, qd)
    b = ctx.qpik_step(x_t, xdot_t, LINK)
    assert (a["status"] == b["status"]).all() and (a["iters"] == b["iters"]).all()
    assert np.abs(a["out"] - b["out"]).max() < 1e-12
    # QPIK with an explicit desired task velocity
    des = 0.2 * np.random.default_rng(0).normal(size=(777, 6))
    c = ctx.qpik(des, LINK)
    ref = oracle.cycle(0, q, qd, None, des, f)
    same = c["iters"] == ref["iters"]
    assert same.mean() > 0.99 and np.abs(c["out"] - ref["out"])[same].max() < 1e-4


def test_fused_cycle_leaves_the_full_state_cache(gpu_ctx, oracle):
    """updateState + QPIKStep in one call: the dynamics (M, M^-1, g, nle) are computed by a kernel that runs next to the ADMM
    solve (QPIK does not read them); after the call every getter must answer for the NEW state, batches large and small."""
    model, ctx = gpu_ctx
    f = oracle.frame_id(LINK)
    for B, seed in ((9000, 31), (300, 32)):
        q0, qd0, _, _ = workload(oracle.model, B, seed + 100)
        ctx.update_state(q0, qd0)                       # stale cache of another state
        q, qd, q_t, xdot_t = workload(oracle.model, B, seed)
        x_t = oracle.update_state(q_t, qd, f)["pose"]
        ctx.cycle_qpik_step(q, qd, x_t, xdot_t, LINK)
        ref = oracle.update_state(q, qd, f)
        dy = ctx.get_dynamics()
        rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
        assert rel(dy["M"], ref["M"]) < 1e-9 and rel(dy["g"], ref["g"]) < 1e-9 and rel(dy["nle"], ref["nle"]) < 1e-9
        assert np.abs(np.einsum("bij,bjk->bik", dy["Minv"], ref["M"]) - np.eye(7)).max() < 1e-7
        fr = ctx.get_frame(LINK)
        assert rel(fr["J"], ref["J"]) < 1e-12 and rel(fr["pose"], ref["pose"]) < 1e-12


def test_device_pointer_path(gpu_ctx, oracle):
    """torch CUDA tensors -> drc_batch_* (device pointers, async) gives the host path's result."""
    import torch
    model, ctx = gpu_ctx
    q, qd, q_t, xdot_t = workload(oracle.model, 3000, 9)
    x_t = oracle.update_state(q_t, qd, oracle.frame_id(LINK))["pose"]
    h = ctx.cycle_qpik_step(q, qd, x_t, xdot_t, LINK)
    dev = torch.device("cuda", 0)
    d = ctx.cycle_qpik_step(*(torch.from_numpy(a).to(dev) for a in (q, qd, x_t, xdot_t)), LINK)
    torch.cuda.synchronize()
    assert (d["status"].cpu().numpy() == h["status"]).all()
    assert np.abs(d["out"].cpu().numpy() - h["out"]).max() < 1e-12


def test_streams_are_ordered_between_calls(gpu_ctx, oracle):
    """torch inputs run on the caller's (torch) stream, the numpy getters on the context's own non-blocking stream: a getter
    issued right after a torch update_state must see the NEW state (cross-stream ordering of consecutive calls)."""
    import torch
    model, ctx = gpu_ctx
    dev = torch.device("cuda", 0)
    f = oracle.frame_id(LINK)
    B = 65536
    for seed in (41, 42, 43):
        q, qd, _, _ = workload(oracle.model, B, seed)
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):           # a user stream, with work queued in front of the update
            tq, tqd = torch.from_numpy(q).to(dev, non_blocking=True), torch.from_numpy(qd).to(dev, non_blocking=True)
            junk = torch.randn(4096, 4096, device=dev) @ torch.randn(4096, 4096, device=dev)
            ctx.update_state(tq, tqd)
        fr = ctx.get_frame(LINK, want=("pose", "J"))     # host path, context stream
        dy = ctx.get_dynamics(want=("M",))
        idx = np.arange(0, B, 257)
        ref = oracle.update_state(q[idx], qd[idx], f)
        assert rel(fr["pose"][idx], ref["pose"]) < 1e-12 and rel(fr["J"][idx], ref["J"]) < 1e-12 and rel(dy["M"][idx], ref["M"]) < 1e-9
        torch.cuda.synchronize()
        del junk


def test_taskspace_controllers(gpu_ctx, oracle):
    model, ctx = gpu_ctx
    B = 1500
    q, qd, q_t, xdot_t = workload(oracle.model, B, 10, stress=True)
    f = oracle.frame_id(LINK)
    x_t = oracle.update_state(q_t, qd, f)["pose"]
    rng = np.random.default_rng(1)
    null = rng.normal(size=(B, 7))
    ctx.update_state(q, qd)
    for nv in (None, null):
        a = ctx.clik_step(x_t, xdot_t, LINK, null_qdot=nv)
        b = oracle.taskspace(0, q, qd, x_t, xdot_t, f, null_vec=nv)
        assert np.abs(a - b).max() < 1e-7 * max(1.0, np.abs(b).max())
        a = ctx.osf_step(x_t, xdot_t, LINK, null_torque=nv)
        b = oracle.taskspace(1, q, qd, x_t, xdot_t, f, null_vec=nv)
        assert np.abs(a - b).max() < 1e-7 * max(1.0, np.abs(b).max())
    xdd = rng.normal(size=(B, 6))
    assert np.abs(ctx.osf(xdd, LINK) - oracle.taskspace(2, q, qd, None, xdd, f)).max() < 1e-7
    tau = ctx.joint_torque_step(q_t, 0.5 * qd)
    assert rel(tau, oracle.joint_torque_step(q, qd, q_t, 0.5 * qd)) < 1e-10


def test_task_space_cubic(gpu_ctx, oracle):
    from oracle import c_oracle
    model, ctx = gpu_ctx
    q, qd, q_t, xdot_t = workload(oracle.model, 64, 11)
    f = oracle.frame_id(LINK)
    x0 = oracle.update_state(q, qd, f)["pose"]
    x1 = oracle.update_state(q_t, qd, f)["pose"]
    v0 = 0.1 * np.random.default_rng(2).normal(size=(64, 6))
    for t in (-0.1, 0.0, 0.37, 1.0, 1.2):
        xd, xdd = ctx.task_space_cubic(x1, xdot_t, x0, v0, t, 0.0, 1.0)
        for b in range(0, 64, 7):
            rx, rv = c_oracle.task_space_cubic(c_oracle.pose44(x1[b]), xdot_t[b], c_oracle.pose44(x0[b]), v0[b], t, 0.0, 1.0)
            assert np.abs(c_oracle.pose44(xd[b]) - rx).max() < 1e-10
            assert np.abs(xdd[b] - rv).max() < 1e-10


def test_full_size_properties(gpu_ctx, oracle):
    """BASELINE batch (65536): size-independent properties of the solutions + oracle spot check."""
    model, ctx = gpu_ctx
    B = 65536
    q, qd, q_t, xdot_t = workload(oracle.model, B, 12)
    ctx.update_state(q_t, qd)
    x_t = ctx.get_frame(LINK, want=("pose",))["pose"]
    r = ctx.cycle_qpik_step(q, qd, x_t, xdot_t, LINK)
    assert (r["status"] == 1).mean() > 0.999
    assert (r["iters"] % 25 == 0).all() and r["iters"].min() >= 25
    ok = r["status"] == 1
    # joint-velocity bounds hold up to OSQP's primal tolerance (eps_abs + eps_rel*|.|, unscaled)
    viol = np.abs(r["out"][ok]) - model.v_limit
    assert viol.max() < 1e-2
    # failures follow the reference fallback: zeros (robot_controller.cpp:283-287)
    assert np.abs(r["out"][~ok]).max(initial=0.0) == 0.0
    # determinism: the same launch twice is bit-identical
    r2 = ctx.cycle_qpik_step(q, qd, x_t, xdot_t, LINK)
    assert (r2["iters"] == r["iters"]).all() and np.array_equal(r2["out"], r["out"])
    # permutation equivariance: robots are independent
    perm = np.random.default_rng(0).permutation(B)
    r3 = ctx.cycle_qpik_step(q[perm], qd[perm], x_t[perm], xdot_t[perm], LINK)
    assert np.array_equal(r3["out"], r["out"][perm])
    # oracle spot check on a strided sample
    idx = np.arange(0, B, 64)
    ref = oracle.cycle(1, q[idx], qd[idx], x_t[idx], xdot_t[idx], oracle.frame_id(LINK))
    same = ref["iters"] == r["iters"][idx]
    assert same.mean() > 0.99
    assert np.abs(ref["out"] - r["out"][idx])[same].max() < 1e-4


def test_error_paths(gpu_ctx):
    import dyros_robot_controller_b200 as drc
    from dyros_robot_controller_b200._capi import DrcError
    model, ctx = gpu_ctx
    with pytest.raises(KeyError):
        ctx._frame("no_such_link")
    small = drc.Context(model, 8, device=0)
    with pytest.raises(DrcError):
        small.update_state(np.zeros((9, 7)), np.zeros((9, 7)))   # batch larger than the context
    with pytest.raises(DrcError):
        small.set_params(max_iter=0)


def test_reference_api_mirror(oracle):
    """dyros_robot_controller_b200.drc mirrors the reference's Python package: the example's call sequence
    (examples/python/fr3_controller.py:100-176: update_state -> get_pose / get_velocity -> QPIK_cubic) on one robot
    and on a batch."""
    from dyros_robot_controller_b200.drc.manipulator import RobotController, RobotData
    from tests.conftest import SRDF, URDF
    rd = RobotData(URDF, SRDF, max_batch=64)
    rc = RobotController(0.001, rd)
    q, qd, q_t, xdot_t = workload(oracle.model, 64, 21)
    f = oracle.frame_id(LINK)
    x_t = oracle.update_state(q_t, qd, f)["pose"]
    ref = oracle.cycle(1, q, qd, x_t, xdot_t, f)
    # single robot, reference shapes
    assert rd.update_state(q[0], qd[0]) is True
    T = rd.get_pose(LINK)
    assert T.shape == (4, 4) and np.allclose(T[3], [0, 0, 0, 1])
    assert rd.get_jacobian(LINK).shape == (6, 7) and rd.get_mass_matrix().shape == (7, 7)
    assert np.abs(rd.get_velocity(LINK) - rd.get_jacobian(LINK) @ qd[0]).max() < 1e-12
    lo, hi = rd.get_joint_position_limit()
    assert lo.shape == (7,) and (hi > lo).all()
    md = rd.get_min_distance(True, True)
    mm = rd.get_manipulability(True, False, LINK)
    assert np.ndim(md.distance) == 0 and md.grad.shape == (7,) and mm.grad.shape == (7,) and np.abs(mm.grad_dot).max() == 0
    from oracle import c_oracle
    qdot = rc.QPIK_step(c_oracle.pose44(x_t[0]), xdot_t[0], LINK)
    assert qdot.shape == (7,) and np.abs(qdot - ref["out"][0]).max() < 1e-4
    # stateless twin does not disturb the cache
    M_other = rd.compute_mass_matrix(q[1])
    assert np.abs(M_other - oracle.update_state(q[1:2], qd[1:2], f)["M"][0]).max() < 1e-9
    rd._single = True
    assert np.abs(rd.get_pose(LINK) - T).max() == 0.0
    # batched sibling
    rd.update_state(q, qd)
    out = rc.QPIK_step(c_oracle.pose44(x_t), xdot_t, LINK)
    same = rc.last_iters == ref["iters"]
    assert out.shape == (64, 7) and same.mean() > 0.95 and np.abs(out - ref["out"])[same].max() < 1e-4
    tau = rc.QPID_step(c_oracle.pose44(x_t), xdot_t, LINK)
    ref3 = oracle.cycle(3, q, qd, x_t, xdot_t, f)
    same = rc.last_iters == ref3["iters"]
    assert np.abs(tau - ref3["out"])[same].max() < 1e-4 * max(1.0, np.abs(ref3["out"]).max())
    # QPIK_cubic == getTaskSpaceCubic + QPIKStep (robot_controller.cpp:303-317)
    x0 = rd.get_pose(LINK)
    out2 = rc.QPIK_cubic(c_oracle.pose44(x_t), xdot_t, x0, np.zeros((64, 6)), 0.4, 0.0, 1.0, LINK)
    x_des, xd_des = rd._ctx.task_space_cubic(c_oracle.pose44(x_t), xdot_t, x0, np.zeros((64, 6)), 0.4, 0.0, 1.0)
    ref4 = oracle.cycle(1, q, qd, x_des, xd_des, f)
    same = rc.last_iters == ref4["iters"]
    assert same.mean() > 0.95 and np.abs(out2 - ref4["out"])[same].max() < 1e-4
    with pytest.raises(RuntimeError):
        rc.set_task_gain(np.ones(5), np.ones(6))


def test_schedule_hint_changes_nothing_but_the_order(gpu_ctx, oracle):
    """The ADMM launch is ordered by the previous call's iteration counts (longest first); with the hint off, on, and on
    again with a stale hint from a different batch, every robot's result is bit-identical."""
    model, ctx = gpu_ctx
    B = 20000
    q, qd, q_t, xdot_t = workload(oracle.model, B, 13, stress=True)
    x_t = oracle.update_state(q_t, qd, oracle.frame_id(LINK))["pose"]
    ctx.set_params(schedule_hint=0)
    a = ctx.cycle_qpik_step(q, qd, x_t, xdot_t, LINK)
    ctx.set_params(schedule_hint=1)
    b = ctx.cycle_qpik_step(q, qd, x_t, xdot_t, LINK)            # hint = the counts of call a (exact)
    perm = np.random.default_rng(3).permutation(B)
    ctx.cycle_qpik_step(q[perm], qd[perm], x_t[perm], xdot_t[perm], LINK)
    c = ctx.cycle_qpik_step(q, qd, x_t, xdot_t, LINK)            # hint from the permuted batch (wrong for every robot)
    for r in (b, c):
        assert np.array_equal(r["out"], a["out"]) and np.array_equal(r["iters"], a["iters"]) and np.array_equal(r["status"], a["status"])
    assert ctx.get_params().schedule_hint == 1
