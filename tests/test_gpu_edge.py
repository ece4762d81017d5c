"""GPU edge cases of the control-cycle path (through the C ABI): ragged batch sizes, batch-size independence of every robot's
result, non-finite inputs, empty batches.  The reference runs one robot per process (examples/C++/src/fr3_controller.cpp:116-134),
so "a robot's result is a function of that robot's inputs only" is the property a batched drop-in has to keep bit for bit."""
import numpy as np
import pytest

from tests.conftest import LINK, workload

pytestmark = pytest.mark.gpu


def _inputs(ctx, oracle, B, seed):
    q, qd, q_t, xdot_t = workload(oracle.model, B, seed, stress=True)
    ctx.update_state(q_t, qd)
    x_t = ctx.get_frame(LINK, want=("pose",))["pose"]
    return q, qd, x_t, xdot_t


@pytest.mark.parametrize("kind", ["ik", "id"])
def test_ragged_batches_equal_the_rows_of_a_large_one(gpu_ctx, oracle, kind):
    """Prefixes of 1 ... 8191 robots (partial warps, partial blocks, below / above the priority-pipeline batch threshold, with and
    without a schedule from an earlier call) return exactly the rows the 8192-robot call returns."""
    model, ctx = gpu_ctx
    B = 8192
    q, qd, x_t, xdot_t = _inputs(ctx, oracle, B, 21)
    call = ctx.cycle_qpik_step if kind == "ik" else ctx.cycle_qpid_step
    full = call(q, qd, x_t, xdot_t, LINK)
    full = {k: np.array(full[k], copy=True) for k in ("out", "status", "iters")}
    for n in (1, 2, 3, 4, 31, 33, 95, 129, 1000, 8191):
        r = call(q[:n], qd[:n], x_t[:n], xdot_t[:n], LINK)
        assert np.array_equal(r["status"], full["status"][:n]), (kind, n)
        assert np.array_equal(r["iters"], full["iters"][:n]), (kind, n)
        assert np.array_equal(r["out"], full["out"][:n]), (kind, n)
    # a strided subset, in another order
    idx = np.arange(B - 1, 0, -7)
    r = call(q[idx], qd[idx], x_t[idx], xdot_t[idx], LINK)
    assert np.array_equal(r["iters"], full["iters"][idx]) and np.array_equal(r["out"], full["out"][idx])


def test_nonfinite_inputs_stay_in_their_robot(gpu_ctx, oracle):
    """NaN / Inf joint states or targets: the call returns, the affected robots report failure with the reference's fallback (zeros,
    robot_controller.cpp:283-287 -- OSQP never reports `solved` on non-finite data), every other robot's result is bit-identical."""
    model, ctx = gpu_ctx
    B = 4096
    q, qd, x_t, xdot_t = _inputs(ctx, oracle, B, 22)
    clean = ctx.cycle_qpik_step(q, qd, x_t, xdot_t, LINK)
    clean = {k: np.array(clean[k], copy=True) for k in ("out", "status", "iters")}
    q2, qd2, x2, xd2 = q.copy(), qd.copy(), x_t.copy(), xdot_t.copy()
    bad = np.array([0, 1, 32, 100, 1025, 2047, 4095])
    q2[0, 3] = np.nan
    q2[1, :] = np.inf
    qd2[32, 0] = np.nan
    x2[100, 0] = np.nan
    xd2[1025, 5] = -np.inf
    q2[2047, 6] = np.nan
    x2[4095, 11] = np.nan
    r = ctx.cycle_qpik_step(q2, qd2, x2, xd2, LINK)
    good = np.ones(B, bool)
    good[bad] = False
    assert np.array_equal(r["out"][good], clean["out"][good])
    assert np.array_equal(r["iters"][good], clean["iters"][good]) and np.array_equal(r["status"][good], clean["status"][good])
    assert (r["status"][bad] != 1).all(), r["status"][bad]
    assert np.array_equal(r["out"][bad], np.zeros((bad.size, 7))), r["out"][bad]
    # and the context is healthy afterwards
    again = ctx.cycle_qpik_step(q, qd, x_t, xdot_t, LINK)
    assert np.array_equal(again["out"], clean["out"]) and np.array_equal(again["iters"], clean["iters"])


def test_empty_and_oversized_batches_are_rejected(gpu_ctx):
    from dyros_robot_controller_b200._capi import DrcError
    model, ctx = gpu_ctx
    z7, z12, z6 = np.zeros((0, 7)), np.zeros((0, 12)), np.zeros((0, 6))
    with pytest.raises(DrcError):
        ctx.update_state(z7, z7)
    with pytest.raises(DrcError):
        ctx.cycle_qpik_step(z7, z7, z12, z6, LINK)
    big = np.zeros((65537, 7))
    with pytest.raises(DrcError):
        ctx.update_state(big, big)


@pytest.mark.parametrize("robot", ["husky_fr3", "xls_fr3"])
@pytest.mark.parametrize("kind", ["ik", "id"])
def test_whole_body_ragged_batches_and_nonfinite_inputs(robot, kind):
    """The same two properties on the mobile-manipulator path (two robots per warp / one per warp, EPA pass on a side stream,
    dynamics job behind the solver): prefixes return the rows of the large call bit for bit; a NaN state fails alone."""
    import dyros_robot_controller_b200 as drc
    from tests.conftest import MOMA, moma_workload
    d = MOMA[robot]
    model = drc.Model(d["urdf"], d["srdf"]).attach_mobile_base(d["kin"], d["joint_idx"], d["actuator_idx"])
    ctx = drc.Context(model, 4096, device=0)

    class _M:   # moma_workload reads the joint limits off an oracle-like object
        q_lo, q_hi, v_lim = model.q_lower, model.q_upper, model.v_limit
    B = 4096
    q, qd, q_t, xdot_t = moma_workload(_M, model.wheel_num, B, 41)
    ctx.moma_update_state(q_t, qd)
    x_t = ctx.moma_get_state(LINK, want=("pose",))["pose"]
    full = ctx.moma_cycle(kind, q, qd, x_t, xdot_t, LINK)
    full = {k: np.array(full[k], copy=True) for k in ("out", "status", "iters")}
    for n in (1, 2, 3, 33, 129, 1000, 4095):
        r = ctx.moma_cycle(kind, q[:n], qd[:n], x_t[:n], xdot_t[:n], LINK)
        assert np.array_equal(r["status"], full["status"][:n]), (robot, kind, n)
        assert np.array_equal(r["iters"], full["iters"][:n]), (robot, kind, n)
        assert np.array_equal(r["out"], full["out"][:n]), (robot, kind, n)
    q2 = q.copy()
    bad = np.array([0, 63, 2048, 4095])
    q2[bad, 3 + model.wheel_num + 2] = np.nan   # an arm joint (a mecanum wheel ANGLE enters neither the kinematics nor the QPIK record)
    r = ctx.moma_cycle(kind, q2, qd, x_t, xdot_t, LINK)
    good = np.ones(B, bool)
    good[bad] = False
    assert np.array_equal(r["out"][good], full["out"][good]) and np.array_equal(r["iters"][good], full["iters"][good])
    assert (r["status"][bad] != 1).all()
    assert np.isfinite(r["out"][bad]).all() if kind == "ik" else True   # QPIK failure = zeros; QPID failure = gravity of a NaN state


def test_page_locked_output_buffers_take_the_direct_route(gpu_ctx, oracle):
    """Host entry points with page-locked result buffers: the solver launches store straight into them (device alias of the host
    allocation) instead of a staged D2H copy; same bits as with pageable buffers, including the failure fallback rows."""
    import torch
    model, ctx = gpu_ctx
    B = 8192
    q, qd, x_t, xdot_t = _inputs(ctx, oracle, B, 23)
    for call in (ctx.cycle_qpik_step, ctx.cycle_qpid_step):
        ref = call(q, qd, x_t, xdot_t, LINK)
        ref = {k: np.array(ref[k], copy=True) for k in ("out", "status", "iters")}
        p_out = torch.full((B, 7), float("nan"), dtype=torch.float64).pin_memory()
        p_st = torch.full((B,), -7, dtype=torch.int32).pin_memory()
        p_it = torch.full((B,), -7, dtype=torch.int32).pin_memory()
        r = call(q, qd, x_t, xdot_t, LINK, out=p_out.numpy(), status=p_st.numpy(), iters=p_it.numpy())
        assert np.array_equal(p_out.numpy(), ref["out"]) and np.array_equal(p_st.numpy(), ref["status"])
        assert np.array_equal(p_it.numpy(), ref["iters"])
        assert r["out"] is not None


@pytest.mark.parametrize("kind", ["ik", "id"])
def test_whole_body_page_locked_outputs(kind):
    """Same as above on the mobile-manipulator host entry points (targets uploaded behind stage 1, results stored directly)."""
    import torch
    import dyros_robot_controller_b200 as drc
    from tests.conftest import MOMA, moma_workload
    d = MOMA["husky_fr3"]
    model = drc.Model(d["urdf"], d["srdf"]).attach_mobile_base(d["kin"], d["joint_idx"], d["actuator_idx"])
    ctx = drc.Context(model, 4096, device=0)

    class _M:
        q_lo, q_hi, v_lim = model.q_lower, model.q_upper, model.v_limit
    B, a = 4096, model.actuated_dof
    q, qd, q_t, xdot_t = moma_workload(_M, model.wheel_num, B, 43)
    ctx.moma_update_state(q_t, qd)
    x_t = ctx.moma_get_state(LINK, want=("pose",))["pose"]
    ref = ctx.moma_cycle(kind, q, qd, x_t, xdot_t, LINK)
    ref = {k: (None if ref.get(k) is None else np.array(ref[k], copy=True)) for k in ("out", "etadot", "status", "iters")}
    p_out, p_out2 = (torch.full((B, a), float("nan"), dtype=torch.float64).pin_memory() for _ in range(2))
    p_st, p_it = (torch.full((B,), -7, dtype=torch.int32).pin_memory() for _ in range(2))
    ctx.moma_cycle(kind, q, qd, x_t, xdot_t, LINK, out=p_out.numpy(), out2=p_out2.numpy(), status=p_st.numpy(), iters=p_it.numpy())
    assert np.array_equal(p_out.numpy(), ref["out"]) and np.array_equal(p_st.numpy(), ref["status"])
    assert np.array_equal(p_it.numpy(), ref["iters"])
    if kind == "id":
        assert np.array_equal(p_out2.numpy(), ref["etadot"])
