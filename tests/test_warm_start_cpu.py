"""Warm-started rollouts (drc_params_t::rollout_warm_start; SURVEY 8f rank 1 "optional warm start") -- an EXTENSION: the reference
builds a fresh OSQP solver per cycle and never warm starts (include/dyros_robot_controller/QP_base.h:133-177).  CPU suite: the
product's solver body (host emulation, osqp_warm_start semantics on the structured KKT) against the oracle's dense OSQP port with
the same warm start, tick by tick, each side fed with ITS OWN previous solution."""
import numpy as np

from tests.conftest import LINK, workload


def _ticks(step, q, qd, x_t, xd_t, T, dt):
    q, qd = q.copy(), qd.copy()
    its, outs = [], []
    for _ in range(T):
        r = step(q, qd)
        q = q + dt * r["out"]
        qd = r["out"].copy()
        its.append(r["iters"].copy()); outs.append(r["out"].copy())
    return np.array(its), np.array(outs), q


def test_warm_start_from_zeros_is_the_cold_start(emu, oracle):
    B = 96
    q, qd, q_t, xd_t = workload(oracle.model, B, 41, stress=True)
    f = oracle.frame_id(LINK)
    x_t = oracle.update_state(q_t, qd, f)["pose"]
    cold = emu.cycle(1, q, qd, x_t, xd_t, emu.frame_id(LINK))
    wx, wy = np.zeros((B, 23)), np.zeros((B, 39))
    warm = emu.cycle_warm(1, q, qd, x_t, xd_t, emu.frame_id(LINK), wx, wy)
    assert (warm["iters"] == cold["iters"]).all() and (warm["status"] == cold["status"]).all()
    np.testing.assert_allclose(warm["out"], cold["out"], atol=1e-12)
    assert np.abs(wx).max() > 0 and np.abs(wy).max() > 0     # the solution was left for the next tick


def test_warm_started_ticks_match_the_oracle(emu, oracle):
    B, T, dt = 128, 6, 1e-3
    q, qd, q_t, xd_t = workload(oracle.model, B, 42)
    fo, fe = oracle.frame_id(LINK), emu.frame_id(LINK)
    x_t = oracle.update_state(q_t, qd, fo)["pose"]
    nx, ny = oracle.qp_sizes(0)
    ox, oy = np.zeros((B, nx)), np.zeros((B, ny))
    ex, ey = np.zeros((B, 23)), np.zeros((B, 39))
    o_it, o_out, o_q = _ticks(lambda a, b: oracle.cycle_warm(1, a, b, x_t, xd_t, fo, ox, oy), q, qd, x_t, xd_t, T, dt)
    e_it, e_out, e_q = _ticks(lambda a, b: emu.cycle_warm(1, a, b, x_t, xd_t, fe, ex, ey), q, qd, x_t, xd_t, T, dt)
    c_it, c_out, _ = _ticks(lambda a, b: oracle.cycle(1, a, b, x_t, xd_t, fo), q, qd, x_t, xd_t, T, dt)
    same = (o_it == e_it).all(0)
    assert same.mean() > 0.9, same.mean()
    # robot by robot over all ticks; an OSQP run is not a continuous function of its data (a rho-update decision can flip on a rounding
    # difference without changing the iteration count), and the warm-started closed loop carries such a flip forward: a few per cent
    close = np.abs(o_out - e_out)[:, same].max(axis=(0, 2)) < 1e-4
    assert close.mean() > 0.95, close.mean()
    assert np.abs(o_q - e_q)[same][close].max() < 1e-6
    # tick 0 is the cold start; later ticks run other iterates towards the same optimum.  (Measured: with OSQP's defaults the
    # warm start does NOT shorten this QP family's solves -- 75 iterations either way: the eps = 1e-3 iterates it starts from are
    # far from the optimum in the dual and rho restarts at 0.1 -- so the flag is an option, not a speed-up; DESIGN.md.)
    assert (o_it[0] == c_it[0]).all() and np.abs(o_out[0] - c_out[0]).max() < 1e-10
    assert np.abs(o_out[1:] - c_out[1:]).max() > 1e-6
    assert np.median(np.abs(o_out[1:] - c_out[1:]).max(-1)) < 0.05
