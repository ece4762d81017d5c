"""CPU tests of the KERNEL BODIES: tests/kernel_emu compiles the very same DRC_HD routines that nvcc
compiles into libdrc_b200.so (drc_kin.h, drc_geom.h, drc_qp.h, drc_cycle.h) with g++, lanes of a warp
emulated by a loop, and this file checks them against the oracle and the golden vectors.  The GPU
parity tests (tests/test_gpu_parity.py) then only have to establish that the device build of those
bodies behaves like the host build."""
from pathlib import Path

import numpy as np
import pytest

from tests.conftest import LINK, workload

GOLD = Path(__file__).resolve().parent / "golden" / "fr3_golden.npz"


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def test_model_compiler_matches_oracle_loader(emu, oracle):
    """C++ URDF/SRDF model compiler (csrc/model.cpp) == python loader of the oracle (two parsers, one model)."""
    a, m = emu.model_arrays(), oracle.model
    assert emu.nv == m.nv == 7
    assert (a["parent"] == m.parent).all() and (a["jtype"] == m.jtype).all()
    for k, ref in (("axis", m.axis), ("jR", m.jR), ("jp", m.jp), ("mass", m.mass), ("com", m.com), ("q_lo", m.q_lo),
                   ("q_hi", m.q_hi), ("v_lim", m.v_lim), ("geom_prm", m.geom_param), ("geom_R", m.geom_R),
                   ("geom_p", m.geom_p), ("frame_R", m.frame_R), ("frame_p", m.frame_p)):
        assert np.abs(a[k] - ref).max() < 1e-14, k
    I6 = np.stack([m.inertia[:, 0, 0], m.inertia[:, 0, 1], m.inertia[:, 0, 2], m.inertia[:, 1, 1], m.inertia[:, 1, 2],
                   m.inertia[:, 2, 2]], axis=1)
    assert np.abs(a["inertia6"] - I6).max() < 1e-14
    assert (a["geom_type"] == m.geom_type).all() and (a["geom_parent"] == m.geom_parent).all()
    assert sorted(map(tuple, a["pairs"])) == sorted(map(tuple, m.pairs))
    s = emu.sizes()
    assert s["dev_bytes"] < 32 * 1024          # the model travels as a __grid_constant__ kernel parameter (< 32 KB)


@pytest.mark.parametrize("B,seed,stress", [(1, 0, False), (300, 1, True)])
def test_state_update_bodies(emu, oracle, B, seed, stress):
    q, qd, _, _ = workload(oracle.model, B, seed, stress)
    f = oracle.frame_id(LINK)
    ref = oracle.update_state(q, qd, f)
    r = emu.update_and_get(q, qd, emu.frame_id(LINK))
    assert rel(r["pose"], ref["pose"]) < 1e-12 and rel(r["J"], ref["J"]) < 1e-12 and rel(r["Jdot"], ref["Jdot"]) < 1e-11
    assert rel(r["M"], ref["M"]) < 1e-9 and rel(r["g"], ref["g"]) < 1e-9 and rel(r["nle"], ref["nle"]) < 1e-9
    assert rel(r["Minv"], ref["Minv"]) < 1e-8
    mr, gr, gdr = oracle.manipulability(q, qd, f, with_graddot=True)
    assert np.abs(r["mani"] - mr).max() < 1e-11
    assert np.abs(r["mgrad"] - gr).max() < 1e-9 * max(1.0, np.abs(gr).max())
    assert np.abs(r["mgraddot"] - gdr).max() < 1e-8 * max(1.0, np.abs(gdr).max())


def test_self_distance_bodies(emu, oracle):
    q, qd, _, _ = workload(oracle.model, 1500, 4)
    ref = oracle.min_distance(q, qd, with_graddot=True)
    r = emu.min_distance(q, qd)
    same = r["pair"] == ref["pair"]
    assert same.mean() > 0.999
    assert np.abs(r["d"] - ref["d"])[same].max() < 1e-8
    assert np.abs(r["grad"] - ref["grad"])[same].max() < 1e-4
    assert np.abs(r["grad_dot"] - ref["grad_dot"])[same].max() < 1e-4
    assert ((r["d"] < 0) == (ref["d"] < 0)).all()


def test_narrow_phase_primitives(emu, oracle):
    """closed forms / GJK / EPA of the product against the oracle's on random primitive pairs, incl. penetration."""
    rng = np.random.default_rng(11)

    def rand_pose(scale):
        w = rng.normal(size=3); th = np.linalg.norm(w); k = w / th
        K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        T = np.eye(4); T[:3, :3] = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
        T[:3, 3] = scale * rng.normal(size=3)
        return T

    def rand_prm(t):
        if t == 0:
            return np.array([rng.uniform(0.03, 0.15), 0, 0])
        if t == 1:
            return np.array([rng.uniform(0.03, 0.1), rng.uniform(0.03, 0.2), 0])
        return rng.uniform(0.03, 0.15, size=3)

    n_pen = 0
    for _ in range(400):
        ta, tb = rng.integers(0, 3, 2)
        pa, pb = rand_prm(ta), rand_prm(tb)
        Ta, Tb = rand_pose(0.0), rand_pose(0.12)
        d0, wa0, wb0, _ = oracle.shape_distance(int(ta), pa, Ta, int(tb), pb, Tb)
        d1, wa1, wb1, _ = emu.shape_distance(int(ta), pa, Ta, int(tb), pb, Tb)
        assert abs(d0 - d1) < 1e-7, (ta, tb, d0, d1)
        n_pen += d0 < 0
        # certified lower bound used for culling never exceeds the true distance
        from tests.emu import pair_lower_bound
        assert pair_lower_bound(int(ta), pa, Ta, int(tb), pb, Tb) <= max(d0, 0.0) + 1e-9
    assert n_pen > 20


@pytest.mark.parametrize("mode,B,seed,stress", [(0, 200, 2, False), (1, 600, 5, True), (2, 100, 6, False), (3, 200, 7, True)])
def test_control_cycle_bodies(emu, oracle, mode, B, seed, stress):
    """QP record build + structured Schur-complement ADMM (drc_qp.h) == dense OSQP restatement (oracle/src/oqp.h)."""
    q, qd, q_t, xdot_t = workload(oracle.model, B, seed, stress)
    f = oracle.frame_id(LINK)
    x_t = oracle.update_state(q_t, qd, f)["pose"] if mode in (1, 3) else None
    des = xdot_t if mode in (1, 3) else 4.0 * xdot_t
    ref = oracle.cycle(mode, q, qd, x_t, des, f, want_x=True)
    r = emu.cycle(mode, q, qd, x_t, des, emu.frame_id(LINK))
    # identical status / iteration count except for borderline robots (a residual within rounding of its tolerance,
    # e.g. max_iter vs solved_inaccurate at iteration 4000)
    assert (r["status"] == ref["status"]).mean() > 0.99
    same = (r["iters"] == ref["iters"]) & (r["status"] == ref["status"])
    assert same.mean() > 0.99
    scale = max(1.0, np.abs(ref["out"]).max())
    # 1e-5 for (nearly) every robot; exceptions: robots whose ACTIVE self-collision row comes from a GJK pair, whose
    # witness points are only good to ~sqrt(gap * radius) ("GJK witness precision")
    err = np.abs(r["out"] - ref["out"]).max(axis=1)[same]
    assert (err < 1e-5 * scale).mean() >= 0.99 and err.max() < 5e-3 * scale   # slow (1000+ iteration) robots accumulate rounding
    # the full primal vector (core variables, slacks, torques) agrees too -> same active set
    n = 7
    if mode <= 1:
        x_ref = ref["x"]                                            # [qdot s_qmin s_qmax s_sing s_col]
        x_emu = np.concatenate([r["x"][:, :n], r["x"][:, n:2 * n], r["x"][:, 2 * n:3 * n], r["x"][:, 3 * n:3 * n + 2]], axis=1)
        ex = np.abs(x_emu - x_ref).max(axis=1)[same]
        assert (ex < 1e-5 * scale).mean() >= 0.99 and ex.max() < 5e-3 * scale
        # slack in use <=> its CBF row is active; same active set outside a tolerance band around the threshold
        act_ref, act_emu = x_ref[:, n:] > 1e-3, x_emu[:, n:] > 1e-3
        clear = (np.abs(x_ref[:, n:] - 1e-3) > 5e-4) & (np.abs(x_emu[:, n:] - 1e-3) > 5e-4)
        assert ((act_ref == act_emu) | ~clear)[same].all()


def test_taskspace_bodies(emu, oracle):
    B = 300
    q, qd, q_t, xdot_t = workload(oracle.model, B, 10, stress=True)
    f = oracle.frame_id(LINK)
    x_t = oracle.update_state(q_t, qd, f)["pose"]
    null = np.random.default_rng(1).normal(size=(B, 7))
    for nv in (None, null):
        for mode in (0, 1):
            a = emu.taskspace(mode, q, qd, x_t, xdot_t, emu.frame_id(LINK), aux=nv)
            b = oracle.taskspace(mode, q, qd, x_t, xdot_t, f, null_vec=nv)
            assert np.abs(a - b).max() < 1e-7 * max(1.0, np.abs(b).max())
    a = emu.taskspace(3, q, qd, None, None, emu.frame_id(LINK), aux=q_t, aux2=0.5 * qd)
    assert rel(a, oracle.joint_torque_step(q, qd, q_t, 0.5 * qd)) < 1e-10


@pytest.mark.skipif(not GOLD.exists(), reason="golden vectors not generated")
def test_kernel_bodies_reproduce_golden_vectors(emu):
    g = np.load(GOLD)
    f = emu.frame_id(LINK)
    r = emu.update_and_get(g["q"], g["qd"], f)
    for k in ("pose", "J", "Jdot", "M", "g", "nle"):
        assert rel(r[k], g[k]) < 1e-9, k
    assert np.abs(r["mani"] - g["mani"]).max() < 1e-11
    md = emu.min_distance(g["q"], g["qd"])
    assert (md["pair"] == g["dist_pair"]).all() and np.abs(md["d"] - g["dist"]).max() < 1e-8
    for mode, name in ((1, "qpik_step"), (3, "qpid_step")):
        c = emu.cycle(mode, g["q"], g["qd"], g["x_target"], g["xdot_target"], f)
        assert (c["status"] == g[name + "_status"]).all() and (c["iters"] == g[name + "_iters"]).all()
        assert np.abs(c["out"] - g[name + "_out"]).max() < 1e-5 * max(1.0, np.abs(g[name + "_out"]).max())
