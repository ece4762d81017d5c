"""Independent ground truth for the two parts of the oracle that had a single restatement family:

* narrow phase (oracle/src/ogeom.h: closed forms, GJK, EPA): the distance between two convex primitives is recomputed as
  a constrained minimisation of |x_a - x_b| over the two bodies (scipy SLSQP, written from the set definitions, no support
  functions, no simplex logic), and the penetration depth of overlapping bodies as the minimum over sampled + locally
  refined directions of the support-function sum of the Minkowski difference;
* QP optimum: scipy SLSQP on the dense QP of the benchmark states and a KKT certificate (stationarity, feasibility,
  complementary slackness, dual sign) of the tight-tolerance solve; the default-tolerance OSQP iterate (what the
  reference returns, QP_base.h:146-149) is then located relative to that optimum and its active set compared.

Numbers printed by these tests are quoted in DESIGN.md section 2."""
import numpy as np
import pytest
from scipy.optimize import minimize

from tests.conftest import LINK, workload

SPHERE, CYL, BOX = 0, 1, 2


def rand_rot(rng, small=None):
    w = rng.normal(size=3)
    th = np.linalg.norm(w)
    k = w / th
    if small is not None:
        th = small
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def pose(R, p):
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, p
    return T


def rand_prm(rng, t):
    if t == SPHERE:
        return np.array([rng.uniform(0.03, 0.15), 0, 0])
    if t == CYL:
        return np.array([rng.uniform(0.03, 0.1), rng.uniform(0.03, 0.2), 0])   # radius, half length (axis = local z)
    return rng.uniform(0.03, 0.15, size=3)                                      # half extents


def inside_constraints(t, prm, T, sl):
    """inequality constraints g(x) >= 0 of 'x[sl] lies in the body' (body defined in its local frame)."""
    R, c = T[:3, :3], T[:3, 3]
    loc = lambda x: R.T @ (x[sl] - c)
    if t == SPHERE:
        return [lambda x: prm[0] ** 2 - loc(x) @ loc(x)]
    if t == CYL:
        return [lambda x: prm[0] ** 2 - loc(x)[0] ** 2 - loc(x)[1] ** 2, lambda x: prm[1] - loc(x)[2], lambda x: prm[1] + loc(x)[2]]
    return [lambda x, i=i, s=s: prm[i] - s * loc(x)[i] for i in range(3) for s in (1.0, -1.0)]


def slsqp_distance(ta, pa, Ta, tb, pb, Tb, expect=None):
    """min |x_a - x_b| s.t. x_a in A, x_b in B  (convex problem).  More starting points are tried only while the result is
    still ABOVE `expect` (SLSQP can stall on the flat faces); a result below `expect` is returned at once -- it is what the
    caller's assertion is looking for."""
    cons = [dict(type="ineq", fun=f) for f in inside_constraints(ta, pa, Ta, slice(0, 3)) + inside_constraints(tb, pb, Tb, slice(3, 6))]
    best = np.inf
    ca, cb = Ta[:3, 3], Tb[:3, 3]
    for s in (0.0, 0.5, 1.0):
        if expect is not None and best < expect + 1e-6:
            break
        x0 = np.concatenate([ca + s * 0.3 * (cb - ca), cb + s * 0.3 * (ca - cb)])
        r = minimize(lambda x: (x[:3] - x[3:]) @ (x[:3] - x[3:]), x0, jac=lambda x: np.concatenate([2 * (x[:3] - x[3:]), -2 * (x[:3] - x[3:])]),
                     constraints=cons, method="SLSQP", options=dict(ftol=1e-16, maxiter=400))
        if min(c["fun"](r.x) for c in cons) > -1e-9:
            best = min(best, np.sqrt(max(r.fun, 0.0)))
    return best


def support(t, prm, T, d):
    """h(d) = max over the body of <x, d>."""
    R, c = T[:3, :3], T[:3, 3]
    dl = R.T @ d
    if t == SPHERE:
        return c @ d + prm[0] * np.linalg.norm(d)
    if t == CYL:
        return c @ d + prm[0] * np.hypot(dl[0], dl[1]) + prm[1] * abs(dl[2])
    return c @ d + np.abs(dl) @ prm


def sampled_penetration(ta, pa, Ta, tb, pb, Tb, rng):
    """penetration depth of overlapping convex bodies = min over unit n of h_A(n) + h_B(-n)  (distance from the origin to the
    boundary of the Minkowski difference): coarse sphere sampling, then local refinement of the best directions."""
    f = lambda n: support(ta, pa, Ta, n) + support(tb, pb, Tb, -n)
    N = rng.normal(size=(4000, 3))
    N /= np.linalg.norm(N, axis=1, keepdims=True)
    vals = np.array([f(n) for n in N])
    best = np.inf
    for i in np.argsort(vals)[:3]:
        r = minimize(lambda a: f(np.array([np.sin(a[0]) * np.cos(a[1]), np.sin(a[0]) * np.sin(a[1]), np.cos(a[0])])),
                     [np.arccos(np.clip(N[i, 2], -1, 1)), np.arctan2(N[i, 1], N[i, 0])], method="Nelder-Mead",
                     options=dict(xatol=1e-10, fatol=1e-13, maxiter=2000))
        best = min(best, r.fun, vals[i])
    return best


def gen_cases(rng, n):
    """random + near-parallel + touching configurations of the GJK-type pairs (cylinder-cylinder, cylinder-box, box-box) and a
    share of the closed-form ones (sphere-*)."""
    types = [(CYL, CYL), (CYL, BOX), (BOX, CYL), (BOX, BOX), (SPHERE, CYL), (SPHERE, BOX), (SPHERE, SPHERE)]
    for i in range(n):
        ta, tb = types[i % len(types)] if i % 4 else types[i % 4]
        pa, pb = rand_prm(rng, ta), rand_prm(rng, tb)
        Ra = rand_rot(rng)
        kind = i % 5
        if kind == 0:      # near-parallel axes / faces: rotation of B = rotation of A times a tiny tilt
            Rb = Ra @ rand_rot(rng, small=10.0 ** rng.uniform(-9, -2))
        else:
            Rb = rand_rot(rng)
        off = rng.normal(size=3)
        off /= np.linalg.norm(off)
        if kind == 1:      # touching: place B so that the two bodies are separated by ~0 along `off`
            gap = 10.0 ** rng.uniform(-9, -4)
            s = support(ta, pa, pose(Ra, np.zeros(3)), off) + support(tb, pb, pose(Rb, np.zeros(3)), -off) + gap
            cb = s * off
        else:
            cb = rng.uniform(0.0, 0.45) * off
        yield ta, pa, pose(Ra, np.zeros(3)), tb, pb, pose(Rb, cb), kind


def test_narrow_phase_against_constrained_minimisation(oracle):
    rng = np.random.default_rng(2024)
    n_sep = n_pen = n_par = n_touch = 0
    worst_sep = worst_pen = 0.0
    for ta, pa, Ta, tb, pb, Tb, kind in gen_cases(rng, 2100):
        d, wa, wb, _ = oracle.shape_distance(ta, pa, Ta, tb, pb, Tb)
        if d > 1e-7:
            ref = slsqp_distance(ta, pa, Ta, tb, pb, Tb, expect=d)
            if not np.isfinite(ref):
                continue
            # the oracle's d is a MINIMUM: the independent minimiser finds nothing smaller, and reaches it
            assert d <= ref + 1e-7, (ta, tb, kind, d, ref)
            assert abs(d - ref) < 2e-6, (ta, tb, kind, d, ref)
            worst_sep = max(worst_sep, abs(d - ref))
            # witness points lie on the two bodies and realise d
            assert abs(np.linalg.norm(wb - wa) - d) < 1e-9
            for t, prm, T, w in ((ta, pa, Ta, wa), (tb, pb, Tb, wb)):
                assert min(g(np.concatenate([w, w])) for g in inside_constraints(t, prm, T, slice(0, 3))) > -1e-7
            n_sep += 1
            n_par += kind == 0
            n_touch += kind == 1
        elif d < -1e-6 and n_pen < 160:
            ref = sampled_penetration(ta, pa, Ta, tb, pb, Tb, rng)
            # EPA (tolerance 1e-6, hpp-fcl's default) returns the depth of the closest face of the expanded polytope; the sampled
            # minimum is an upper bound of the true depth that the local refinement drives to ~1e-6
            assert -d <= ref + 1e-5, (ta, tb, d, ref)
            assert abs(-d - ref) < 2e-4 * max(1.0, ref / 0.05), (ta, tb, d, ref)
            worst_pen = max(worst_pen, abs(-d - ref))
            n_pen += 1
    print(f"narrow phase truth: {n_sep} separated cases (max |d - slsqp| = {worst_sep:.2e}; {n_par} near-parallel, {n_touch} touching), "
          f"{n_pen} penetrating cases (max |depth - sampled| = {worst_pen:.2e})")
    assert n_sep >= 1200 and n_pen >= 150 and n_par >= 150 and n_touch >= 150


def kkt_certificate(P, q, A, l, u, x, y, tol):
    """necessary and sufficient optimality conditions of the convex QP  min 1/2 x'Px + q'x, l <= Ax <= u."""
    Ax = A @ x
    stat = np.abs(P @ x + q + A.T @ y).max()
    feas = max((l - Ax).max(), (Ax - u).max(), 0.0)
    at_l, at_u = np.abs(Ax - l) < tol, np.abs(Ax - u) < tol
    # y_i < 0 only on rows at their lower bound, y_i > 0 only at their upper bound
    comp = max(np.abs(y[(y < 0) & ~at_l]).max(initial=0.0), np.abs(y[(y > 0) & ~at_u]).max(initial=0.0))
    return stat, feas, comp


def test_qp_optimum_against_scipy_and_kkt(oracle):
    """>= 1000 benchmark QPs (QPIK of the BASELINE state distribution, incl. the stress set): the tight solve is the optimum
    (KKT certificate; SLSQP agrees on a subset), and the default-tolerance OSQP iterate -- the reference's output -- is
    located relative to it."""
    m = oracle.model
    f = oracle.frame_id(LINK)
    B = 1000
    q, qd, q_t, xdot_t = workload(m, B, 77, stress=True)
    x_t = oracle.update_state(q_t, qd, f)["pose"]
    ref = oracle.cycle(1, q, qd, x_t, xdot_t, f, want_x=True, want_y=True)
    # the desired task velocity QPIKStep feeds the QP (robot_controller.cpp:292-300): Kp e + Kv edot, rebuilt here from the oracle's
    # own task error so that build_qp sees the same problem
    dist_x, agree, n_act, slsqp_gap = [], [], [], []
    oracle.set_qp_settings(eps_abs=1e-10, eps_rel=1e-10, max_iter=400000)
    tight = oracle.cycle(1, q, qd, x_t, xdot_t, f, want_x=True, want_y=True)
    oracle.set_qp_settings()
    assert (tight["status"] == 1).mean() > 0.995
    des = oracle.desired_task(1, q, qd, x_t, xdot_t, f)
    for b in range(B):
        if tight["status"][b] != 1 or ref["status"][b] != 1:
            continue
        P, qv, A, l, u = oracle.build_qp(0, q[b], qd[b], des[b], f)
        xs, ys = tight["x"][b], tight["y"][b]
        stat, feas, comp = kkt_certificate(P, qv, A, l, u, xs, ys, 1e-7)
        scale = 1.0 + np.abs(ys).max()
        assert stat < 1e-6 * scale and feas < 1e-7 and comp < 1e-6 * scale, (b, stat, feas, comp)
        dist_x.append(np.abs(ref["x"][b][:7] - xs[:7]).max())
        act_ref, act_opt = np.abs(ref["y"][b]) > 1e-9, np.abs(ys) > 1e-9
        agree.append((act_ref == act_opt).mean())
        n_act.append(act_opt.sum())
        if b % 10 == 0:   # SLSQP from the OSQP iterate: an independent optimiser must not find a better feasible point
            obj = lambda x: 0.5 * x @ P @ x + qv @ x
            fin_l, fin_u = l > -1e29, u < 1e29
            cons = [dict(type="ineq", fun=lambda x: (A @ x - l)[fin_l], jac=lambda x: A[fin_l]),
                    dict(type="ineq", fun=lambda x: (u - A @ x)[fin_u], jac=lambda x: -A[fin_u])]
            r = minimize(obj, ref["x"][b], jac=lambda x: P @ x + qv, constraints=cons, method="SLSQP", options=dict(ftol=1e-14, maxiter=500))
            # SLSQP ends ~1e-6 infeasible on these slack-weight-1000 problems ("positive directional derivative"): by weak duality
            # any point x satisfies obj(x) >= obj* - |y*|_1 * infeasibility(x), so it must not beat the certified optimum by more
            Axr = A @ r.x
            infeas = max((l - Axr).max(), (Axr - u).max(), 0.0)
            if infeas < 1e-4:
                slsqp_gap.append(np.abs(r.x[:7] - xs[:7]).max())
                assert obj(r.x) >= obj(xs) - np.abs(ys).sum() * infeas - 1e-6 * (1.0 + abs(obj(xs)))
    dist_x, agree = np.array(dist_x), np.array(agree)
    print(f"QP truth over {len(dist_x)} QPs: |qdot_osqp - qdot_opt| median {np.median(dist_x):.2e}, p99 {np.quantile(dist_x, 0.99):.2e}, "
          f"max {dist_x.max():.2e}; active-set agreement of the eps=1e-3 iterate with the optimum: mean {agree.mean():.4f}, "
          f"identical for {np.mean(agree == 1.0):.3f} of the QPs; mean active rows {np.mean(n_act):.1f}; SLSQP checked {len(slsqp_gap)} "
          f"(median |qdot_slsqp - qdot_opt| {np.median(slsqp_gap):.1e})")
    assert len(dist_x) >= 950 and len(slsqp_gap) >= 80
    # OSQP at eps 1e-3 is OSQP-accurate, not exact: bounded distance, mostly the same active set
    assert np.median(dist_x) < 5e-2 and agree.mean() > 0.97


def test_explicit_inverse_of_the_schur_complement_is_accurate():
    """the ADMM kernel inverts its NC x NC Schur complement explicitly (Gauss-Jordan, no pivoting) instead of
    keeping a Cholesky factor.  On the stress set, for rho over OSQP's whole clip range, the equilibrated S has a condition number
    below 20 and the inverse is accurate to rounding (tools/gj_accuracy.py; numbers in profiles/r02_gauss_jordan_accuracy.json)."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))
    from gj_accuracy import measure
    r = measure(48)
    assert r["max_condition_number"] < 20
    assert r["max_abs_SSinv_minus_I"] < 1e-13 and r["max_rel_error_vs_cholesky_solve"] < 1e-13
