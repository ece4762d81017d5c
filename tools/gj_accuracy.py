#!/usr/bin/env python3
"""How accurate is the explicit Gauss-Jordan inverse of the ADMM kernel's Schur complement (csrc/drc_qp.h `factor`)?

For the QPIK problems of the stress set (10 % of the states at a joint limit, 5 % near singular) the scaled problem is rebuilt in
numpy (OSQP's Ruiz passes), the NC x NC Schur complement S = K_cc - K_cd K_dd^-1 K_dc of K = P + sigma I + A' diag(rho) A is formed
for rho over OSQP's whole clip range [1e-6, 1e6], inverted with the kernel's algorithm (in-place Gauss-Jordan, no pivoting, fp64) and
compared with a Cholesky factorisation:  max |S S^-1 - I|, and the error of S^-1 u against the Cholesky solve for random u.

    python tools/gj_accuracy.py [n_states]      -> one JSON line (also asserted by tests/test_oracle_truth.py)
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def ruiz(P, q, A, scaling=10):
    P, q, A = P.copy(), q.copy(), A.copy()
    n, m = q.size, A.shape[0]
    D, E, c = np.ones(n), np.ones(m), 1.0
    lim = lambda v: np.minimum(np.where(v < 1e-4, 1.0, v), 1e4)
    for _ in range(scaling):
        dt = 1 / np.sqrt(lim(np.maximum(np.abs(P).max(axis=0), np.abs(A).max(axis=0))))
        et = 1 / np.sqrt(lim(np.abs(A).max(axis=1)))
        P = dt[:, None] * P * dt[None, :]; A = et[:, None] * A * dt[None, :]; q = dt * q
        D, E = D * dt, E * et
        ct = 1.0 / float(lim(np.array([max(np.abs(P).max(axis=0).mean(), float(lim(np.array([np.abs(q).max()]))[0]))]))[0])
        P, q, c = P * ct, q * ct, c * ct
    return P, q, A, D, E, c


def gauss_jordan_inverse(S):
    """the kernel's algorithm: in place, pivot order 0..n-1, no pivoting (S is SPD)"""
    S = S.copy()
    n = len(S)
    for k in range(n):
        piv = 1.0 / S[k, k]
        col = S[:, k] * piv
        for j in range(n):
            if j == k:
                continue
            f = col[j]
            for i in range(n):
                if i != k:
                    S[j, i] -= f * S[k, i]
            S[j, k] = -f
        S[k, :] = S[k, :] * piv
        S[k, k] = piv
    return S


def measure(n_states=256, seed=3):
    from oracle.c_oracle import Oracle
    from tests.conftest import LINK, SRDF, URDF, workload
    o = Oracle(URDF, SRDF, threads=8)
    q, qd, q_t, xd_t = workload(o.model, n_states, seed, stress=True)
    f = o.frame_id(LINK)
    x_t = o.update_state(q_t, qd, f)["pose"]
    des = o.desired_task(1, q, qd, x_t, xd_t, f)
    nc, sigma = o.nv, 1e-6
    worst = dict(ssinv=0.0, solve_rel=0.0, cond=0.0)
    rng = np.random.default_rng(0)
    for b in range(n_states):
        P, qv, A, l, u = o.build_qp(0, q[b], qd[b], des[b], f)
        Ps, qs, As, D, E, c = ruiz(P, qv, A)
        ls, us = E * l, E * u
        inf = 1e30 * 1e-4
        ctype = np.where((ls < -inf) & (us > inf), -1, np.where(us - ls < 1e-4, 1, 0))
        for rho in (1e-6, 1e-4, 1e-2, 0.1, 1.0, 1e2, 1e4, 1e6):
            rv = np.where(ctype == -1, 1e-6, np.where(ctype == 1, 1e3 * rho, rho))
            K = Ps + sigma * np.eye(len(qs)) + As.T @ (rv[:, None] * As)
            Kcc, Kcd, Kdd = K[:nc, :nc], K[:nc, nc:], K[nc:, nc:]
            S = Kcc - Kcd @ np.linalg.solve(Kdd, Kcd.T)
            Si = gauss_jordan_inverse(S)
            worst["ssinv"] = max(worst["ssinv"], float(np.abs(S @ Si - np.eye(nc)).max()))
            worst["cond"] = max(worst["cond"], float(np.linalg.cond(S)))
            L = np.linalg.cholesky(S)
            for _ in range(2):
                uvec = rng.normal(size=nc)
                ref = np.linalg.solve(L.T, np.linalg.solve(L, uvec))
                worst["solve_rel"] = max(worst["solve_rel"], float(np.abs(Si @ uvec - ref).max() / np.abs(ref).max()))
    return dict(states=n_states, rho_values=8, max_abs_SSinv_minus_I=worst["ssinv"], max_rel_error_vs_cholesky_solve=worst["solve_rel"],
                max_condition_number=worst["cond"])


if __name__ == "__main__":
    print(json.dumps(measure(int(sys.argv[1]) if len(sys.argv) > 1 else 256)))
