"""ORACLE (test infrastructure) -- second, independent numpy restatement used to cross-check
oracle/src/*.h (two restatements agreeing is the strongest pin available: the reference has no
tests or goldens and its arithmetic lives in un-vendored Pinocchio / hpp-fcl / OSQP).

Deliberately written with DIFFERENT formulations from the C++ oracle:
  * kinematics: direct geometric formulas in the world frame (no spatial-motion transforms);
  * mass matrix: M = sum_k  Jv_k' m_k Jv_k + Jw_k' I_k Jw_k  (link-wise kinetic energy);
  * gravity: g = -sum_k m_k Jv_k' gvec;   Coriolis: Christoffel symbols by finite differences of M;
  * OSQP: the literal quasi-definite KKT system [P+sI A'; A -1/rho] solved with numpy each iteration.
Pure numpy, small cases only.
"""
from __future__ import annotations

import numpy as np

from .urdf_model import FlatModel, JOINT_REVOLUTE


def _rot(axis, th):
    a = np.asarray(axis, float)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def fk(m: FlatModel, q):
    """World placements (R_i, p_i) of every joint frame."""
    R, p = [None] * m.nv, [None] * m.nv
    for i in range(m.nv):
        Rl, pl = m.jR[i], m.jp[i]
        if m.jtype[i] == JOINT_REVOLUTE:
            Rl = Rl @ _rot(m.axis[i], q[i])
        else:
            pl = pl + m.jR[i] @ (m.axis[i] * q[i])
        if m.parent[i] < 0:
            R[i], p[i] = Rl, pl
        else:
            R[i], p[i] = R[m.parent[i]] @ Rl, R[m.parent[i]] @ pl + p[m.parent[i]]
    return R, p


def _ancestors(m, i):
    out = []
    while i >= 0:
        out.append(i)
        i = m.parent[i]
    return out


def point_jacobian(m: FlatModel, q, joint: int, point_world):
    """6 x nv LOCAL_WORLD_ALIGNED Jacobian of a point rigidly attached to `joint`."""
    R, p = fk(m, q)
    J = np.zeros((6, m.nv))
    for j in _ancestors(m, joint):
        a = R[j] @ m.axis[j]
        if m.jtype[j] == JOINT_REVOLUTE:
            J[:3, j] = np.cross(a, point_world - p[j])
            J[3:, j] = a
        else:
            J[:3, j] = a
    return J


def frame_pose(m: FlatModel, q, frame: int):
    R, p = fk(m, q)
    pj = m.frame_parent[frame]
    if pj < 0:
        return m.frame_R[frame], m.frame_p[frame]
    return R[pj] @ m.frame_R[frame], R[pj] @ m.frame_p[frame] + p[pj]


def frame_jacobian(m: FlatModel, q, frame: int):
    _, pf = frame_pose(m, q, frame)
    return point_jacobian(m, q, m.frame_parent[frame], pf)


def frame_jacobian_dot(m: FlatModel, q, qd, frame: int, h=1e-6):
    """d/dt of the LWA frame Jacobian by central differences along qd."""
    q, qd = np.asarray(q, float), np.asarray(qd, float)
    return (frame_jacobian(m, q + h * qd, frame) - frame_jacobian(m, q - h * qd, frame)) / (2 * h)


def mass_matrix(m: FlatModel, q):
    R, p = fk(m, q)
    M = np.zeros((m.nv, m.nv))
    for k in range(m.nv):
        if m.mass[k] <= 0:
            continue
        c = R[k] @ m.com[k] + p[k]
        J = point_jacobian(m, q, k, c)
        Iw = R[k] @ m.inertia[k] @ R[k].T
        M += m.mass[k] * J[:3].T @ J[:3] + J[3:].T @ Iw @ J[3:]
    return M


def gravity(m: FlatModel, q):
    R, p = fk(m, q)
    g = np.zeros(m.nv)
    for k in range(m.nv):
        c = R[k] @ m.com[k] + p[k]
        J = point_jacobian(m, q, k, c)
        g -= m.mass[k] * J[:3].T @ m.gravity
    return g


def coriolis(m: FlatModel, q, qd, h=1e-5):
    """c(q,qd) = C(q,qd) qd from Christoffel symbols of M (finite differences)."""
    n = m.nv
    q = np.asarray(q, float)
    dM = np.zeros((n, n, n))
    for k in range(n):
        e = np.zeros(n); e[k] = h
        dM[:, :, k] = (mass_matrix(m, q + e) - mass_matrix(m, q - e)) / (2 * h)
    c = np.zeros(n)
    for i in range(n):
        for j in range(n):
            for k in range(n):
                c[i] += 0.5 * (dM[i, j, k] + dM[i, k, j] - dM[j, k, i]) * qd[j] * qd[k]
    return c


def manipulability(m: FlatModel, q, frame: int, cols=None):
    J = frame_jacobian(m, q, frame)
    if cols is not None:
        J = J[:, cols]
    return np.sqrt(max(np.linalg.det(J @ J.T), 0.0))


def manipulability_grad_fd(m: FlatModel, q, frame: int, cols=None, h=1e-6):
    q = np.asarray(q, float)
    idx = range(m.nv) if cols is None else cols
    g = []
    for i in idx:
        e = np.zeros(m.nv); e[i] = h
        g.append((manipulability(m, q + e, frame, cols) - manipulability(m, q - e, frame, cols)) / (2 * h))
    return np.array(g)


# ----------------------------------------------------------------------------- OSQP (literal KKT form)
OSQP_INFTY = 1e30


def osqp_literal(P, q, A, l, u, rho=0.1, sigma=1e-6, alpha=1.6, eps_abs=1e-3, eps_rel=1e-3, max_iter=4000,
                 check=25, scaling=10, adaptive_rho_interval=50, adaptive_rho_tolerance=5.0):
    """OSQP v0.6 algorithm with the KKT system solved literally (np.linalg.solve). Returns dict."""
    P, q, A, l, u = (np.array(a, float) for a in (P, q, A, l, u))
    n, m = q.size, l.size
    D, E, c = np.ones(n), np.ones(m), 1.0

    def lim(v):
        v = np.where(v < 1e-4, 1.0, v)
        return np.minimum(v, 1e4)

    for _ in range(scaling):
        dt = np.maximum(np.abs(P).max(axis=0), np.abs(A).max(axis=0) if m else 0)
        et = np.abs(A).max(axis=1) if m else np.zeros(0)
        dt, et = 1 / np.sqrt(lim(dt)), 1 / np.sqrt(lim(et))
        P = dt[:, None] * P * dt[None, :]
        A = et[:, None] * A * dt[None, :]
        q = dt * q
        D, E = D * dt, E * et
        ct = max(np.abs(P).max(axis=0).mean(), float(lim(np.array([np.abs(q).max()]))[0]))
        ct = 1.0 / float(lim(np.array([ct]))[0])
        P, q, c = P * ct, q * ct, c * ct
    l, u = E * l, E * u
    inf = OSQP_INFTY * 1e-4
    ctype = np.where((l < -inf) & (u > inf), -1, np.where(u - l < 1e-4, 1, 0))

    def rho_vec(r):
        return np.where(ctype == -1, 1e-6, np.where(ctype == 1, 1e3 * r, r))

    rv = rho_vec(rho)

    def kkt(rv):
        return np.block([[P + sigma * np.eye(n), A.T], [A, -np.diag(1.0 / rv)]])

    K = kkt(rv)
    x, z, y = np.zeros(n), np.zeros(m), np.zeros(m)
    status, it, nupd = "max_iter", 0, 0
    for it in range(1, max_iter + 1):
        xp, zp = x, z
        sol = np.linalg.solve(K, np.concatenate([sigma * xp - q, zp - y / rv]))
        xt, nu = sol[:n], sol[n:]
        zt = zp + (nu - y) / rv
        x = alpha * xt + (1 - alpha) * xp
        zr = alpha * zt + (1 - alpha) * zp
        z = np.clip(zr + y / rv, l, u)
        y = y + rv * (zr - z)
        chk = it % check == 0
        upd = adaptive_rho_interval and it % adaptive_rho_interval == 0
        if chk or upd:
            Ax, Px, Aty = A @ x, P @ x, A.T @ y
            rp, rd = Ax - z, Px + q + Aty
            pri, dua = np.abs(rp / E).max(), np.abs(rd / D).max() / c
            ep = eps_abs + eps_rel * max(np.abs(z / E).max(), np.abs(Ax / E).max())
            ed = eps_abs + eps_rel / c * max(np.abs(q / D).max(), np.abs(Aty / D).max(), np.abs(Px / D).max())
            if chk and pri < ep and dua < ed:
                status = "solved"
                break
            if upd:
                pr = np.abs(rp).max() / (max(np.abs(z).max(), np.abs(Ax).max()) + 1e-10)
                dr = np.abs(rd).max() / (max(np.abs(q).max(), np.abs(Aty).max(), np.abs(Px).max()) + 1e-10)
                rn = float(np.clip(rho * np.sqrt(pr / (dr + 1e-10)), 1e-6, 1e6))
                if rn > rho * adaptive_rho_tolerance or rn < rho / adaptive_rho_tolerance:
                    rho, nupd = rn, nupd + 1
                    rv = rho_vec(rho)
                    K = kkt(rv)
    return dict(status=status, iters=it, x=D * x, y=E * y / c, z=z / E, rho=rho, rho_updates=nupd)
