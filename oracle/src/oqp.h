// ORACLE -- test infrastructure only. PARITY UNPINNED (no reference goldens exist).
// Restatement of the OSQP algorithm (Stellato et al., "OSQP: an operator splitting solver for
// quadratic programs", Math. Prog. Comp. 2020; solver v0.6.x defaults) as driven by the reference's
// QPBase::solveQP (include/dyros_robot_controller/QP_base.h:100-180): fresh solver every call,
// cold start (:146), default settings except verbose (:149), success iff status == Solved (:167).
// OSQP itself is an un-vendored, un-pinned dependency (package.xml:36) -- see SURVEY.md 8(c).
//
//   min 1/2 x'Px + q'x   s.t.  l <= Ax <= u          (dense row-major P (n x n), A (m x n))
//
// Steps restated: Ruiz equilibration of the KKT matrix with cost scaling (scaling = 10),
// rho vector by constraint type, ADMM iteration with relaxation alpha = 1.6, termination /
// infeasibility checks every 25 iterations on UNSCALED residuals, adaptive rho.
// Linear system: OSQP factorises the quasi-definite KKT matrix [P+sI A'; A -1/rho] with QDLDL;
// the oracle solves the algebraically identical reduced system (P + sI + A' diag(rho) A) x = rhs
// by dense Cholesky, which yields the same iterates up to rounding.
// Deviation: OSQP 0.6 picks the adaptive-rho interval from wall-clock time; here it is
// a fixed iteration interval (default 50, a multiple of check_termination).
#pragma once
#include "omath.h"

namespace orc {

constexpr double OSQP_INFTY = 1e30;
constexpr double OSQP_MIN_SCALING = 1e-4, OSQP_MAX_SCALING = 1e4;
constexpr double OSQP_RHO_MIN = 1e-6, OSQP_RHO_MAX = 1e6, OSQP_RHO_TOL = 1e-4, OSQP_RHO_EQ_OVER_INEQ = 1e3;

enum QpStatus { QP_SOLVED = 1, QP_MAX_ITER = 2, QP_PRIMAL_INFEASIBLE = 3, QP_DUAL_INFEASIBLE = 4, QP_NON_CVX = 5,
                QP_SOLVED_INACCURATE = 6 };

struct QpSettings {
  double rho = 0.1, sigma = 1e-6, alpha = 1.6;
  double eps_abs = 1e-3, eps_rel = 1e-3, eps_prim_inf = 1e-4, eps_dual_inf = 1e-4;
  int max_iter = 4000, check_termination = 25, scaling = 10;
  int adaptive_rho = 1, adaptive_rho_interval = 50;
  double adaptive_rho_tolerance = 5.0;
};

struct QpProblem {
  int n = 0, m = 0;
  Mat P, q, A, l, u;
  // optional warm start (osqp_warm_start: x = D^-1 x0, z = A x, y = c E^-1 y0); null = OSQP's cold start, which is what the
  // reference uses (QP_base.h:146).  Vectors in this problem's own variable / row order.
  const double* x0 = nullptr;
  const double* y0 = nullptr;
  void resize(int n_, int m_) {
    n = n_; m = m_;
    P.assign(n * n, 0.0); q.assign(n, 0.0); A.assign(m * n, 0.0);
    l.assign(m, -OSQP_INFTY); u.assign(m, OSQP_INFTY);
  }
};

struct QpResult {
  int status = 0, iters = 0, rho_updates = 0;
  double pri_res = 0, dua_res = 0, rho = 0;
  // diagnostics of the adaptive-rho decisions (schedule studies): smallest | log(rho_new / rho) | - log(tolerance) | over the
  // adaptation points (how close a decision was to flipping), and log(rho_new / rho) at the first one
  double rho_margin = 1e9, rho_first = 0;
  Mat x, y, z;
};

inline double inf_norm(const Mat& v) { double r = 0; for (double a : v) r = std::max(r, std::fabs(a)); return r; }
inline void limit_scaling(double& d) {
  d = d < OSQP_MIN_SCALING ? 1.0 : d;
  d = d > OSQP_MAX_SCALING ? OSQP_MAX_SCALING : d;
}

struct QpWork {
  // compressed rows of the scaled A (zeros skipped, like OSQP's CSC storage after sparseView())
  std::vector<int> rptr, cidx;
  std::vector<double> aval;
  Mat P, q, l, u, D, E, Dinv, Einv, rho_vec, K, x, z, y, xt, zt, xp, zp, dx, dy, Ax, Px, Aty, tmpn, tmpm;
  std::vector<int> ctype;
  double c = 1, cinv = 1;
};

inline void a_mul(const QpWork& w, int m, const double* x, double* out) {
  for (int i = 0; i < m; ++i) {
    double s = 0;
    for (int k = w.rptr[i]; k < w.rptr[i + 1]; ++k) s += w.aval[k] * x[w.cidx[k]];
    out[i] = s;
  }
}
inline void at_mul(const QpWork& w, int m, int n, const double* y, double* out) {
  for (int j = 0; j < n; ++j) out[j] = 0;
  for (int i = 0; i < m; ++i)
    for (int k = w.rptr[i]; k < w.rptr[i + 1]; ++k) out[w.cidx[k]] += w.aval[k] * y[i];
}

inline bool qp_factor(QpWork& w, int n, int m, double sigma) {
  // K = P + sigma I + A' diag(rho) A
  w.K = w.P;
  for (int j = 0; j < n; ++j) w.K[j * n + j] += sigma;
  for (int i = 0; i < m; ++i)
    for (int k1 = w.rptr[i]; k1 < w.rptr[i + 1]; ++k1)
      for (int k2 = w.rptr[i]; k2 < w.rptr[i + 1]; ++k2)
        w.K[w.cidx[k1] * n + w.cidx[k2]] += w.rho_vec[i] * w.aval[k1] * w.aval[k2];
  return cholesky(w.K.data(), n);
}

inline void qp_solve(const QpProblem& pb, const QpSettings& st, QpResult& res, QpWork& w) {
  const int n = pb.n, m = pb.m;
  // ---------------- setup: copy + scale_data (OSQP scaling.c)
  w.P = pb.P; w.q = pb.q; w.l = pb.l; w.u = pb.u;
  Mat As = pb.A;
  w.D.assign(n, 1.0); w.E.assign(m, 1.0);
  w.c = 1.0;
  Mat Dt(n), Et(m);
  for (int it = 0; it < st.scaling; ++it) {
    for (int j = 0; j < n; ++j) {
      double d = 0;
      for (int i = 0; i < n; ++i) d = std::max(d, std::fabs(w.P[i * n + j]));
      for (int i = 0; i < m; ++i) d = std::max(d, std::fabs(As[i * n + j]));
      Dt[j] = d;
    }
    for (int i = 0; i < m; ++i) {
      double e = 0;
      for (int j = 0; j < n; ++j) e = std::max(e, std::fabs(As[i * n + j]));
      Et[i] = e;
    }
    for (double& d : Dt) { limit_scaling(d); d = 1.0 / std::sqrt(d); }
    for (double& e : Et) { limit_scaling(e); e = 1.0 / std::sqrt(e); }
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) w.P[i * n + j] *= Dt[i] * Dt[j];
    for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) As[i * n + j] *= Et[i] * Dt[j];
    for (int j = 0; j < n; ++j) { w.q[j] *= Dt[j]; w.D[j] *= Dt[j]; }
    for (int i = 0; i < m; ++i) w.E[i] *= Et[i];
    // cost normalisation
    double mean = 0;
    for (int j = 0; j < n; ++j) {
      double d = 0;
      for (int i = 0; i < n; ++i) d = std::max(d, std::fabs(w.P[i * n + j]));
      mean += d;
    }
    double ct = n > 0 ? mean / n : 1.0;
    double qn = inf_norm(w.q);
    limit_scaling(qn);
    ct = std::max(ct, qn);
    limit_scaling(ct);
    ct = 1.0 / ct;
    for (double& p : w.P) p *= ct;
    for (double& qq : w.q) qq *= ct;
    w.c *= ct;
  }
  w.cinv = 1.0 / w.c;
  w.Dinv.resize(n); w.Einv.resize(m);
  for (int j = 0; j < n; ++j) w.Dinv[j] = 1.0 / w.D[j];
  for (int i = 0; i < m; ++i) { w.Einv[i] = 1.0 / w.E[i]; w.l[i] *= w.E[i]; w.u[i] *= w.E[i]; }
  // compressed rows
  w.rptr.assign(m + 1, 0); w.cidx.clear(); w.aval.clear();
  for (int i = 0; i < m; ++i) {
    for (int j = 0; j < n; ++j)
      if (pb.A[i * n + j] != 0.0) { w.cidx.push_back(j); w.aval.push_back(As[i * n + j]); }
    w.rptr[i + 1] = int(w.cidx.size());
  }
  // ---------------- rho vector (set_rho_vec)
  double rho = std::min(std::max(st.rho, OSQP_RHO_MIN), OSQP_RHO_MAX);
  w.rho_vec.resize(m); w.ctype.resize(m);
  auto set_rho = [&]() {
    for (int i = 0; i < m; ++i) {
      if (w.l[i] < -OSQP_INFTY * OSQP_MIN_SCALING && w.u[i] > OSQP_INFTY * OSQP_MIN_SCALING) { w.ctype[i] = -1; w.rho_vec[i] = OSQP_RHO_MIN; }
      else if (w.u[i] - w.l[i] < OSQP_RHO_TOL) { w.ctype[i] = 1; w.rho_vec[i] = OSQP_RHO_EQ_OVER_INEQ * rho; }
      else { w.ctype[i] = 0; w.rho_vec[i] = rho; }
    }
  };
  set_rho();
  res.x.assign(n, 0.0); res.y.assign(m, 0.0); res.z.assign(m, 0.0);
  res.status = 0; res.iters = 0; res.rho_updates = 0; res.rho_margin = 1e9; res.rho_first = 0;
  if (!qp_factor(w, n, m, st.sigma)) { res.status = QP_NON_CVX; return; }
  // ---------------- ADMM (osqp_solve), cold start
  w.x.assign(n, 0.0); w.z.assign(m, 0.0); w.y.assign(m, 0.0);
  if (pb.x0 && pb.y0) {
    for (int j = 0; j < n; ++j) w.x[j] = w.Dinv[j] * pb.x0[j];
    a_mul(w, m, w.x.data(), w.z.data());
    for (int i = 0; i < m; ++i) w.y[i] = w.c * w.Einv[i] * pb.y0[i];
  }
  w.xp.assign(n, 0.0); w.zp.assign(m, 0.0); w.xt.assign(n, 0.0); w.zt.assign(m, 0.0);
  w.dx.assign(n, 0.0); w.dy.assign(m, 0.0); w.Ax.assign(m, 0.0); w.Px.assign(n, 0.0); w.Aty.assign(n, 0.0);
  w.tmpn.assign(n, 0.0); w.tmpm.assign(m, 0.0);
  double pri_res = 0, dua_res = 0, pri_res_s = 0, dua_res_s = 0;
  auto update_info = [&]() {
    // primal residual ||Einv (Ax - z)||, dual residual cinv ||Dinv (Px + q + A'y)||
    a_mul(w, m, w.x.data(), w.Ax.data());
    pri_res = 0; pri_res_s = 0;
    for (int i = 0; i < m; ++i) {
      double r = w.Ax[i] - w.z[i];
      pri_res = std::max(pri_res, std::fabs(w.Einv[i] * r));
      pri_res_s = std::max(pri_res_s, std::fabs(r));
    }
    matvec(w.P.data(), w.x.data(), w.Px.data(), n, n);
    at_mul(w, m, n, w.y.data(), w.Aty.data());
    dua_res = 0; dua_res_s = 0;
    for (int j = 0; j < n; ++j) {
      double r = w.q[j] + w.Px[j] + w.Aty[j];
      dua_res = std::max(dua_res, std::fabs(w.Dinv[j] * r));
      dua_res_s = std::max(dua_res_s, std::fabs(r));
    }
    dua_res *= w.cinv;
  };
  auto is_primal_infeasible = [&]() {
    for (int i = 0; i < m; ++i) {
      if (w.u[i] > OSQP_INFTY * OSQP_MIN_SCALING) {
        if (w.l[i] < -OSQP_INFTY * OSQP_MIN_SCALING) w.dy[i] = 0.0;
        else w.dy[i] = std::min(w.dy[i], 0.0);
      } else if (w.l[i] < -OSQP_INFTY * OSQP_MIN_SCALING) {
        w.dy[i] = std::max(w.dy[i], 0.0);
      }
    }
    double nd = 0;
    for (int i = 0; i < m; ++i) nd = std::max(nd, std::fabs(w.E[i] * w.dy[i]));
    if (nd > st.eps_prim_inf) {
      double lhs = 0;
      for (int i = 0; i < m; ++i) lhs += w.u[i] * std::max(w.dy[i], 0.0) + w.l[i] * std::min(w.dy[i], 0.0);
      if (lhs < -st.eps_prim_inf * nd) {
        at_mul(w, m, n, w.dy.data(), w.tmpn.data());
        double na = 0;
        for (int j = 0; j < n; ++j) na = std::max(na, std::fabs(w.Dinv[j] * w.tmpn[j]));
        return na < st.eps_prim_inf * nd;
      }
    }
    return false;
  };
  auto is_dual_infeasible = [&]() {
    double ndx = 0;
    for (int j = 0; j < n; ++j) ndx = std::max(ndx, std::fabs(w.D[j] * w.dx[j]));
    const double cs = w.c;
    if (ndx > st.eps_dual_inf) {
      double qdx = 0;
      for (int j = 0; j < n; ++j) qdx += w.q[j] * w.dx[j];
      if (qdx < -cs * st.eps_dual_inf * ndx) {
        matvec(w.P.data(), w.dx.data(), w.tmpn.data(), n, n);
        double np = 0;
        for (int j = 0; j < n; ++j) np = std::max(np, std::fabs(w.Dinv[j] * w.tmpn[j]));
        if (np < cs * st.eps_dual_inf * ndx) {
          a_mul(w, m, w.dx.data(), w.tmpm.data());
          for (int i = 0; i < m; ++i) {
            double a = w.Einv[i] * w.tmpm[i];
            if ((w.u[i] < OSQP_INFTY * OSQP_MIN_SCALING && a > st.eps_dual_inf * ndx) ||
                (w.l[i] > -OSQP_INFTY * OSQP_MIN_SCALING && a < -st.eps_dual_inf * ndx))
              return false;
          }
          return true;
        }
      }
    }
    return false;
  };
  auto check_termination = [&](bool approximate) {
    double ea = st.eps_abs, er = st.eps_rel, epi = st.eps_prim_inf, edi = st.eps_dual_inf;
    if (approximate) { ea *= 10; er *= 10; epi *= 10; edi *= 10; }
    if (pri_res > OSQP_INFTY || dua_res > OSQP_INFTY) { res.status = QP_NON_CVX; return true; }
    bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
    if (m == 0) prim_ok = true;
    else {
      double nz = 0, nax = 0;
      for (int i = 0; i < m; ++i) { nz = std::max(nz, std::fabs(w.Einv[i] * w.z[i])); nax = std::max(nax, std::fabs(w.Einv[i] * w.Ax[i])); }
      double eps_prim = ea + er * std::max(nz, nax);
      if (pri_res < eps_prim) prim_ok = true;
      else prim_inf = is_primal_infeasible();
    }
    double nq = 0, naty = 0, npx = 0;
    for (int j = 0; j < n; ++j) {
      nq = std::max(nq, std::fabs(w.Dinv[j] * w.q[j]));
      naty = std::max(naty, std::fabs(w.Dinv[j] * w.Aty[j]));
      npx = std::max(npx, std::fabs(w.Dinv[j] * w.Px[j]));
    }
    double eps_dual = ea + er * w.cinv * std::max(nq, std::max(naty, npx));
    if (dua_res < eps_dual) dual_ok = true;
    else dual_inf = is_dual_infeasible();
    if (prim_ok && dual_ok) { res.status = approximate ? QP_SOLVED_INACCURATE : QP_SOLVED; return true; }
    if (prim_inf) { res.status = QP_PRIMAL_INFEASIBLE; return true; }
    if (dual_inf) { res.status = QP_DUAL_INFEASIBLE; return true; }
    return false;
  };
  int iter;
  bool done = false;
  for (iter = 1; iter <= st.max_iter; ++iter) {
    std::swap(w.x, w.xp); std::swap(w.z, w.zp);
    // update_xz_tilde
    for (int i = 0; i < m; ++i) w.tmpm[i] = w.rho_vec[i] * w.zp[i] - w.y[i];
    at_mul(w, m, n, w.tmpm.data(), w.xt.data());
    for (int j = 0; j < n; ++j) w.xt[j] += st.sigma * w.xp[j] - w.q[j];
    chol_solve(w.K.data(), w.xt.data(), n);
    a_mul(w, m, w.xt.data(), w.zt.data());
    // update_x, update_z, update_y
    for (int j = 0; j < n; ++j) { w.x[j] = st.alpha * w.xt[j] + (1 - st.alpha) * w.xp[j]; w.dx[j] = w.x[j] - w.xp[j]; }
    for (int i = 0; i < m; ++i) {
      double zr = st.alpha * w.zt[i] + (1 - st.alpha) * w.zp[i];
      double zz = zr + w.y[i] / w.rho_vec[i];
      w.z[i] = std::min(std::max(zz, w.l[i]), w.u[i]);
      w.dy[i] = w.rho_vec[i] * (zr - w.z[i]);
      w.y[i] += w.dy[i];
    }
    bool can_check = st.check_termination && (iter % st.check_termination == 0);
    if (can_check) {
      update_info();
      if (check_termination(false)) { done = true; break; }
    }
    if (st.adaptive_rho && st.adaptive_rho_interval && (iter % st.adaptive_rho_interval == 0)) {
      if (!can_check) update_info();
      // compute_rho_estimate: scaled residuals, normalised
      double nz = inf_norm(w.z), nax = inf_norm(w.Ax);
      double pr = pri_res_s / (std::max(nz, nax) + 1e-10);
      double dn = std::max(inf_norm(w.q), std::max(inf_norm(w.Aty), inf_norm(w.Px)));
      double dr = dua_res_s / (dn + 1e-10);
      double rho_new = rho * std::sqrt(pr / (dr + 1e-10));
      rho_new = std::min(std::max(rho_new, OSQP_RHO_MIN), OSQP_RHO_MAX);
      {
        const double lr = std::log(rho_new / rho);
        if (res.rho_margin > 1e8) res.rho_first = lr;
        res.rho_margin = std::min(res.rho_margin, std::fabs(std::fabs(lr) - std::log(st.adaptive_rho_tolerance)));
      }
      if (rho_new > rho * st.adaptive_rho_tolerance || rho_new < rho / st.adaptive_rho_tolerance) {
        rho = rho_new;
        for (int i = 0; i < m; ++i) {
          if (w.ctype[i] == 0) w.rho_vec[i] = rho;
          else if (w.ctype[i] == 1) w.rho_vec[i] = OSQP_RHO_EQ_OVER_INEQ * rho;
        }
        qp_factor(w, n, m, st.sigma);
        res.rho_updates++;
      }
    }
  }
  if (!done) {
    iter = st.max_iter;
    if (!(st.check_termination && (iter % st.check_termination == 0))) { update_info(); check_termination(false); }
    if (res.status == 0 && !check_termination(true)) res.status = QP_MAX_ITER;
  }
  res.iters = iter;
  res.pri_res = pri_res; res.dua_res = dua_res; res.rho = rho;
  for (int j = 0; j < n; ++j) res.x[j] = w.D[j] * w.x[j];
  for (int i = 0; i < m; ++i) { res.y[i] = w.cinv * w.E[i] * w.y[i]; res.z[i] = w.Einv[i] * w.z[i]; }
}

}  // namespace orc
