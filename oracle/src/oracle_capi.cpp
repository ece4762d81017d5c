// ORACLE -- test infrastructure only: loaded through ctypes by tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs.  Never linked or imported by the product.
// PARITY UNPINNED: the reference ships no tests or golden vectors and its arithmetic lives in
// un-vendored Pinocchio / hpp-fcl / OSQP; this library restates that arithmetic.
//
// All batch arrays are batch-major ("AoS"): element (b, k) of a (B, K) array sits at [b*K + k].
// Poses are 12 doubles: the top three rows of the 4x4 homogeneous matrix, row-major.
#include <memory>
#include <omp.h>

#include "omoma.h"

using namespace orc;

struct OrcHandle {
  Model m;
  CtrlParams cp;
  GeomParams gp;
  QpSettings qs;
  int threads = 1;
  // CPU-baseline mode: 1 = a fresh workspace (solver data, scratch vectors) is heap-allocated and released EVERY control cycle, as
  // the reference does (new OsqpEigen::Solver per solveQP, QP_base.h:143-177; new pinocchio::Data per getManipulability,
  // robot_data.cpp:542); 0 = one preallocated workspace per thread (optimistic)
  int fresh_workspace = 0;
  // optional per-robot solver diagnostics of orc_cycle*: [rho_margin, rho_first, rho_updates, final rho] (schedule studies)
  double* diag = nullptr;
};

static SE3 pose_from12(const double* a) {
  SE3 T;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T.R(i, j) = a[4 * i + j];
    T.p[i] = a[4 * i + 3];
  }
  return T;
}
static void pose_to12(const SE3& T, double* a) {
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) a[4 * i + j] = T.R(i, j);
    a[4 * i + 3] = T.p[i];
  }
}
static M3 m3_from(const double* a) { M3 R; for (int i = 0; i < 9; ++i) R.m[i] = a[i]; return R; }

extern "C" {

OrcHandle* orc_model_create(int nv, const int* parent, const int* jtype, const double* axis, const double* jR,
                            const double* jp, const double* mass, const double* com, const double* inertia,
                            const double* q_lo, const double* q_hi, const double* v_lim, int nf, const int* frame_parent,
                            const double* frame_R, const double* frame_p, int ng, const int* geom_type,
                            const double* geom_param, const int* geom_parent, const double* geom_R, const double* geom_p,
                            int npairs, const int* pairs, const double* gravity) {
  if (nv > MAXV) return nullptr;
  OrcHandle* h = new OrcHandle;
  Model& m = h->m;
  m.nv = nv;
  for (int i = 0; i < nv; ++i) {
    m.parent[i] = parent[i];
    m.jtype[i] = jtype[i];
    m.axis[i] = V3(axis[3 * i], axis[3 * i + 1], axis[3 * i + 2]);
    m.jplace[i] = SE3(m3_from(jR + 9 * i), V3(jp[3 * i], jp[3 * i + 1], jp[3 * i + 2]));
    m.inertia[i].m = mass[i];
    m.inertia[i].c = V3(com[3 * i], com[3 * i + 1], com[3 * i + 2]);
    m.inertia[i].I = m3_from(inertia + 9 * i);
    m.q_lo[i] = q_lo[i]; m.q_hi[i] = q_hi[i]; m.v_lim[i] = v_lim[i];
  }
  for (int i = 0; i < nv; ++i)
    for (int j = 0; j < nv; ++j) {
      bool a = false;
      for (int k = i; k >= 0; k = m.parent[k]) if (k == j) { a = true; break; }
      m.anc[i][j] = a;
    }
  m.gravity = V3(gravity[0], gravity[1], gravity[2]);
  m.nf = nf;
  for (int f = 0; f < nf; ++f) {
    m.frame_parent.push_back(frame_parent[f]);
    m.frame_place.push_back(SE3(m3_from(frame_R + 9 * f), V3(frame_p[3 * f], frame_p[3 * f + 1], frame_p[3 * f + 2])));
  }
  m.ng = ng;
  for (int g = 0; g < ng; ++g) {
    m.geom_type.push_back(geom_type[g]);
    m.geom_parent.push_back(geom_parent[g]);
    m.geom_param.push_back(V3(geom_param[3 * g], geom_param[3 * g + 1], geom_param[3 * g + 2]));
    m.geom_place.push_back(SE3(m3_from(geom_R + 9 * g), V3(geom_p[3 * g], geom_p[3 * g + 1], geom_p[3 * g + 2])));
  }
  for (int k = 0; k < npairs; ++k) { m.pair_a.push_back(pairs[2 * k]); m.pair_b.push_back(pairs[2 * k + 1]); }
  return h;
}
// mesh geometry (GEOM_CONVEX): hull vertices per geometry, from the Python loader (scipy / Qhull)
void orc_model_set_hulls(OrcHandle* h, int nvert_total, const double* verts, const int* off, const int* n) {
  Model& m = h->m;
  m.hull.assign(verts, verts + 3 * (size_t)nvert_total);
  m.hull_off.assign(off, off + m.ng);
  m.hull_n.assign(n, n + m.ng);
}
void orc_model_destroy(OrcHandle* h) { delete h; }
void orc_set_threads(OrcHandle* h, int t) { h->threads = t > 0 ? t : 1; }
void orc_set_fresh_workspace(OrcHandle* h, int on) { h->fresh_workspace = on != 0; }
void orc_set_diag(OrcHandle* h, double* buf) { h->diag = buf; }
void orc_set_task_gains(OrcHandle* h, const double* kp, const double* kv) {
  for (int i = 0; i < 6; ++i) { h->cp.Kp_task[i] = kp[i]; h->cp.Kv_task[i] = kv[i]; }
}
void orc_set_joint_gains(OrcHandle* h, const double* kp, const double* kv) {
  for (int i = 0; i < h->m.nv; ++i) { h->cp.Kp_joint[i] = kp[i]; h->cp.Kv_joint[i] = kv[i]; }
}
void orc_set_qp_settings(OrcHandle* h, double rho, double sigma, double alpha, double eps_abs, double eps_rel,
                         double eps_prim_inf, double eps_dual_inf, int max_iter, int check_termination, int scaling,
                         int adaptive_rho, int adaptive_rho_interval, double adaptive_rho_tolerance) {
  QpSettings& s = h->qs;
  s.rho = rho; s.sigma = sigma; s.alpha = alpha; s.eps_abs = eps_abs; s.eps_rel = eps_rel;
  s.eps_prim_inf = eps_prim_inf; s.eps_dual_inf = eps_dual_inf; s.max_iter = max_iter;
  s.check_termination = check_termination; s.scaling = scaling; s.adaptive_rho = adaptive_rho;
  s.adaptive_rho_interval = adaptive_rho_interval; s.adaptive_rho_tolerance = adaptive_rho_tolerance;
}
void orc_set_geom_params(OrcHandle* h, double gjk_tol, int gjk_max_iter, double epa_tol, int epa_max_iter) {
  h->gp.gjk_tol = gjk_tol; h->gp.gjk_max_iter = gjk_max_iter; h->gp.epa_tol = epa_tol; h->gp.epa_max_iter = epa_max_iter;
}

// updateState + cached getters for one frame.  Any output pointer may be NULL.
void orc_update_state(OrcHandle* h, int B, const double* q, const double* qd, int frame, double* pose, double* J,
                      double* Jdot, double* M, double* Minv, double* g, double* nle, double* oMi) {
  const Model& m = h->m;
  const int n = m.nv;
#pragma omp parallel num_threads(h->threads)
  {
    State s;
    std::vector<double> tmp(6 * n);
#pragma omp for schedule(static)
    for (int b = 0; b < B; ++b) {
      update_state(m, s, q + b * n, qd + b * n);
      if (pose) pose_to12(frame_pose(m, s.oMi, frame), pose + 12 * b);
      if (J) frame_jacobian(m, s.oMi, s.J, frame, J + 6 * n * b);
      if (Jdot) frame_jacobian_time_variation(m, s.oMi, s.ov, s.J, s.dJ, frame, Jdot + 6 * n * b);
      if (M) std::copy(s.M, s.M + n * n, M + n * n * b);
      if (Minv) std::copy(s.Minv, s.Minv + n * n, Minv + n * n * b);
      if (g) std::copy(s.g, s.g + n, g + n * b);
      if (nle) std::copy(s.nle, s.nle + n, nle + n * b);
      if (oMi) for (int i = 0; i < n; ++i) pose_to12(s.oMi[i], oMi + 12 * (n * b + i));
    }
  }
}

void orc_manipulability(OrcHandle* h, int B, const double* q, const double* qd, int frame, int with_graddot, double* mani,
                        double* grad, double* grad_dot) {
  const Model& m = h->m;
  const int n = m.nv;
#pragma omp parallel num_threads(h->threads)
  {
    State s;
    ManipResult r;
#pragma omp for schedule(static)
    for (int b = 0; b < B; ++b) {
      update_state(m, s, q + b * n, qd + b * n);
      manipulability(m, s, frame, true, with_graddot != 0, 0, n, r);
      mani[b] = r.m;
      if (grad) std::copy(r.grad, r.grad + n, grad + n * b);
      if (grad_dot) std::copy(r.grad_dot, r.grad_dot + n, grad_dot + n * b);
    }
  }
}

void orc_min_distance(OrcHandle* h, int B, const double* q, const double* qd, int with_graddot, double* dist, double* grad,
                      double* grad_dot, int* pair, double* pa, double* pb) {
  const Model& m = h->m;
  const int n = m.nv;
#pragma omp parallel num_threads(h->threads)
  {
    State s;
    MinDistResult r;
#pragma omp for schedule(static)
    for (int b = 0; b < B; ++b) {
      update_state(m, s, q + b * n, qd + b * n);
      min_distance(m, s, true, with_graddot != 0, h->gp, r);
      dist[b] = r.d;
      if (grad) std::copy(r.grad, r.grad + n, grad + n * b);
      if (grad_dot) std::copy(r.grad_dot, r.grad_dot + n, grad_dot + n * b);
      if (pair) pair[b] = r.pair;
      if (pa) { pa[3 * b] = r.pa.x; pa[3 * b + 1] = r.pa.y; pa[3 * b + 2] = r.pa.z; }
      if (pb) { pb[3 * b] = r.pb.x; pb[3 * b + 1] = r.pb.y; pb[3 * b + 2] = r.pb.z; }
    }
  }
}

// all pair distances of one configuration (for geometry unit tests)
void orc_pair_distances(OrcHandle* h, const double* q, double* d_out, double* pa, double* pb, int* gjk_iters) {
  const Model& m = h->m;
  State s;
  std::vector<double> qd(m.nv, 0.0);
  update_state(m, s, q, qd.data());
  for (size_t k = 0; k < m.pair_a.size(); ++k) {
    int ga = m.pair_a[k], gb = m.pair_b[k];
    Shape A = make_shape(m, ga, m.geom_parent[ga] < 0 ? m.geom_place[ga] : s.oMi[m.geom_parent[ga]] * m.geom_place[ga]);
    Shape Bs = make_shape(m, gb, m.geom_parent[gb] < 0 ? m.geom_place[gb] : s.oMi[m.geom_parent[gb]] * m.geom_place[gb]);
    DistResult r = shape_distance(A, Bs, h->gp);
    d_out[k] = r.d;
    if (pa) { pa[3 * k] = r.pa.x; pa[3 * k + 1] = r.pa.y; pa[3 * k + 2] = r.pa.z; }
    if (pb) { pb[3 * k] = r.pb.x; pb[3 * k + 1] = r.pb.y; pb[3 * k + 2] = r.pb.z; }
    if (gjk_iters) gjk_iters[k] = r.gjk_iters + 1000 * r.epa_iters;
  }
}

// generic shape pair (type, param[3], pose12) for unit tests
double orc_shape_distance(OrcHandle* h, int ta, const double* pa_, const double* Ta, int tb, const double* pb_,
                          const double* Tb, double* wa, double* wb, int* iters) {
  Shape A, Bs;
  A.type = ta; A.prm = V3(pa_[0], pa_[1], pa_[2]); A.T = pose_from12(Ta);
  Bs.type = tb; Bs.prm = V3(pb_[0], pb_[1], pb_[2]); Bs.T = pose_from12(Tb);
  GeomParams gp = h ? h->gp : GeomParams();
  DistResult r = shape_distance(A, Bs, gp);
  if (wa) { wa[0] = r.pa.x; wa[1] = r.pa.y; wa[2] = r.pa.z; }
  if (wb) { wb[0] = r.pb.x; wb[1] = r.pb.y; wb[2] = r.pb.z; }
  if (iters) { iters[0] = r.gjk_iters; iters[1] = r.epa_iters; }
  return r.d;
}

// Build the dense QP of one robot. kind: 0 = QPIK, 1 = QPID.  Sizes via orc_qp_sizes.
void orc_qp_sizes(OrcHandle* h, int kind, int* nx, int* nc) {
  const int n = h->m.nv;
  if (kind == 0) { *nx = 3 * n + 2; *nc = *nx + 2 * n + 2; }
  else { *nx = 6 * n + 2; *nc = *nx + 4 * n + 2 + n; }
}
void orc_build_qp(OrcHandle* h, int kind, const double* q, const double* qd, const double* des, int frame, double* P,
                  double* qv, double* A, double* l, double* u) {
  State s;
  update_state(h->m, s, q, qd);
  QpProblem pb;
  if (kind == 0) build_qpik(h->m, s, frame, des, h->cp, h->gp, pb);
  else build_qpid(h->m, s, frame, des, h->cp, h->gp, pb);
  std::copy(pb.P.begin(), pb.P.end(), P);
  std::copy(pb.q.begin(), pb.q.end(), qv);
  std::copy(pb.A.begin(), pb.A.end(), A);
  std::copy(pb.l.begin(), pb.l.end(), l);
  std::copy(pb.u.begin(), pb.u.end(), u);
}
// Generic dense OSQP-algorithm solve.
int orc_solve_qp(OrcHandle* h, int n, int m, const double* P, const double* qv, const double* A, const double* l,
                 const double* u, double* x, double* y, int* iters, double* info /* pri_res, dua_res, rho, rho_updates */) {
  QpProblem pb;
  pb.resize(n, m);
  std::copy(P, P + n * n, pb.P.begin());
  std::copy(qv, qv + n, pb.q.begin());
  std::copy(A, A + m * n, pb.A.begin());
  std::copy(l, l + m, pb.l.begin());
  std::copy(u, u + m, pb.u.begin());
  QpWork w;
  QpResult r;
  qp_solve(pb, h->qs, r, w);
  if (x) std::copy(r.x.begin(), r.x.end(), x);
  if (y) std::copy(r.y.begin(), r.y.end(), y);
  if (iters) *iters = r.iters;
  if (info) { info[0] = r.pri_res; info[1] = r.dua_res; info[2] = r.rho; info[3] = r.rho_updates; }
  return r.status;
}

// mode: 0 QPIK(xdot_des given, 6)      1 QPIKStep(x_target 12, xdot_target 6)
//       2 QPID(xddot_des given, 6)     3 QPIDStep(x_target, xdot_target)
// out: (B, n) qdot* (modes 0/1) or tau* (modes 2/3); status/iters: (B,)
void orc_cycle_xy(OrcHandle* h, int mode, int B, const double* q, const double* qd, const double* x_target,
                  const double* xdot_target, int frame, double* out, int* status, int* iters, double* qp_x, double* qp_y);
void orc_cycle(OrcHandle* h, int mode, int B, const double* q, const double* qd, const double* x_target,
               const double* xdot_target, int frame, double* out, int* status, int* iters, double* qp_x) {
  orc_cycle_xy(h, mode, B, q, qd, x_target, xdot_target, frame, out, status, iters, qp_x, nullptr);
}
// same, plus the dual vector y of the QP (rows in the reference's order [bounds; inequalities; equalities], QP_base.h:204-226)
void orc_cycle_xy(OrcHandle* h, int mode, int B, const double* q, const double* qd, const double* x_target,
                  const double* xdot_target, int frame, double* out, int* status, int* iters, double* qp_x, double* qp_y) {
  const Model& m = h->m;
  const int n = m.nv;
#pragma omp parallel num_threads(h->threads)
  {
    std::unique_ptr<Workspace> keep(new Workspace);
#pragma omp for schedule(dynamic, 16)
    for (int b = 0; b < B; ++b) {
      std::unique_ptr<Workspace> fresh(h->fresh_workspace ? new Workspace : nullptr);
      Workspace& ws = fresh ? *fresh : *keep;
      update_state(m, ws.s, q + b * n, qd + b * n);
      double des[6];
      if (mode == 0 || mode == 2) std::copy(xdot_target + 6 * b, xdot_target + 6 * b + 6, des);
      else desired_from_error(m, ws.s, frame, pose_from12(x_target + 12 * b), xdot_target + 6 * b, h->cp, true, des);
      int st;
      if (mode <= 1) st = ctrl_qpik(m, ws, frame, des, h->cp, h->gp, h->qs, out + n * b);
      else st = ctrl_qpid(m, ws, frame, des, h->cp, h->gp, h->qs, out + n * b, nullptr);
      if (status) status[b] = st;
      if (iters) iters[b] = ws.res.iters;
      if (h->diag) { double* d = h->diag + 4 * size_t(b); d[0] = ws.res.rho_margin; d[1] = ws.res.rho_first; d[2] = ws.res.rho_updates; d[3] = ws.res.rho; }
      if (qp_x) std::copy(ws.res.x.begin(), ws.res.x.end(), qp_x + size_t(b) * ws.res.x.size());
      if (qp_y) std::copy(ws.res.y.begin(), ws.res.y.end(), qp_y + size_t(b) * ws.res.y.size());
    }
  }
}

// QPIKStep / QPIDStep cycle with a WARM START from (qp_x, qp_y) -- the previous tick's primal / dual solution of every robot, in the
// oracle's own variable / row order -- which are then overwritten with this tick's (zeros after an infeasible / non-convex solve:
// the next tick of that robot starts cold, and all zeros IS the cold start).  An extension, not a reference behaviour (QP_base.h:146).
void orc_cycle_warm(OrcHandle* h, int mode, int B, const double* q, const double* qd, const double* x_target,
                    const double* xdot_target, int frame, double* out, int* status, int* iters, int nx, int ny, double* qp_x, double* qp_y) {
  const Model& m = h->m;
  const int n = m.nv;
#pragma omp parallel num_threads(h->threads)
  {
    std::unique_ptr<Workspace> keep(new Workspace);
#pragma omp for schedule(dynamic, 16)
    for (int b = 0; b < B; ++b) {
      Workspace& ws = *keep;
      update_state(m, ws.s, q + b * n, qd + b * n);
      double des[6];
      desired_from_error(m, ws.s, frame, pose_from12(x_target + 12 * b), xdot_target + 6 * b, h->cp, true, des);
      ws.pb.x0 = qp_x + size_t(b) * nx; ws.pb.y0 = qp_y + size_t(b) * ny;
      int st;
      if (mode <= 1) st = ctrl_qpik(m, ws, frame, des, h->cp, h->gp, h->qs, out + n * b);
      else st = ctrl_qpid(m, ws, frame, des, h->cp, h->gp, h->qs, out + n * b, nullptr);
      ws.pb.x0 = ws.pb.y0 = nullptr;
      if (status) status[b] = st;
      if (iters) iters[b] = ws.res.iters;
      const bool keep_sol = st == QP_SOLVED || st == QP_MAX_ITER || st == QP_SOLVED_INACCURATE;
      if ((int)ws.res.x.size() != nx || (int)ws.res.y.size() != ny) { if (status) status[b] = -1; continue; }
      for (int i = 0; i < nx; ++i) qp_x[size_t(b) * nx + i] = keep_sol ? ws.res.x[i] : 0.0;
      for (int i = 0; i < ny; ++i) qp_y[size_t(b) * ny + i] = keep_sol ? ws.res.y[i] : 0.0;
    }
  }
}

// desired task signal of the Step controllers: Kp e + Kv edot (QPIKStep / QPIDStep, robot_controller.cpp:292-300, 335-345)
void orc_desired_task(OrcHandle* h, int mode, int B, const double* q, const double* qd, const double* x_target,
                      const double* xdot_target, int frame, double* des) {
  (void)mode;
  const Model& m = h->m;
  const int n = m.nv;
  State s;
  for (int b = 0; b < B; ++b) {
    update_state(m, s, q + b * n, qd + b * n);
    desired_from_error(m, s, frame, pose_from12(x_target + 12 * b), xdot_target + 6 * b, h->cp, true, des + 6 * b);
  }
}

// CLIKStep (mode 0) / OSFStep (mode 1) / OSF with given xddot in xdot_target (mode 2)
void orc_taskspace(OrcHandle* h, int mode, int B, const double* q, const double* qd, const double* x_target,
                   const double* xdot_target, const double* null_vec, int frame, double* out) {
  const Model& m = h->m;
  const int n = m.nv;
#pragma omp parallel num_threads(h->threads)
  {
    State s;
#pragma omp for schedule(static)
    for (int b = 0; b < B; ++b) {
      update_state(m, s, q + b * n, qd + b * n);
      const double* nv = null_vec ? null_vec + n * b : nullptr;
      if (mode == 0) ctrl_clik_step(m, s, frame, pose_from12(x_target + 12 * b), xdot_target + 6 * b, nv, h->cp, out + n * b);
      else if (mode == 1) {
        double des[6];
        desired_from_error(m, s, frame, pose_from12(x_target + 12 * b), xdot_target + 6 * b, h->cp, true, des);
        ctrl_osf(m, s, frame, des, nv, out + n * b);
      } else ctrl_osf(m, s, frame, xdot_target + 6 * b, nv, out + n * b);
    }
  }
}

void orc_joint_torque_step(OrcHandle* h, int B, const double* q, const double* qd, const double* q_t, const double* qd_t,
                           double* tau) {
  const Model& m = h->m;
  const int n = m.nv;
#pragma omp parallel num_threads(h->threads)
  {
    State s;
#pragma omp for schedule(static)
    for (int b = 0; b < B; ++b) {
      update_state(m, s, q + b * n, qd + b * n);
      ctrl_joint_torque_step(m, s, q_t + n * b, qd_t + n * b, h->cp, tau + n * b);
    }
  }
}

void orc_task_space_cubic(const double* x_target, const double* xdot_target, const double* x_init, const double* xdot_init,
                          double t, double t0, double dur, double* x_des, double* xdot_des) {
  SE3 xd;
  task_space_cubic(pose_from12(x_target), xdot_target, pose_from12(x_init), xdot_init, t, t0, dur, xd, xdot_des);
  pose_to12(xd, x_des);
}

void orc_pinv(const double* A, int m, int n, double* out) { pinv_cod(A, m, n, out); }


// ---- mobile base / mobile manipulator (omoma.h)
void orc_model_set_moma(OrcHandle* h, int drive_type, double wheel_radius, double base_width, double wheel_offset, int wheel_num,
                        const double* roller_angles, const double* b2w_x, const double* b2w_y, const double* b2w_ang,
                        int virtual_start, int mani_start, int mobi_start, int act_mani_start, int act_mobi_start) {
  Model& m = h->m;
  m.drive_type = drive_type; m.wheel_radius = wheel_radius; m.base_width = base_width; m.wheel_offset = wheel_offset;
  m.wheel_num = wheel_num; m.virtual_start = virtual_start; m.mani_start = mani_start; m.mobi_start = mobi_start;
  m.act_mani_start = act_mani_start; m.act_mobi_start = act_mobi_start;
  const int np = drive_type == 2 ? wheel_num / 2 : wheel_num;
  m.roller_angles.assign(wheel_num, 0.0); m.b2w_x.assign(np, 0.0); m.b2w_y.assign(np, 0.0); m.b2w_ang.assign(wheel_num, 0.0);
  for (int i = 0; i < wheel_num; ++i) { if (roller_angles) m.roller_angles[i] = roller_angles[i]; if (b2w_ang) m.b2w_ang[i] = b2w_ang[i]; }
  for (int i = 0; i < np; ++i) { if (b2w_x) m.b2w_x[i] = b2w_x[i]; if (b2w_y) m.b2w_y[i] = b2w_y[i]; }
}
// Mobile::RobotData: FK Jacobian (3 x w) and base velocity for B wheel states
void orc_mobile_state(OrcHandle* h, int B, const double* wheel_pos, const double* wheel_vel, double* J, double* base_vel) {
  const Model& m = h->m;
  const int w = m.wheel_num;
  for (int b = 0; b < B; ++b) {
    double Jm[3 * 8];
    mobile_fk_jacobian(m, wheel_pos + w * b, Jm);
    if (J) std::copy(Jm, Jm + 3 * w, J + 3 * w * b);
    if (base_vel)
      for (int r = 0; r < 3; ++r) {
        double v = 0;
        for (int k = 0; k < w; ++k) v += Jm[r * w + k] * wheel_vel[w * b + k];
        base_vel[3 * b + r] = v;
      }
  }
}
// Mobile::RobotData / Mobile::RobotController without a URDF: the base is its KinematicParam (type_define.h:58-72).
// fk != 0: J (B, 3, w) and out = base velocity (B, 3) from in = wheel velocities (B, w)
// fk == 0: J (B, w, 3) and out = wheel velocities (B, w) from in = base velocity (B, 3) [VelocityCommand when saturate]
void orc_mobile_base(int drive_type, double wheel_radius, double base_width, double wheel_offset, double max_lin_speed,
                     double max_ang_speed, int wheel_num, const double* roller_angles, const double* b2w_x, const double* b2w_y,
                     const double* b2w_ang, int fk, int saturate, int B, const double* wheel_pos, const double* in, double* J,
                     double* out) {
  Model m;
  m.drive_type = drive_type; m.wheel_radius = wheel_radius; m.base_width = base_width; m.wheel_offset = wheel_offset;
  m.wheel_num = wheel_num;
  const int w = wheel_num, np = drive_type == 2 ? w / 2 : w;
  m.roller_angles.assign(w, 0.0); m.b2w_x.assign(np, 0.0); m.b2w_y.assign(np, 0.0); m.b2w_ang.assign(w, 0.0);
  for (int i = 0; i < w; ++i) { if (roller_angles) m.roller_angles[i] = roller_angles[i]; if (b2w_ang) m.b2w_ang[i] = b2w_ang[i]; }
  for (int i = 0; i < np; ++i) { if (b2w_x) m.b2w_x[i] = b2w_x[i]; if (b2w_y) m.b2w_y[i] = b2w_y[i]; }
  std::vector<double> zeros(w, 0.0), Jm(3 * w);
  for (int b = 0; b < B; ++b) {
    const double* wp = wheel_pos ? wheel_pos + w * b : zeros.data();
    if (fk) {
      mobile_fk_jacobian(m, wp, Jm.data());
      if (J) std::copy(Jm.begin(), Jm.end(), J + 3 * w * b);
      if (out && in)
        for (int r = 0; r < 3; ++r) {
          double v = 0;
          for (int k = 0; k < w; ++k) v += Jm[r * w + k] * in[w * b + k];
          out[3 * b + r] = v;
        }
    } else {
      if (J) mobile_ik_jacobian(m, wp, J + 3 * w * b);
      if (out && in) mobile_wheel_velocity(m, max_lin_speed, max_ang_speed, wp, in + 3 * b, saturate != 0, out + w * b);
    }
  }
}
// MobileManipulator::RobotData::updateState on full-dof vectors: S (n x act), M~, M~^-1, g~, nle~, J~, J~dot, manipulability
void orc_moma_update_state(OrcHandle* h, int B, const double* q, const double* qd, int frame, double* S, double* Mact,
                           double* Minv_act, double* g_act, double* nle_act, double* Jt, double* Jtd, double* mani,
                           double* mani_grad, double* mani_graddot) {
  const Model& m = h->m;
  const int n = m.nv;
#pragma omp parallel num_threads(h->threads)
  {
    State s;
    MomaState ms;
#pragma omp for schedule(static)
    for (int b = 0; b < B; ++b) {
      update_state(m, s, q + b * n, qd + b * n);
      moma_update(m, s, ms);
      const int act = ms.act, k = ms.mani;
      if (S) std::copy(ms.S, ms.S + n * act, S + size_t(b) * n * act);
      if (Mact) std::copy(ms.M, ms.M + act * act, Mact + size_t(b) * act * act);
      if (Minv_act) std::copy(ms.Minv, ms.Minv + act * act, Minv_act + size_t(b) * act * act);
      if (g_act) std::copy(ms.g, ms.g + act, g_act + size_t(b) * act);
      if (nle_act) std::copy(ms.nle, ms.nle + act, nle_act + size_t(b) * act);
      if (Jt) moma_jacobians(m, s, ms, frame, Jt + size_t(b) * 6 * act, Jtd ? Jtd + size_t(b) * 6 * act : nullptr);
      if (mani) {
        ManipResult mr;
        manipulability(m, s, frame, true, mani_graddot != nullptr, m.mani_start, k, mr);
        mani[b] = mr.m;
        if (mani_grad) std::copy(mr.grad, mr.grad + k, mani_grad + size_t(b) * k);
        if (mani_graddot) std::copy(mr.grad_dot, mr.grad_dot + k, mani_graddot + size_t(b) * k);
      }
    }
  }
}
void orc_moma_qp_sizes(OrcHandle* h, int kind, int* nx, int* nc) {
  const int n = h->m.nv, w = h->m.wheel_num, k = n - 3 - w, act = w + k;
  if (kind == 0) { *nx = act; *nc = act + 2 * k + 2; }
  else { *nx = 2 * act; *nc = 4 * k + 2 + act; }
}
void orc_moma_build_qp(OrcHandle* h, int kind, const double* q, const double* qd, const double* des, int frame, double* P,
                       double* qv, double* A, double* l, double* u) {
  Workspace ws;
  MomaState ms;
  update_state(h->m, ws.s, q, qd);
  moma_update(h->m, ws.s, ms);
  if (kind == 0) build_moma_qpik(h->m, ws.s, ms, frame, des, h->cp, h->gp, ws.pb);
  else build_moma_qpid(h->m, ws.s, ms, frame, des, h->cp, h->gp, ws.pb);
  std::copy(ws.pb.P.begin(), ws.pb.P.end(), P); std::copy(ws.pb.q.begin(), ws.pb.q.end(), qv);
  std::copy(ws.pb.A.begin(), ws.pb.A.end(), A); std::copy(ws.pb.l.begin(), ws.pb.l.end(), l);
  std::copy(ws.pb.u.begin(), ws.pb.u.end(), u);
}
// mode: 0 QPIK(xdot_des) 1 QPIKStep 2 QPID(xddot_des) 3 QPIDStep.  out: (B, act) eta* | tau*;  out2: (B, act) eta_dot* (modes 2/3)
void orc_moma_cycle_xy(OrcHandle* h, int mode, int B, const double* q, const double* qd, const double* x_target,
                       const double* xdot_target, int frame, double* out, double* out2, int* status, int* iters, double* qp_x,
                       double* qp_y);
void orc_moma_cycle(OrcHandle* h, int mode, int B, const double* q, const double* qd, const double* x_target,
                    const double* xdot_target, int frame, double* out, double* out2, int* status, int* iters) {
  orc_moma_cycle_xy(h, mode, B, q, qd, x_target, xdot_target, frame, out, out2, status, iters, nullptr, nullptr);
}
void orc_moma_cycle_xy(OrcHandle* h, int mode, int B, const double* q, const double* qd, const double* x_target,
                       const double* xdot_target, int frame, double* out, double* out2, int* status, int* iters, double* qp_x,
                       double* qp_y) {
  const Model& m = h->m;
  const int n = m.nv;
#pragma omp parallel num_threads(h->threads)
  {
    std::unique_ptr<Workspace> keep(new Workspace);
    MomaState ms;
#pragma omp for schedule(dynamic, 16)
    for (int b = 0; b < B; ++b) {
      std::unique_ptr<Workspace> fresh(h->fresh_workspace ? new Workspace : nullptr);
      Workspace& ws = fresh ? *fresh : *keep;
      update_state(m, ws.s, q + b * n, qd + b * n);
      moma_update(m, ws.s, ms);
      const int act = ms.act;
      double des[6];
      if (mode == 0 || mode == 2) std::copy(xdot_target + 6 * b, xdot_target + 6 * b + 6, des);
      else desired_from_error(m, ws.s, frame, pose_from12(x_target + 12 * b), xdot_target + 6 * b, h->cp, mode == 3, des);
      int st;
      if (mode <= 1) st = ctrl_moma_qpik(m, ws, ms, frame, des, h->cp, h->gp, h->qs, out + size_t(b) * act);
      else {
        double ed[MAXV];
        st = ctrl_moma_qpid(m, ws, ms, frame, des, h->cp, h->gp, h->qs, out2 ? out2 + size_t(b) * act : ed, out + size_t(b) * act);
      }
      if (status) status[b] = st;
      if (iters) iters[b] = ws.res.iters;
      if (qp_x) std::copy(ws.res.x.begin(), ws.res.x.end(), qp_x + size_t(b) * ws.res.x.size());
      if (qp_y) std::copy(ws.res.y.begin(), ws.res.y.end(), qp_y + size_t(b) * ws.res.y.size());
    }
  }
}

}  // extern "C"
