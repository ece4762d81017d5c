// ORACLE -- test infrastructure only. PARITY UNPINNED (no reference goldens exist).
// QP formulations and controller wrappers restated from the reference:
//   Manipulator::QPIK   src/manipulator/QP_IK.cpp:69-131      
//   Manipulator::QPID   src/manipulator/QP_ID.cpp:92-193      (B2)
//   RobotController     src/manipulator/robot_controller.cpp:115-125,156-171,208-247,277-360
//   DyrosMath helpers   include/math_type_define.h:62-144,235-298,633-685
#pragma once
#include "ogeom.h"
#include "oqp.h"

namespace orc {

struct CtrlParams {
  double alpha = 50.0;          // CBF gain (QP_IK.cpp:101, QP_ID.cpp:130)
  double slack_weight = 1000.0; // QP_IK.cpp:83-86
  double ik_reg = 1.0;          // QP_IK.cpp:81
  double moma_ik_reg = 0.01;    // mobile_manipulator/QP_IK.cpp:71
  double mani_thresh = 0.01;    // QP_IK.cpp:122
  double dist_thresh = 0.05;    // QP_IK.cpp:130
  double Kp_task[6] = {100, 100, 100, 100, 100, 100};  // robot_controller.cpp:12-13
  double Kv_task[6] = {20, 20, 20, 20, 20, 20};
  double Kp_joint[MAXV], Kv_joint[MAXV];               // robot_controller.cpp:14-15
  CtrlParams() { for (int i = 0; i < MAXV; ++i) { Kp_joint[i] = 400; Kv_joint[i] = 40; } }
};

// ------------------------------------------------------------------ DyrosMath
inline double cubic(double t, double t0, double tf, double x0, double xf, double v0, double vf) {
  if (t < t0) return x0;
  if (t > tf) return xf;
  double e = t - t0, T = tf - t0, T2 = T * T, T3 = T2 * T, dx = xf - x0;
  return x0 + v0 * e + (3 * dx / T2 - 2 * v0 / T - vf / T) * e * e + (-2 * dx / T3 + (v0 + vf) / T2) * e * e * e;
}
inline double cubic_dot(double t, double t0, double tf, double x0, double xf, double v0, double vf) {
  if (t < t0) return v0;
  if (t > tf) return vf;
  double e = t - t0, T = tf - t0, T2 = T * T, T3 = T2 * T, dx = xf - x0;
  return v0 + 2 * (3 * dx / T2 - 2 * v0 / T - vf / T) * e + 3 * (-2 * dx / T3 + (v0 + vf) / T2) * e * e;
}
// vee(log(R)) for a rotation matrix (principal branch), and exp(skew(w)) by Rodrigues.
inline V3 rot_log(const M3& R) {
  double tr = R(0, 0) + R(1, 1) + R(2, 2);
  double c = std::min(std::max(0.5 * (tr - 1), -1.0), 1.0);
  double th = std::acos(c);
  V3 w(R(2, 1) - R(1, 2), R(0, 2) - R(2, 0), R(1, 0) - R(0, 1));
  if (th < 1e-8) return 0.5 * w;
  if (M_PI - th < 1e-6) {  // near pi: axis from the symmetric part
    V3 ax;
    int k = 0;
    if (R(1, 1) > R(k, k)) k = 1;
    if (R(2, 2) > R(k, k)) k = 2;
    V3 col = V3(R(0, k), R(1, k), R(2, k));
    col[k] += 1.0;
    ax = (1.0 / norm(col)) * col;
    if (dot(ax, w) < 0) ax = -ax;
    return th * ax;
  }
  return (th / (2 * std::sin(th))) * w;
}
inline M3 rot_exp(const V3& w) {
  double th = norm(w);
  if (th < 1e-12) return M3::identity() + skew(w);
  return axis_angle((1.0 / th) * w, th);
}
// getPhi(current_rotation, desired_rotation) (math_type_define.h:283-298)
inline V3 get_phi(const M3& Rc, const M3& Rd) {
  V3 s;
  for (int i = 0; i < 3; ++i) s += cross(Rc.col(i), Rd.col(i));
  return -0.5 * s;
}
// getTaskSpaceError: note the target rotation is passed as getPhi's first argument (:642)
inline void task_space_error(const SE3& x_target, const double* xdot_target, const SE3& x, const double* xdot,
                             double* x_err, double* xdot_err) {
  V3 ep = x_target.p - x.p, eo = get_phi(x_target.R, x.R);
  x_err[0] = ep.x; x_err[1] = ep.y; x_err[2] = ep.z; x_err[3] = eo.x; x_err[4] = eo.y; x_err[5] = eo.z;
  for (int i = 0; i < 6; ++i) xdot_err[i] = xdot_target[i] - xdot[i];
}
// getTaskSpaceCubic (math_type_define.h:647-685)
inline void task_space_cubic(const SE3& x_target, const double* xdot_target, const SE3& x_init, const double* xdot_init,
                             double t, double t0, double dur, SE3& x_des, double* xdot_des) {
  const double tf = t0 + dur;
  for (int i = 0; i < 3; ++i) {
    x_des.p[i] = cubic(t, t0, tf, x_init.p[i], x_target.p[i], xdot_init[i], xdot_target[i]);
    xdot_des[i] = cubic_dot(t, t0, tf, x_init.p[i], x_target.p[i], xdot_init[i], xdot_target[i]);
  }
  V3 r = rot_log(transpose(x_init.R) * x_target.R);
  if (t >= tf) x_des.R = x_target.R;
  else if (t < t0) x_des.R = x_init.R;
  else {
    double tau = cubic(t, t0, tf, 0, 1, 0, 0);
    x_des.R = x_init.R * rot_exp(tau * r);
  }
  V3 rd(cubic_dot(t, t0, tf, 0, r.x, 0, 0), cubic_dot(t, t0, tf, 0, r.y, 0, 0), cubic_dot(t, t0, tf, 0, r.z, 0, 0));
  rd = x_init.R * rd;
  double tau = (t - t0) / (tf - t0);
  if (tau < 0 || tau > 1) rd = V3();
  xdot_des[3] = rd.x; xdot_des[4] = rd.y; xdot_des[5] = rd.z;
}

// ------------------------------------------------------------------ QP builders
// Manipulator::QPIK  x = [qdot(n); s_qmin(n); s_qmax(n); s_sing; s_col]
inline void build_qpik(const Model& m, const State& s, int frame, const double* xdot_des, const CtrlParams& cp,
                       const GeomParams& gp, QpProblem& pb) {
  const int n = m.nv, nx = 3 * n + 2, nineq = 2 * n + 2, nc = nx + nineq;
  pb.resize(nx, nc);
  std::vector<double> J(6 * n);
  frame_jacobian(m, s.oMi, s.J, frame, J.data());
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double v = 0;
      for (int r = 0; r < 6; ++r) v += J[r * n + i] * J[r * n + j];
      pb.P[i * nx + j] = 2.0 * v + (i == j ? cp.ik_reg : 0.0);
    }
  for (int i = 0; i < n; ++i) {
    double v = 0;
    for (int r = 0; r < 6; ++r) v += J[r * n + i] * xdot_des[r];
    pb.q[i] = -2.0 * v;
  }
  for (int i = n; i < nx; ++i) pb.q[i] = cp.slack_weight;
  // bounds (identity block first, QP_base.h:204-210)
  for (int i = 0; i < nx; ++i) pb.A[i * nx + i] = 1.0;
  for (int i = 0; i < n; ++i) { pb.l[i] = -m.v_lim[i]; pb.u[i] = m.v_lim[i]; }
  for (int i = n; i < nx; ++i) pb.l[i] = 0.0;
  // inequality rows
  const int r0 = nx;
  for (int i = 0; i < n; ++i) {
    pb.A[(r0 + i) * nx + i] = 1.0; pb.A[(r0 + i) * nx + n + i] = 1.0;
    pb.l[r0 + i] = -cp.alpha * (s.q[i] - m.q_lo[i]);
    pb.A[(r0 + n + i) * nx + i] = -1.0; pb.A[(r0 + n + i) * nx + 2 * n + i] = 1.0;
    pb.l[r0 + n + i] = -cp.alpha * (m.q_hi[i] - s.q[i]);
  }
  ManipResult mr;
  manipulability(m, s, frame, true, false, 0, n, mr);
  for (int i = 0; i < n; ++i) pb.A[(r0 + 2 * n) * nx + i] = mr.grad[i];
  pb.A[(r0 + 2 * n) * nx + 3 * n] = 1.0;
  pb.l[r0 + 2 * n] = -cp.alpha * (mr.m - cp.mani_thresh);
  MinDistResult md;
  min_distance(m, s, true, false, gp, md);
  for (int i = 0; i < n; ++i) pb.A[(r0 + 2 * n + 1) * nx + i] = md.grad[i];
  pb.A[(r0 + 2 * n + 1) * nx + 3 * n + 1] = 1.0;
  pb.l[r0 + 2 * n + 1] = -cp.alpha * (md.d - cp.dist_thresh);
}

// Manipulator::QPID  x = [qddot; tau; s_qmin; s_qmax; s_vmin; s_vmax; s_sing; s_col]
inline void build_qpid(const Model& m, const State& s, int frame, const double* xddot_des, const CtrlParams& cp,
                       const GeomParams& gp, QpProblem& pb) {
  const int n = m.nv, nx = 6 * n + 2, nineq = 4 * n + 2, neq = n, nc = nx + nineq + neq;
  const double a = cp.alpha;
  pb.resize(nx, nc);
  std::vector<double> J(6 * n), Jd(6 * n);
  frame_jacobian(m, s.oMi, s.J, frame, J.data());
  frame_jacobian_time_variation(m, s.oMi, s.ov, s.J, s.dJ, frame, Jd.data());
  double rhs[6];
  for (int r = 0; r < 6; ++r) {
    double v = 0;
    for (int j = 0; j < n; ++j) v += Jd[r * n + j] * s.qd[j];
    rhs[r] = xddot_des[r] - v;
  }
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) {
      double v = 0;
      for (int r = 0; r < 6; ++r) v += J[r * n + i] * J[r * n + j];
      pb.P[i * nx + j] = 2.0 * v;
    }
    double v = 0;
    for (int r = 0; r < 6; ++r) v += J[r * n + i] * rhs[r];
    pb.q[i] = -2.0 * v;
  }
  for (int i = 2 * n; i < nx; ++i) pb.q[i] = cp.slack_weight;
  for (int i = 0; i < nx; ++i) pb.A[i * nx + i] = 1.0;
  for (int i = 2 * n; i < nx; ++i) pb.l[i] = 0.0;
  const int r0 = nx;
  for (int i = 0; i < n; ++i) {
    // joint angle limits (2nd-order CBF)
    pb.A[(r0 + i) * nx + i] = 1.0; pb.A[(r0 + i) * nx + 2 * n + i] = 1.0;
    pb.l[r0 + i] = -(a + a) * s.qd[i] - a * a * (s.q[i] - m.q_lo[i]);
    pb.A[(r0 + n + i) * nx + i] = -1.0; pb.A[(r0 + n + i) * nx + 3 * n + i] = 1.0;
    pb.l[r0 + n + i] = +(a + a) * s.qd[i] - a * a * (m.q_hi[i] - s.q[i]);
    // joint velocity limits (CBF)
    pb.A[(r0 + 2 * n + i) * nx + i] = 1.0; pb.A[(r0 + 2 * n + i) * nx + 4 * n + i] = 1.0;
    pb.l[r0 + 2 * n + i] = -a * (s.qd[i] + m.v_lim[i]);
    pb.A[(r0 + 3 * n + i) * nx + i] = -1.0; pb.A[(r0 + 3 * n + i) * nx + 5 * n + i] = 1.0;
    pb.l[r0 + 3 * n + i] = -a * (m.v_lim[i] - s.qd[i]);
  }
  ManipResult mr;
  manipulability(m, s, frame, true, true, 0, n, mr);
  double gd = 0, gq = 0;
  for (int i = 0; i < n; ++i) { pb.A[(r0 + 4 * n) * nx + i] = mr.grad[i]; gd += mr.grad_dot[i] * s.qd[i]; gq += mr.grad[i] * s.qd[i]; }
  pb.A[(r0 + 4 * n) * nx + 6 * n] = 1.0;
  pb.l[r0 + 4 * n] = -gd - (a + a) * gq - a * a * (mr.m - cp.mani_thresh);
  MinDistResult md;
  min_distance(m, s, true, true, gp, md);
  gd = 0; gq = 0;
  for (int i = 0; i < n; ++i) { pb.A[(r0 + 4 * n + 1) * nx + i] = md.grad[i]; gd += md.grad_dot[i] * s.qd[i]; gq += md.grad[i] * s.qd[i]; }
  pb.A[(r0 + 4 * n + 1) * nx + 6 * n + 1] = 1.0;
  pb.l[r0 + 4 * n + 1] = -gd - (a + a) * gq - a * a * (md.d - cp.dist_thresh);
  // equality  [M -I][qddot; tau] = -g   (gravity only)
  const int e0 = r0 + nineq;
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) pb.A[(e0 + i) * nx + j] = s.M[i * n + j];
    pb.A[(e0 + i) * nx + n + i] = -1.0;
    pb.l[e0 + i] = -s.g[i];
    pb.u[e0 + i] = -s.g[i];
  }
}

// ------------------------------------------------------------------ controllers (manipulator)
struct Workspace {
  State s;
  QpProblem pb;
  QpWork w;
  QpResult res;
};

inline void desired_from_error(const Model& m, const State& s, int frame, const SE3& x_target, const double* xdot_target,
                               const CtrlParams& cp, bool use_kv, double* out) {
  const int n = m.nv;
  SE3 x = frame_pose(m, s.oMi, frame);
  std::vector<double> J(6 * n);
  frame_jacobian(m, s.oMi, s.J, frame, J.data());
  double xdot[6], xe[6], xde[6];
  matvec(J.data(), s.qd, xdot, 6, n);
  task_space_error(x_target, xdot_target, x, xdot, xe, xde);
  for (int i = 0; i < 6; ++i) out[i] = cp.Kp_task[i] * xe[i] + (use_kv ? cp.Kv_task[i] * xde[i] : xdot_target[i]);
}

// RobotController::QPIK (robot_controller.cpp:277-290): zeros on failure
inline int ctrl_qpik(const Model& m, Workspace& ws, int frame, const double* xdot_des, const CtrlParams& cp,
                     const GeomParams& gp, const QpSettings& st, double* qdot_out) {
  build_qpik(m, ws.s, frame, xdot_des, cp, gp, ws.pb);
  qp_solve(ws.pb, st, ws.res, ws.w);
  for (int i = 0; i < m.nv; ++i) qdot_out[i] = ws.res.status == QP_SOLVED ? ws.res.x[i] : 0.0;
  return ws.res.status;
}
// RobotController::QPID (robot_controller.cpp:319-333): gravity on failure
inline int ctrl_qpid(const Model& m, Workspace& ws, int frame, const double* xddot_des, const CtrlParams& cp,
                     const GeomParams& gp, const QpSettings& st, double* tau_out, double* qddot_out) {
  build_qpid(m, ws.s, frame, xddot_des, cp, gp, ws.pb);
  qp_solve(ws.pb, st, ws.res, ws.w);
  const int n = m.nv;
  for (int i = 0; i < n; ++i) {
    tau_out[i] = ws.res.status == QP_SOLVED ? ws.res.x[n + i] : ws.s.g[i];
    if (qddot_out) qddot_out[i] = ws.res.status == QP_SOLVED ? ws.res.x[i] : 0.0;
  }
  return ws.res.status;
}
// CLIKStep (robot_controller.cpp:156-171)
inline void ctrl_clik_step(const Model& m, const State& s, int frame, const SE3& x_target, const double* xdot_target,
                           const double* null_qdot, const CtrlParams& cp, double* qdot_out) {
  const int n = m.nv;
  double des[6];
  desired_from_error(m, s, frame, x_target, xdot_target, cp, false, des);
  std::vector<double> J(6 * n), Jp(n * 6), N(n * n);
  frame_jacobian(m, s.oMi, s.J, frame, J.data());
  pinv_cod(J.data(), 6, n, Jp.data());
  matmul(Jp.data(), J.data(), N.data(), n, 6, n);
  for (int i = 0; i < n; ++i) {
    double v = 0;
    for (int r = 0; r < 6; ++r) v += Jp[i * 6 + r] * des[r];
    for (int j = 0; j < n; ++j) v += ((i == j ? 1.0 : 0.0) - N[i * n + j]) * (null_qdot ? null_qdot[j] : 0.0);
    qdot_out[i] = v;
  }
}
// OSF (robot_controller.cpp:208-225)
inline void ctrl_osf(const Model& m, const State& s, int frame, const double* xddot, const double* null_tau, double* tau) {
  const int n = m.nv;
  std::vector<double> J(6 * n), JMi(6 * n), L(36), Li(36), JTp(6 * n);
  frame_jacobian(m, s.oMi, s.J, frame, J.data());
  matmul(J.data(), s.Minv, JMi.data(), 6, n, n);
  matmul_nt(JMi.data(), J.data(), Li.data(), 6, n, 6);
  pinv_cod(Li.data(), 6, 6, L.data());
  matmul(L.data(), JMi.data(), JTp.data(), 6, 6, n);  // J_T_pinv = M_task J Minv (6 x n)
  double F[6];
  matvec(L.data(), xddot, F, 6, 6);
  for (int i = 0; i < n; ++i) {
    double v = 0;
    for (int r = 0; r < 6; ++r) v += J[r * n + i] * F[r];
    if (null_tau)
      for (int j = 0; j < n; ++j) {
        double pij = (i == j ? 1.0 : 0.0);
        for (int r = 0; r < 6; ++r) pij -= J[r * n + i] * JTp[r * n + j];
        v += pij * null_tau[j];
      }
    tau[i] = v + s.g[i];
  }
}
// moveJointTorqueStep(q_target, qdot_target) (robot_controller.cpp:115-125)
inline void ctrl_joint_torque_step(const Model& m, const State& s, const double* q_t, const double* qd_t,
                                   const CtrlParams& cp, double* tau) {
  const int n = m.nv;
  double acc[MAXV];
  for (int i = 0; i < n; ++i) acc[i] = cp.Kp_joint[i] * (q_t[i] - s.q[i]) + cp.Kv_joint[i] * (qd_t[i] - s.qd[i]);
  for (int i = 0; i < n; ++i) {
    double v = s.g[i];
    for (int j = 0; j < n; ++j) v += s.M[i * n + j] * acc[j];
    tau[i] = v;
  }
}

}  // namespace orc
