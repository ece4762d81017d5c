// ORACLE -- test infrastructure only. PARITY UNPINNED (no reference goldens exist).
// Mobile base and mobile-manipulator path restated from the reference:
//   Mobile::RobotData            src/mobile/robot_data.cpp:104-204            (base FK Jacobians, base velocity)
//   MobileManipulator::RobotData src/mobile_manipulator/robot_data.cpp:7-144,407-496 (selection matrix S, actuated dynamics,
//                                                                                 actuated Jacobians, manipulability override)
//   MobileManipulator::QPIK      src/mobile_manipulator/QP_IK.cpp:59-128      (hard CBF rows, 0.01 I regulariser, free bounds)
//   MobileManipulator::QPID      src/mobile_manipulator/QP_ID.cpp:75-184      (no bound rows, M~ eta_dot - tau = -g~)
//   MobileManipulator::RobotController  src/mobile_manipulator/robot_controller.cpp:147-250
#pragma once
#include "octrl.h"

namespace orc {

// Mobile::RobotData::computeFKJacobian (mobile/robot_data.cpp:122-204): 3 x w, row-major
inline void mobile_fk_jacobian(const Model& m, const double* wheel_pos, double* J) {
  const int w = m.wheel_num;
  std::fill(J, J + 3 * w, 0.0);
  if (m.drive_type == 0) {                     // DifferentialFKJacobian :138-147
    J[0 * w + 0] = m.wheel_radius / 2; J[0 * w + 1] = m.wheel_radius / 2;
    J[2 * w + 0] = -m.wheel_radius / m.base_width; J[2 * w + 1] = m.wheel_radius / m.base_width;
  } else if (m.drive_type == 1) {              // MecanumFKJacobian :149-177 = PinvCOD(J_inv)
    std::vector<double> Ji(w * 3);
    for (int i = 0; i < w; ++i) {
      const double r = m.wheel_radius, g = m.roller_angles[i], px = m.b2w_x[i], py = m.b2w_y[i], pt = m.b2w_ang[i];
      const double A1[2][3] = {{1, 0, -py}, {0, 1, px}};
      const double A2[2][2] = {{std::cos(pt), std::sin(pt)}, {-std::sin(pt), std::cos(pt)}};
      const double A3[2] = {1.0, std::tan(g)};
      for (int c = 0; c < 3; ++c) {
        double v = 0;
        for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) v += A3[a] * A2[a][b] * A1[b][c];
        Ji[i * 3 + c] = v / r;
      }
    }
    pinv_cod(Ji.data(), w, 3, J);
  } else {                                     // CasterFKJacobian :179-203 (Holmberg & Khatib)
    const int ns = w / 2;
    std::vector<double> Jp(w * 3, 0.0), Jq(w * w, 0.0), JtJ(9), JtJi(9), T(3 * w);
    for (int i = 0; i < ns; ++i) {
      const double r = m.wheel_radius, b = m.wheel_offset, px = m.b2w_x[i], py = m.b2w_y[i], phi = wheel_pos[2 * i];
      Jp[(2 * i) * 3 + 0] = 1; Jp[(2 * i) * 3 + 2] = -(py + b * std::sin(phi));
      Jp[(2 * i + 1) * 3 + 1] = 1; Jp[(2 * i + 1) * 3 + 2] = px + b * std::cos(phi);
      Jq[(2 * i) * w + 2 * i] = b * std::sin(phi); Jq[(2 * i) * w + 2 * i + 1] = r * std::cos(phi);
      Jq[(2 * i + 1) * w + 2 * i] = -b * std::cos(phi); Jq[(2 * i + 1) * w + 2 * i + 1] = r * std::sin(phi);
    }
    matmul_tn(Jp.data(), Jp.data(), JtJ.data(), w, 3, 3);
    pinv_cod(JtJ.data(), 3, 3, JtJi.data());
    matmul_tn(Jp.data(), Jq.data(), T.data(), w, 3, w);   // Jp^T Jq  (3 x w)
    matmul(JtJi.data(), T.data(), J, 3, 3, w);
  }
}

// Mobile::RobotController::computeIKJacobian (mobile/robot_controller.cpp:50-124): w x 3, row-major
inline void mobile_ik_jacobian(const Model& m, const double* wheel_pos, double* Ji) {
  const int w = m.wheel_num;
  std::fill(Ji, Ji + 3 * w, 0.0);
  if (m.drive_type == 0) {                     // DifferentialIKJacobian :65-74
    Ji[0] = 1 / m.wheel_radius; Ji[2] = -m.base_width / (2 * m.wheel_radius);
    Ji[3] = 1 / m.wheel_radius; Ji[5] = m.base_width / (2 * m.wheel_radius);
  } else if (m.drive_type == 1) {              // MecanumIKJacobian :76-102
    for (int i = 0; i < w; ++i) {
      const double r = m.wheel_radius, g = m.roller_angles[i], px = m.b2w_x[i], py = m.b2w_y[i], pt = m.b2w_ang[i];
      const double A1[2][3] = {{1, 0, -py}, {0, 1, px}};
      const double A2[2][2] = {{std::cos(pt), std::sin(pt)}, {-std::sin(pt), std::cos(pt)}};
      const double A3[2] = {1.0, std::tan(g)};
      for (int c = 0; c < 3; ++c) {
        double v = 0;
        for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) v += A3[a] * A2[a][b] * A1[b][c];
        Ji[i * 3 + c] = v / r;
      }
    }
  } else {                                     // CasterIKJacobian :104-123
    for (int i = 0; i < w / 2; ++i) {
      const double r = m.wheel_radius, b = m.wheel_offset, px = m.b2w_x[i], py = m.b2w_y[i], phi = wheel_pos[2 * i];
      Ji[(2 * i) * 3 + 0] = -std::sin(phi) / b; Ji[(2 * i) * 3 + 1] = std::cos(phi) / b;
      Ji[(2 * i) * 3 + 2] = (px * std::cos(phi) + py * std::sin(phi)) / b - 1;
      Ji[(2 * i + 1) * 3 + 0] = std::cos(phi) / r; Ji[(2 * i + 1) * 3 + 1] = std::sin(phi) / r;
      Ji[(2 * i + 1) * 3 + 2] = (px * std::sin(phi) - py * std::cos(phi)) / r;
    }
  }
}

// Mobile::RobotController::VelocityCommand (:14-41) when saturate, computeWheelVel (:43-47) otherwise
inline void mobile_wheel_velocity(const Model& m, double max_lin_speed, double max_ang_speed, const double* wheel_pos,
                                  const double* base_vel, bool saturate, double* wheel_vel) {
  double v[3] = {base_vel[0], base_vel[1], base_vel[2]};
  if (saturate) {
    double speed = std::sqrt(v[0] * v[0] + v[1] * v[1]);
    double dir[2] = {0, 0};
    if (!(std::fabs(speed) < 1e-4)) { dir[0] = v[0] / speed; dir[1] = v[1] / speed; }
    speed = std::min(std::max(speed, -max_lin_speed), max_lin_speed);
    v[0] = dir[0] * speed; v[1] = dir[1] * speed;
    v[2] = std::min(std::max(v[2], -max_ang_speed), max_ang_speed);
  }
  std::vector<double> Ji(3 * m.wheel_num);
  mobile_ik_jacobian(m, wheel_pos, Ji.data());
  for (int k = 0; k < m.wheel_num; ++k) wheel_vel[k] = Ji[k * 3] * v[0] + Ji[k * 3 + 1] * v[1] + Ji[k * 3 + 2] * v[2];
}

struct MomaState {
  int act = 0, mani = 0;
  double S[MAXV * MAXV];                          // dof x act
  double M[MAXV * MAXV], Minv[MAXV * MAXV];       // act x act
  double g[MAXV], nle[MAXV], c[MAXV];
  double q_act[MAXV], qd_act[MAXV];
  double J_mobile[3 * 8], base_vel[3];
};

// MobileManipulator::RobotData::updateKinematics / updateDynamics (mobile_manipulator/robot_data.cpp:104-144)
inline void moma_update(const Model& m, const State& s, MomaState& ms) {
  const int n = m.nv, w = m.wheel_num, mani = n - 3 - w, act = w + mani;
  ms.act = act; ms.mani = mani;
  std::fill(ms.S, ms.S + n * act, 0.0);
  for (int i = 0; i < mani; ++i) ms.S[(m.mani_start + i) * act + m.act_mani_start + i] = 1.0;
  for (int i = 0; i < w; ++i) ms.S[(m.mobi_start + i) * act + m.act_mobi_start + i] = 1.0;
  mobile_fk_jacobian(m, s.q + m.mobi_start, ms.J_mobile);
  for (int r = 0; r < 3; ++r) {
    double v = 0;
    for (int k = 0; k < w; ++k) v += ms.J_mobile[r * w + k] * s.qd[m.mobi_start + k];
    ms.base_vel[r] = v;
  }
  const double yaw = s.q[m.virtual_start + 2], cy = std::cos(yaw), sy = std::sin(yaw);
  const double Rz[3][3] = {{cy, -sy, 0}, {sy, cy, 0}, {0, 0, 1}};
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < w; ++k) {
      double v = 0;
      for (int a = 0; a < 3; ++a) v += Rz[r][a] * ms.J_mobile[a * w + k];
      ms.S[(m.virtual_start + r) * act + m.act_mobi_start + k] = v;
    }
  for (int i = 0; i < w; ++i) { ms.q_act[m.act_mobi_start + i] = s.q[m.mobi_start + i]; ms.qd_act[m.act_mobi_start + i] = s.qd[m.mobi_start + i]; }
  for (int i = 0; i < mani; ++i) { ms.q_act[m.act_mani_start + i] = s.q[m.mani_start + i]; ms.qd_act[m.act_mani_start + i] = s.qd[m.mani_start + i]; }
  std::vector<double> MS(n * act);
  matmul(s.M, ms.S, MS.data(), n, n, act);
  matmul_tn(ms.S, MS.data(), ms.M, n, act, act);
  pinv_cod(ms.M, act, act, ms.Minv);
  matvec_t(ms.S, s.g, ms.g, n, act);
  matvec_t(ms.S, s.nle, ms.nle, n, act);
  for (int i = 0; i < act; ++i) ms.c[i] = ms.nle[i] - ms.g[i];
}

// getJacobianActuated / getJacobianActuatedTimeVariation (robot_data.cpp:407-415; S-dot neglected)
inline void moma_jacobians(const Model& m, const State& s, const MomaState& ms, int frame, double* Jt, double* Jtd) {
  const int n = m.nv;
  std::vector<double> J(6 * n), Jd(6 * n);
  frame_jacobian(m, s.oMi, s.J, frame, J.data());
  matmul(J.data(), ms.S, Jt, 6, n, ms.act);
  if (Jtd) {
    frame_jacobian_time_variation(m, s.oMi, s.ov, s.J, s.dJ, frame, Jd.data());
    matmul(Jd.data(), ms.S, Jtd, 6, n, ms.act);
  }
}

// MobileManipulator::QPIK  x = eta (act); A = [I (free bounds); q_min; q_max; sing; col]
inline void build_moma_qpik(const Model& m, const State& s, const MomaState& ms, int frame, const double* xdot_des,
                            const CtrlParams& cp, const GeomParams& gp, QpProblem& pb) {
  const int act = ms.act, k = ms.mani, nx = act, nineq = 2 * k + 2, nc = nx + nineq, am = m.act_mani_start;
  pb.resize(nx, nc);
  std::vector<double> Jt(6 * act);
  moma_jacobians(m, s, ms, frame, Jt.data(), nullptr);
  for (int i = 0; i < act; ++i) {
    for (int j = 0; j < act; ++j) {
      double v = 0;
      for (int r = 0; r < 6; ++r) v += Jt[r * act + i] * Jt[r * act + j];
      pb.P[i * nx + j] = 2.0 * v + (i == j ? cp.moma_ik_reg : 0.0);
    }
    double v = 0;
    for (int r = 0; r < 6; ++r) v += Jt[r * act + i] * xdot_des[r];
    pb.q[i] = -2.0 * v;
    pb.A[i * nx + i] = 1.0;   // bound rows stay (-inf, inf): setBoundConstraint is commented out (QP_IK.cpp:75-83)
  }
  const int r0 = nx;
  for (int i = 0; i < k; ++i) {
    const double qi = s.q[m.mani_start + i];
    pb.A[(r0 + i) * nx + am + i] = 1.0;
    pb.l[r0 + i] = -cp.alpha * (qi - m.q_lo[m.mani_start + i]);
    pb.A[(r0 + k + i) * nx + am + i] = -1.0;
    pb.l[r0 + k + i] = -cp.alpha * (m.q_hi[m.mani_start + i] - qi);
  }
  ManipResult mr;
  manipulability(m, s, frame, true, false, m.mani_start, k, mr);
  for (int i = 0; i < k; ++i) pb.A[(r0 + 2 * k) * nx + am + i] = mr.grad[i];
  pb.l[r0 + 2 * k] = -cp.alpha * (mr.m - cp.mani_thresh);
  MinDistResult md;
  min_distance(m, s, true, false, gp, md);
  for (int i = 0; i < k; ++i) pb.A[(r0 + 2 * k + 1) * nx + am + i] = md.grad[m.mani_start + i];
  pb.l[r0 + 2 * k + 1] = -cp.alpha * (md.d - cp.dist_thresh);
}

// MobileManipulator::QPID  x = [eta_dot (act); tau (act)]; no bound rows; A = [ineq (4k+2); eq (act)]
inline void build_moma_qpid(const Model& m, const State& s, const MomaState& ms, int frame, const double* xddot_des,
                            const CtrlParams& cp, const GeomParams& gp, QpProblem& pb) {
  const int act = ms.act, k = ms.mani, nx = 2 * act, nineq = 4 * k + 2, nc = nineq + act, am = m.act_mani_start;
  const double a = cp.alpha;
  pb.resize(nx, nc);
  std::vector<double> Jt(6 * act), Jtd(6 * act);
  moma_jacobians(m, s, ms, frame, Jt.data(), Jtd.data());
  double rhs[6];
  for (int r = 0; r < 6; ++r) {
    double v = 0;
    for (int j = 0; j < act; ++j) v += Jtd[r * act + j] * ms.qd_act[j];
    rhs[r] = xddot_des[r] - v;
  }
  for (int i = 0; i < act; ++i) {
    for (int j = 0; j < act; ++j) {
      double v = 0;
      for (int r = 0; r < 6; ++r) v += Jt[r * act + i] * Jt[r * act + j];
      pb.P[i * nx + j] = 2.0 * v;
    }
    double v = 0;
    for (int r = 0; r < 6; ++r) v += Jt[r * act + i] * rhs[r];
    pb.q[i] = -2.0 * v;
  }
  for (int i = 0; i < k; ++i) {
    const double qi = s.q[m.mani_start + i], qdi = s.qd[m.mani_start + i];
    const double lo = m.q_lo[m.mani_start + i], hi = m.q_hi[m.mani_start + i], vl = m.v_lim[m.mani_start + i];
    pb.A[i * nx + am + i] = 1.0;            pb.l[i] = -(a + a) * qdi - a * a * (qi - lo);
    pb.A[(k + i) * nx + am + i] = -1.0;     pb.l[k + i] = +(a + a) * qdi - a * a * (hi - qi);
    pb.A[(2 * k + i) * nx + am + i] = 1.0;  pb.l[2 * k + i] = -a * (qdi + vl);
    pb.A[(3 * k + i) * nx + am + i] = -1.0; pb.l[3 * k + i] = -a * (vl - qdi);
  }
  ManipResult mr;
  manipulability(m, s, frame, true, true, m.mani_start, k, mr);
  double gd = 0, gq = 0;
  for (int i = 0; i < k; ++i) {
    const double qdi = s.qd[m.mani_start + i];
    pb.A[(4 * k) * nx + am + i] = mr.grad[i]; gd += mr.grad_dot[i] * qdi; gq += mr.grad[i] * qdi;
  }
  pb.l[4 * k] = -gd - (a + a) * gq - a * a * (mr.m - cp.mani_thresh);
  MinDistResult md;
  min_distance(m, s, true, true, gp, md);
  gd = 0; gq = 0;
  for (int i = 0; i < k; ++i) {
    const double qdi = s.qd[m.mani_start + i];
    pb.A[(4 * k + 1) * nx + am + i] = md.grad[m.mani_start + i];
    gd += md.grad_dot[m.mani_start + i] * qdi; gq += md.grad[m.mani_start + i] * qdi;
  }
  pb.l[4 * k + 1] = -gd - (a + a) * gq - a * a * (md.d - cp.dist_thresh);
  const int e0 = nineq;
  for (int i = 0; i < act; ++i) {
    for (int j = 0; j < act; ++j) pb.A[(e0 + i) * nx + j] = ms.M[i * act + j];
    pb.A[(e0 + i) * nx + act + i] = -1.0;
    pb.l[e0 + i] = -ms.g[i];
    pb.u[e0 + i] = -ms.g[i];
  }
}

// MobileManipulator::RobotController::QPIKStep desired signal: Kp e + xdot_target (no Kv term, robot_controller.cpp:181)
// QPIDStep: Kp e + Kv edot (:225).  The frame velocity is J qdot of the FULL model (getVelocity).
// QPIK (:147-166): zeros on failure.  QPID (:192-213): the reference falls back to the FULL-dof gravity vector sliced with
// actuator indices (a latent indexing bug); this restatement returns the actuated gravity g~ instead (documented deviation).
inline int ctrl_moma_qpik(const Model& m, Workspace& ws, const MomaState& ms, int frame, const double* xdot_des,
                          const CtrlParams& cp, const GeomParams& gp, const QpSettings& st, double* eta_out) {
  build_moma_qpik(m, ws.s, ms, frame, xdot_des, cp, gp, ws.pb);
  qp_solve(ws.pb, st, ws.res, ws.w);
  for (int i = 0; i < ms.act; ++i) eta_out[i] = ws.res.status == QP_SOLVED ? ws.res.x[i] : 0.0;
  return ws.res.status;
}
inline int ctrl_moma_qpid(const Model& m, Workspace& ws, const MomaState& ms, int frame, const double* xddot_des,
                          const CtrlParams& cp, const GeomParams& gp, const QpSettings& st, double* etadot_out, double* tau_out) {
  build_moma_qpid(m, ws.s, ms, frame, xddot_des, cp, gp, ws.pb);
  qp_solve(ws.pb, st, ws.res, ws.w);
  for (int i = 0; i < ms.act; ++i) {
    etadot_out[i] = ws.res.status == QP_SOLVED ? ws.res.x[i] : 0.0;
    tau_out[i] = ws.res.status == QP_SOLVED ? ws.res.x[ms.act + i] : ms.g[i];
  }
  return ws.res.status;
}

}  // namespace orc
