// ORACLE -- test infrastructure only. PARITY UNPINNED (no reference goldens exist).
// Rigid-body kinematics/dynamics restated from the published algorithms the reference calls
// through Pinocchio (un-vendored, un-pinned: package.xml:34):
//   computeJointJacobians / computeJointJacobiansTimeVariation   <- robot_data.cpp:103-104
//   crba / computeGeneralizedGravity / nonLinearEffects          <- robot_data.cpp:111-113
//   getFrameJacobian / getFrameJacobianTimeVariation (LOCAL_WORLD_ALIGNED) <- robot_data.cpp:392-417
// Conventions (Featherstone RBDA as used by Pinocchio): spatial vectors are (linear; angular),
// joint i's frame oMi = oM_parent * placement_i * Xjoint(q_i), gravity (0,0,-9.81).
#pragma once
#include "omath.h"

namespace orc {

constexpr int MAXV = 16;
enum { JOINT_REVOLUTE = 0, JOINT_PRISMATIC = 1 };
enum { GEOM_SPHERE = 0, GEOM_CYLINDER = 1, GEOM_BOX = 2, GEOM_CAPSULE = 3, GEOM_CONVEX = 4 };

struct Motion { V3 lin, ang; };
struct Force { V3 lin, ang; };

inline Motion operator+(const Motion& a, const Motion& b) { return {a.lin + b.lin, a.ang + b.ang}; }
inline Force operator+(const Force& a, const Force& b) { return {a.lin + b.lin, a.ang + b.ang}; }
// M.act(v): motion expressed in frame B -> frame A, with M = aMb
inline Motion act(const SE3& M, const Motion& v) {
  V3 w = M.R * v.ang;
  return {M.R * v.lin + cross(M.p, w), w};
}
inline Motion act_inv(const SE3& M, const Motion& v) {
  return {tmul(M.R, v.lin - cross(M.p, v.ang)), tmul(M.R, v.ang)};
}
inline Force act(const SE3& M, const Force& f) {
  V3 l = M.R * f.lin;
  return {l, M.R * f.ang + cross(M.p, l)};
}
// motion x motion
inline Motion cross_mm(const Motion& a, const Motion& b) {
  return {cross(a.ang, b.lin) + cross(a.lin, b.ang), cross(a.ang, b.ang)};
}
// motion x* force
inline Force cross_mf(const Motion& v, const Force& f) {
  return {cross(v.ang, f.lin), cross(v.ang, f.ang) + cross(v.lin, f.lin)};
}

// Spatial inertia (mass, centre of mass, rotational inertia about the centre of mass).
struct Inertia {
  double m = 0;
  V3 c;
  M3 I;
};
inline Force mul(const Inertia& Y, const Motion& v) {
  V3 fl = Y.m * (v.lin - cross(Y.c, v.ang));
  return {fl, Y.I * v.ang + cross(Y.c, fl)};
}
inline Inertia act(const SE3& M, const Inertia& Y) {
  Inertia r;
  r.m = Y.m;
  r.c = M.R * Y.c + M.p;
  r.I = M.R * Y.I * transpose(M.R);
  return r;
}
inline Inertia add(const Inertia& a, const Inertia& b) {
  Inertia r;
  r.m = a.m + b.m;
  if (r.m <= 0) return r;
  r.c = (1.0 / r.m) * (a.m * a.c + b.m * b.c);
  auto shift = [&](const Inertia& y) {
    V3 d = y.c - r.c;
    return y.I + y.m * (dot(d, d) * M3::identity() - outer(d, d));
  };
  r.I = shift(a) + shift(b);
  return r;
}

struct Model {
  int nv = 0;
  int parent[MAXV];
  int jtype[MAXV];
  V3 axis[MAXV];
  SE3 jplace[MAXV];
  Inertia inertia[MAXV];
  double q_lo[MAXV], q_hi[MAXV], v_lim[MAXV];
  bool anc[MAXV][MAXV];  // anc[i][j]: joint j is i or an ancestor of i
  V3 gravity;
  // frames
  int nf = 0;
  std::vector<int> frame_parent;
  std::vector<SE3> frame_place;
  // collision geometry
  int ng = 0;
  std::vector<int> geom_type, geom_parent;
  std::vector<V3> geom_param;
  std::vector<SE3> geom_place;
  std::vector<double> hull;             // GEOM_CONVEX: hull vertices of mesh geometry (xyz triples, geometry frame)
  std::vector<int> hull_off, hull_n;   // per geometry: first vertex / vertex count
  std::vector<int> pair_a, pair_b;
  // mobile-manipulator extension (filled by orc_model_set_moma)
  int drive_type = -1;  // 0 differential, 1 mecanum, 2 caster
  int wheel_num = 0, virtual_start = 0, mani_start = 0, mobi_start = 0, act_mani_start = 0, act_mobi_start = 0;
  double wheel_radius = 0, base_width = 0, wheel_offset = 0;
  std::vector<double> roller_angles, b2w_x, b2w_y, b2w_ang;
};

inline Motion joint_subspace(const Model& m, int i) {
  return m.jtype[i] == JOINT_REVOLUTE ? Motion{V3(), m.axis[i]} : Motion{m.axis[i], V3()};
}

struct State {
  double q[MAXV], qd[MAXV];
  SE3 liMi[MAXV], oMi[MAXV];
  double J[6 * MAXV], dJ[6 * MAXV];  // 6 x nv, row-major, world frame (Pinocchio data.J / data.dJ)
  Motion v[MAXV], ov[MAXV];
  double M[MAXV * MAXV], Minv[MAXV * MAXV], g[MAXV], nle[MAXV], c[MAXV];
};

inline void set_col(double* A, int nv, int j, const Motion& mo) {
  A[0 * nv + j] = mo.lin.x; A[1 * nv + j] = mo.lin.y; A[2 * nv + j] = mo.lin.z;
  A[3 * nv + j] = mo.ang.x; A[4 * nv + j] = mo.ang.y; A[5 * nv + j] = mo.ang.z;
}
inline Motion get_col(const double* A, int nv, int j) {
  return {V3(A[0 * nv + j], A[1 * nv + j], A[2 * nv + j]), V3(A[3 * nv + j], A[4 * nv + j], A[5 * nv + j])};
}

// pinocchio::forwardKinematics (positions) + computeJointJacobians
inline void forward_kinematics(const Model& m, const double* q, SE3* liMi, SE3* oMi) {
  for (int i = 0; i < m.nv; ++i) {
    SE3 X;
    if (m.jtype[i] == JOINT_REVOLUTE) X.R = axis_angle(m.axis[i], q[i]);
    else X.p = q[i] * m.axis[i];
    liMi[i] = m.jplace[i] * X;
    oMi[i] = m.parent[i] < 0 ? liMi[i] : oMi[m.parent[i]] * liMi[i];
  }
}
inline void joint_jacobians(const Model& m, const SE3* oMi, double* J) {
  for (int i = 0; i < m.nv; ++i) set_col(J, m.nv, i, act(oMi[i], joint_subspace(m, i)));
}
// pinocchio::computeJointJacobiansTimeVariation: dJ[:,i] = ov_i x J[:,i]
inline void joint_jacobians_time_variation(const Model& m, const double* q, const double* qd, SE3* liMi, SE3* oMi,
                                           Motion* v, Motion* ov, double* J, double* dJ) {
  forward_kinematics(m, q, liMi, oMi);
  for (int i = 0; i < m.nv; ++i) {
    Motion S = joint_subspace(m, i);
    Motion vj{qd[i] * S.lin, qd[i] * S.ang};
    v[i] = m.parent[i] < 0 ? vj : act_inv(liMi[i], v[m.parent[i]]) + vj;
    ov[i] = act(oMi[i], v[i]);
    Motion Jc = act(oMi[i], S);
    set_col(J, m.nv, i, Jc);
    set_col(dJ, m.nv, i, cross_mm(ov[i], Jc));
  }
}

// pinocchio::crba (upper triangle) + symmetrisation (reference robot_data.cpp:116-117)
inline void crba(const Model& m, const SE3* liMi, double* M) {
  const int n = m.nv;
  Inertia Y[MAXV];
  Force F[MAXV];  // F[j]: column j expressed in the frame of the joint currently being processed
  std::fill(M, M + n * n, 0.0);
  for (int i = 0; i < n; ++i) Y[i] = m.inertia[i];
  for (int i = n - 1; i >= 0; --i) {
    Motion S = joint_subspace(m, i);
    F[i] = mul(Y[i], S);
    for (int j = i; j < n; ++j)
      if (m.anc[j][i]) M[i * n + j] = dot(S.lin, F[j].lin) + dot(S.ang, F[j].ang);
    if (m.parent[i] >= 0) {
      Y[m.parent[i]] = add(Y[m.parent[i]], act(liMi[i], Y[i]));
      for (int j = i; j < n; ++j)
        if (m.anc[j][i]) F[j] = act(liMi[i], F[j]);
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) M[j * n + i] = M[i * n + j];
}

// pinocchio::rnea (used as computeGeneralizedGravity with qd=qdd=0 and nonLinearEffects with qdd=0)
inline void rnea(const Model& m, const SE3* liMi, const double* qd, const double* qdd, double* tau) {
  const int n = m.nv;
  Motion v[MAXV], a[MAXV];
  Force f[MAXV];
  const Motion a0{-1.0 * m.gravity, V3()};
  for (int i = 0; i < n; ++i) {
    Motion S = joint_subspace(m, i);
    Motion vj{(qd ? qd[i] : 0.0) * S.lin, (qd ? qd[i] : 0.0) * S.ang};
    Motion aj{(qdd ? qdd[i] : 0.0) * S.lin, (qdd ? qdd[i] : 0.0) * S.ang};
    if (m.parent[i] < 0) {
      v[i] = vj;
      a[i] = act_inv(liMi[i], a0) + aj;
    } else {
      v[i] = act_inv(liMi[i], v[m.parent[i]]) + vj;
      a[i] = act_inv(liMi[i], a[m.parent[i]]) + aj + cross_mm(v[i], vj);
    }
    f[i] = mul(m.inertia[i], a[i]) + cross_mf(v[i], mul(m.inertia[i], v[i]));
  }
  for (int i = n - 1; i >= 0; --i) {
    Motion S = joint_subspace(m, i);
    tau[i] = dot(S.lin, f[i].lin) + dot(S.ang, f[i].ang);
    if (m.parent[i] >= 0) f[m.parent[i]] = f[m.parent[i]] + act(liMi[i], f[i]);
  }
}

// reference Manipulator::RobotData::updateState (robot_data.cpp:91-124)
inline void update_state(const Model& m, State& s, const double* q, const double* qd) {
  const int n = m.nv;
  for (int i = 0; i < n; ++i) { s.q[i] = q[i]; s.qd[i] = qd[i]; }
  // updateKinematics: computeJointJacobians then computeJointJacobiansTimeVariation (which redoes FK)
  forward_kinematics(m, q, s.liMi, s.oMi);
  joint_jacobians(m, s.oMi, s.J);
  joint_jacobians_time_variation(m, q, qd, s.liMi, s.oMi, s.v, s.ov, s.J, s.dJ);
  // updateDynamics
  crba(m, s.liMi, s.M);
  rnea(m, s.liMi, nullptr, nullptr, s.g);
  rnea(m, s.liMi, qd, nullptr, s.nle);
  pinv_cod(s.M, n, n, s.Minv);
  for (int i = 0; i < n; ++i) s.c[i] = s.nle[i] - s.g[i];
}

// Frame pose: oMi[parent] * placement (the "fresh" pose).
inline SE3 frame_pose(const Model& m, const SE3* oMi, int frame) {
  int pj = m.frame_parent[frame];
  return pj < 0 ? m.frame_place[frame] : oMi[pj] * m.frame_place[frame];
}

// pinocchio::getFrameJacobian(LOCAL_WORLD_ALIGNED): columns on the support of the parent joint,
// linear part translated to the frame origin.
inline void translate_lwa(const Model& m, int joint, const V3& p, const double* Jw, double* out) {
  const int n = m.nv;
  std::fill(out, out + 6 * n, 0.0);
  if (joint < 0) return;
  for (int j = 0; j < n; ++j) {
    if (!m.anc[joint][j]) continue;
    Motion c = get_col(Jw, n, j);
    set_col(out, n, j, Motion{c.lin - cross(p, c.ang), c.ang});
  }
}
inline void frame_jacobian(const Model& m, const SE3* oMi, const double* Jw, int frame, double* Jf) {
  SE3 oMf = frame_pose(m, oMi, frame);
  translate_lwa(m, m.frame_parent[frame], oMf.p, Jw, Jf);
}
// pinocchio (3.x) getFrameJacobianTimeVariation(LOCAL_WORLD_ALIGNED): translated dJ minus
// (velocity of the frame origin) x (angular part of J)  -> exact d/dt of the LWA Jacobian.
inline void lwa_time_variation(const Model& m, int joint, const V3& p, const Motion* ov, const double* Jw,
                               const double* dJw, double* out) {
  const int n = m.nv;
  translate_lwa(m, joint, p, dJw, out);
  if (joint < 0) return;
  V3 vp = ov[joint].lin + cross(ov[joint].ang, p);
  for (int j = 0; j < n; ++j) {
    if (!m.anc[joint][j]) continue;
    Motion c = get_col(Jw, n, j);
    V3 corr = cross(vp, c.ang);
    out[0 * n + j] -= corr.x; out[1 * n + j] -= corr.y; out[2 * n + j] -= corr.z;
  }
}
inline void frame_jacobian_time_variation(const Model& m, const SE3* oMi, const Motion* ov, const double* Jw,
                                          const double* dJw, int frame, double* dJf) {
  SE3 oMf = frame_pose(m, oMi, frame);
  lwa_time_variation(m, m.frame_parent[frame], oMf.p, ov, Jw, dJw, dJf);
}

// reference getManipulability (robot_data.cpp:519-573).  col0/ncols select the manipulator block
// for the mobile-manipulator override (mobile_manipulator/robot_data.cpp:439-496); the dJ/dq_i
// matrices are produced the way the reference does it: a J-time-variation pass with qdot = e_i.
struct ManipResult { double m; double grad[MAXV]; double grad_dot[MAXV]; };
inline void manipulability(const Model& mdl, const State& s, int frame, bool with_grad, bool with_graddot,
                           int col0, int ncols, ManipResult& out) {
  const int n = mdl.nv, k = ncols;
  out.m = 0;
  for (int i = 0; i < MAXV; ++i) out.grad[i] = out.grad_dot[i] = 0;
  std::vector<double> Jf(6 * n), J(6 * k), JJt(36), JJt_inv(36);
  frame_jacobian(mdl, s.oMi, s.J, frame, Jf.data());
  for (int r = 0; r < 6; ++r) for (int c = 0; c < k; ++c) J[r * k + c] = Jf[r * n + col0 + c];
  matmul_nt(J.data(), J.data(), JJt.data(), 6, k, 6);
  out.m = std::sqrt(determinant(JJt.data(), 6));
  if (!with_grad && !with_graddot) return;
  pinv_cod(JJt.data(), 6, 6, JJt_inv.data());
  // G = J^T (JJt)^-1   (k x 6)
  std::vector<double> G(k * 6);
  matmul_tn(J.data(), JJt_inv.data(), G.data(), 6, k, 6);
  std::vector<std::vector<double>> dJdq(k, std::vector<double>(6 * k));
  State tmp;  // fresh pinocchio::Data (robot_data.cpp:542)
  std::vector<double> e(n), dJf(6 * n);
  for (int i = 0; i < k; ++i) {
    std::fill(e.begin(), e.end(), 0.0);
    e[col0 + i] = 1.0;
    joint_jacobians_time_variation(mdl, s.q, e.data(), tmp.liMi, tmp.oMi, tmp.v, tmp.ov, tmp.J, tmp.dJ);
    frame_jacobian_time_variation(mdl, tmp.oMi, tmp.ov, tmp.J, tmp.dJ, frame, dJf.data());
    for (int r = 0; r < 6; ++r) for (int c = 0; c < k; ++c) dJdq[i][r * k + c] = dJf[r * n + col0 + c];
    double tr = 0;  // trace(dJdq_i * G)
    for (int r = 0; r < 6; ++r) for (int c = 0; c < k; ++c) tr += dJdq[i][r * k + c] * G[c * 6 + r];
    out.grad[i] = out.m * tr;
  }
  if (!with_graddot) return;
  std::vector<double> Jdf(6 * n), Jd(6 * k);
  frame_jacobian_time_variation(mdl, s.oMi, s.ov, s.J, s.dJ, frame, Jdf.data());
  for (int r = 0; r < 6; ++r) for (int c = 0; c < k; ++c) Jd[r * k + c] = Jdf[r * n + col0 + c];
  double trd = 0;
  for (int r = 0; r < 6; ++r) for (int c = 0; c < k; ++c) trd += Jd[r * k + c] * G[c * 6 + r];
  const double mani_dot = out.m * trd;
  // JJt_dot = 2 Jdot J^T (sic, robot_data.cpp:561); JJt_inv_dot = -(JJt_inv JJt_dot JJt_inv)
  std::vector<double> JJt_dot(36), t1(36), JJt_inv_dot(36);
  matmul_nt(Jd.data(), J.data(), JJt_dot.data(), 6, k, 6);
  for (double& x : JJt_dot) x *= 2.0;
  matmul(JJt_inv.data(), JJt_dot.data(), t1.data(), 6, 6, 6);
  matmul(t1.data(), JJt_inv.data(), JJt_inv_dot.data(), 6, 6, 6);
  for (double& x : JJt_inv_dot) x = -x;
  // H = Jdot^T JJt_inv + J^T JJt_inv_dot   (k x 6)
  std::vector<double> H1(k * 6), H2(k * 6);
  matmul_tn(Jd.data(), JJt_inv.data(), H1.data(), 6, k, 6);
  matmul_tn(J.data(), JJt_inv_dot.data(), H2.data(), 6, k, 6);
  for (int i = 0; i < k; ++i) {
    double tr1 = 0, tr2 = 0;
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < k; ++c) {
        tr1 += dJdq[i][r * k + c] * G[c * 6 + r];
        tr2 += dJdq[i][r * k + c] * (H1[c * 6 + r] + H2[c * 6 + r]);
      }
    out.grad_dot[i] = mani_dot * tr1 + out.m * tr2;
  }
}

}  // namespace orc
