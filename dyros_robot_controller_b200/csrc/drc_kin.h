// drc_b200 -- per-thread rigid-body kinematics and dynamics (one robot per thread, fp64).
//
// B200-native replacement of what the reference's RobotData::updateState obtains from Pinocchio
// (reference src/manipulator/robot_data.cpp:91-124, getters :378-422, getManipulability :519-573).
// Formulation: everything is expressed in ONE frame (the world frame, optionally re-centred at a
// reference joint), so the joint-chain recursion carries no 6x6 transforms:
//   S_i   = (p_i x a_i ; a_i)  (revolute)   |  (a_i ; 0) (prismatic)         world screw axis
//   v_i   = v_parent + S_i qd_i                                              spatial velocity
//   M_ij  = S_i . (Ic_j S_j)   with composite inertias Ic_j summed in place  (CRBA)
//   tau_i = S_i . sum_subtree f_j,  f_j = I_j a_j + v_j x* (I_j v_j)         (RNEA)
// Loops are fully unrolled on the compile-time dof NV; CHAIN=true (parent[i] = i-1) keeps every
// index static so the whole state lives in registers.
#pragma once
#include "drc_math.h"

namespace drc {

template <bool CHAIN>
DRC_HD int parent_of(const DrcModelDev& m, int i) { return CHAIN ? i - 1 : m.parent[i]; }
template <bool CHAIN>
DRC_HD bool is_anc(const DrcModelDev& m, int i, int j) {  // j is i or an ancestor of i
  return CHAIN ? (j <= i) : (((m.anc_mask[i] >> j) & 1u) != 0u);
}

struct Spatial {  // (linear; angular) for motions, (force; moment) for forces
  Vec3 l, a;
};
DRC_HD Spatial operator+(Spatial x, Spatial y) { return Spatial{x.l + y.l, x.a + y.a}; }
DRC_HD double dot6(Spatial x, Spatial y) { return dot(x.l, y.l) + dot(x.a, y.a); }
DRC_HD Spatial cross_motion(Spatial v, Spatial s) { return Spatial{cross(v.a, s.l) + cross(v.l, s.a), cross(v.a, s.a)}; }
DRC_HD Spatial cross_force(Spatial v, Spatial f) { return Spatial{cross(v.a, f.l), cross(v.a, f.a) + cross(v.l, f.l)}; }

// Spatial inertia about the (shifted) world origin: mass, first moment h = m c, rotational inertia Io.
struct WInertia {
  double m;
  Vec3 h;
  double I[6];  // xx xy xz yy yz zz, about the origin
};
DRC_HD Vec3 sym_mul(const double* I, Vec3 v) {
  return Vec3{I[0] * v.x + I[1] * v.y + I[2] * v.z, I[1] * v.x + I[3] * v.y + I[4] * v.z, I[2] * v.x + I[4] * v.y + I[5] * v.z};
}
DRC_HD Spatial apply(const WInertia& Y, Spatial v) {
  return Spatial{Y.m * v.l - cross(Y.h, v.a), sym_mul(Y.I, v.a) + cross(Y.h, v.l)};
}
DRC_HD void accumulate(WInertia& A, const WInertia& B) {
  A.m += B.m;
  A.h = A.h + B.h;
#pragma unroll
  for (int k = 0; k < 6; ++k) A.I[k] += B.I[k];
}
// World inertia of link i from its joint-frame description and the joint's world placement.
DRC_HD WInertia world_inertia(const DrcModelDev& m, int i, const Mat3& R, Vec3 p) {
  WInertia Y;
  Y.m = m.mass[i];
  const Vec3 c = mul(R, v3(m.com[i][0], m.com[i][1], m.com[i][2])) + p;
  Y.h = Y.m * c;
  // Ic_world = R I R^T
  const double* I = m.inertia[i];
  Mat3 Il = {{I[0], I[1], I[2], I[1], I[3], I[4], I[2], I[4], I[5]}};
  Mat3 T = mul(R, Il);
  double W[6];
  // (T R^T)_{rc} = sum_k T[r][k] R[c][k]
  W[0] = T.m[0] * R.m[0] + T.m[1] * R.m[1] + T.m[2] * R.m[2];
  W[1] = T.m[0] * R.m[3] + T.m[1] * R.m[4] + T.m[2] * R.m[5];
  W[2] = T.m[0] * R.m[6] + T.m[1] * R.m[7] + T.m[2] * R.m[8];
  W[3] = T.m[3] * R.m[3] + T.m[4] * R.m[4] + T.m[5] * R.m[5];
  W[4] = T.m[3] * R.m[6] + T.m[4] * R.m[7] + T.m[5] * R.m[8];
  W[5] = T.m[6] * R.m[6] + T.m[7] * R.m[7] + T.m[8] * R.m[8];
  const double cc = dot(c, c);
  Y.I[0] = W[0] + Y.m * (cc - c.x * c.x);
  Y.I[1] = W[1] - Y.m * c.x * c.y;
  Y.I[2] = W[2] - Y.m * c.x * c.z;
  Y.I[3] = W[3] + Y.m * (cc - c.y * c.y);
  Y.I[4] = W[4] - Y.m * c.y * c.z;
  Y.I[5] = W[5] + Y.m * (cc - c.z * c.z);
  return Y;
}

template <int NV>
struct KinState {
  Mat3 R[NV];  // world rotation of joint frame i
  Vec3 p[NV];  // world origin of joint frame i (relative to `origin`)
  Vec3 a[NV];  // world joint axis
  Vec3 origin; // shift applied to all positions (0 for fixed-base arms)
};

// Forward kinematics.  q is a per-thread register array.
template <int NV, bool CHAIN>
DRC_HD void forward_kinematics(const DrcModelDev& m, const double* q, KinState<NV>& k) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const Vec3 ax = v3(m.axis[i][0], m.axis[i][1], m.axis[i][2]);
    Mat3 Rl = mat3_from(m.jR[i]);
    Vec3 pl = v3(m.jp[i][0], m.jp[i][1], m.jp[i][2]);
    if (m.jtype[i] == kRevolute) {
      double s, c;
      sincos(q[i], &s, &c);
      Rl = mul(Rl, rot_axis(ax, s, c));
    } else {
      pl = pl + mul(Rl, q[i] * ax);
    }
    const int pr = parent_of<CHAIN>(m, i);
    if (pr < 0) {
      k.R[i] = Rl;
      k.p[i] = pl;
    } else {
      k.R[i] = mul(k.R[pr], Rl);
      k.p[i] = mul(k.R[pr], pl) + k.p[pr];
    }
    k.a[i] = mul(k.R[i], ax);
  }
}

template <int NV>
DRC_HD Spatial screw(const DrcModelDev& m, const KinState<NV>& k, int i) {
  return m.jtype[i] == kRevolute ? Spatial{cross(k.p[i], k.a[i]), k.a[i]} : Spatial{k.a[i], v3(0, 0, 0)};
}

// Spatial velocity of every joint frame (world coordinates, at the shifted origin).
template <int NV, bool CHAIN>
DRC_HD void joint_velocities(const DrcModelDev& m, const KinState<NV>& k, const double* qd, Spatial* v) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const Spatial S = screw<NV>(m, k, i);
    const int pr = parent_of<CHAIN>(m, i);
    const Spatial vj = Spatial{qd[i] * S.l, qd[i] * S.a};
    v[i] = pr < 0 ? vj : v[pr] + vj;
  }
}

// Frame placement in the world.
template <int NV>
DRC_HD void frame_pose(const KinState<NV>& k, const DrcFrame& f, Mat3& Rf, Vec3& pf) {
  Rf = mat3_from(f.R);
  pf = v3(f.p[0], f.p[1], f.p[2]);
  if (f.parent >= 0) {
    pf = mul(k.R[f.parent], pf) + k.p[f.parent];
    Rf = mul(k.R[f.parent], Rf);
  } else {
    pf = pf - k.origin;
  }
}

// LOCAL_WORLD_ALIGNED Jacobian of a point pf attached to joint `pj`:  J[r*NV + j], rows = lin(3), ang(3).
template <int NV, bool CHAIN>
DRC_HD void point_jacobian(const DrcModelDev& m, const KinState<NV>& k, int pj, Vec3 pf, double* J) {
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    Vec3 l = v3(0, 0, 0), w = v3(0, 0, 0);
    if (pj >= 0 && is_anc<CHAIN>(m, pj, j)) {
      if (m.jtype[j] == kRevolute) { l = cross(k.a[j], pf - k.p[j]); w = k.a[j]; }
      else l = k.a[j];
    }
    J[0 * NV + j] = l.x; J[1 * NV + j] = l.y; J[2 * NV + j] = l.z;
    J[3 * NV + j] = w.x; J[4 * NV + j] = w.y; J[5 * NV + j] = w.z;
  }
}
// Exact time derivative of the above for a body-fixed point (Pinocchio 3.x semantics):
//   col j = [ adot_j x (pf - p_j) + a_j x (pfdot - pdot_j) ; adot_j ],  adot_j = w_j x a_j.
template <int NV, bool CHAIN>
DRC_HD void point_jacobian_dot(const DrcModelDev& m, const KinState<NV>& k, const Spatial* v, int pj, Vec3 pf, double* Jd) {
  Vec3 pfd = v3(0, 0, 0);
  if (pj >= 0) pfd = v[pj].l + cross(v[pj].a, pf);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    Vec3 l = v3(0, 0, 0), w = v3(0, 0, 0);
    if (pj >= 0 && is_anc<CHAIN>(m, pj, j)) {
      const Vec3 ad = cross(v[j].a, k.a[j]);
      if (m.jtype[j] == kRevolute) {
        const Vec3 pjd = v[j].l + cross(v[j].a, k.p[j]);
        l = cross(ad, pf - k.p[j]) + cross(k.a[j], pfd - pjd);
        w = ad;
      } else {
        l = ad;
      }
    }
    Jd[0 * NV + j] = l.x; Jd[1 * NV + j] = l.y; Jd[2 * NV + j] = l.z;
    Jd[3 * NV + j] = w.x; Jd[4 * NV + j] = w.y; Jd[5 * NV + j] = w.z;
  }
}

// CRBA mass matrix (dense, symmetric, row-major NV x NV).
template <int NV, bool CHAIN>
DRC_HD void mass_matrix(const DrcModelDev& m, const KinState<NV>& k, double* M) {
  WInertia Yc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) Yc[i] = world_inertia(m, i, k.R[i], k.p[i]);
#pragma unroll
  for (int j = NV - 1; j >= 0; --j) {
    const Spatial F = apply(Yc[j], screw<NV>(m, k, j));
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i <= j) {
        const double v = is_anc<CHAIN>(m, j, i) ? dot6(screw<NV>(m, k, i), F) : 0.0;
        M[i * NV + j] = v;
        M[j * NV + i] = v;
      }
    }
    const int pr = parent_of<CHAIN>(m, j);
    if (pr >= 0) accumulate(Yc[pr], Yc[j]);
  }
}

// RNEA with qdd = 0: tau = C(q,qd) qd + g(q)  (nonLinearEffects); pass with_vel=false for gravity only.
template <int NV, bool CHAIN>
DRC_HD void rnea_bias(const DrcModelDev& m, const KinState<NV>& k, const Spatial* v, bool with_vel, const double* qd, double* tau) {
  Spatial f[NV], acc[NV];
  const Spatial a0 = Spatial{v3(-m.gravity[0], -m.gravity[1], -m.gravity[2]), v3(0, 0, 0)};
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const WInertia Y = world_inertia(m, i, k.R[i], k.p[i]);
    const int pr = parent_of<CHAIN>(m, i);
    Spatial ai = pr < 0 ? a0 : acc[pr];
    if (with_vel) {
      const Spatial S = screw<NV>(m, k, i);
      const Spatial c = cross_motion(v[i], Spatial{qd[i] * S.l, qd[i] * S.a});
      ai = ai + c;
      acc[i] = ai;
      f[i] = apply(Y, ai) + cross_force(v[i], apply(Y, v[i]));
    } else {
      acc[i] = ai;
      f[i] = apply(Y, ai);
    }
  }
#pragma unroll
  for (int i = NV - 1; i >= 0; --i) {
    tau[i] = dot6(screw<NV>(m, k, i), f[i]);
    const int pr = parent_of<CHAIN>(m, i);
    if (pr >= 0) f[pr] = f[pr] + f[i];
  }
}

// Inverse of the SPD mass matrix.  The reference uses PinvCOD(M) (robot_data.cpp:118); for a
// positive-definite M that is the inverse, taken here by Cholesky, with the rank-revealing route
// as the guard for (near-)singular input.
// `defer`: when the guard fails, return false with Ainv untouched -- the caller hands the robot to the group-cooperative
// rank-revealing kernel (k_pinv_list) instead of running the thread-serial routine here.
template <int N>
DRC_HD bool spd_pinv(const double* A, double* Ainv, double threshold, bool defer = false) {
  double L[N * N];
  double maxd = 0;
#pragma unroll
  for (int i = 0; i < N * N; ++i) L[i] = A[i];
#pragma unroll
  for (int i = 0; i < N; ++i) maxd = dmax(maxd, A[i * N + i]);
  const double minpiv = chol_inplace<N>(L);
  if (minpiv > 1e-5 * maxd) { chol_inverse<N>(L, Ainv); return true; }
  if (defer) return false;
  // below the guard the rank-revealing factorisation decides, as in the reference: it either truncates (its pseudo-inverse is the
  // result) or keeps every pivot -- then PinvCOD(A) = A^-1 and the Cholesky factor at hand gives it without forming Q
  if (pinv_cpqr<N, N>(A, Ainv, threshold, nullptr, minpiv > 0)) chol_inverse<N>(L, Ainv);
  return true;
}

// Cholesky route of the manipulability for a well-conditioned 6 x 6 SPD matrix: inverse, product of the factor's diagonal, and the
// certificate kappa <= min(tr A, |A|_inf) * min(tr A^-1, |A^-1|_inf) < 1.25e4 (see `manipulability`).  NOT inlined: one copy of this
// arithmetic per translation unit, so every kernel that evaluates a robot gets the same bits (as with pinv_cpqr).
DRC_HD_NOINLINE bool chol6_certified(const double* A, double* Ainv, double* prod_diag) {
  double L[36];
  double trA = 0, infA = 0;
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    double rs = 0;
#pragma unroll
    for (int c = 0; c < 6; ++c) { L[r * 6 + c] = A[r * 6 + c]; rs += fabs(A[r * 6 + c]); }
    trA += A[r * 6 + r]; infA = dmax(infA, rs);
  }
  if (!(chol_inplace<6>(L) > 0)) return false;
  chol_inverse<6>(L, Ainv);
  double trI = 0, infI = 0, det = 1.0;
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    double rs = 0;
#pragma unroll
    for (int c = 0; c < 6; ++c) rs += fabs(Ainv[r * 6 + c]);
    trI += Ainv[r * 6 + r]; infI = dmax(infI, rs);
    det *= L[r * 6 + r];
  }
  *prod_diag = det;
  return dmin(trA, infA) * dmin(trI, infI) < 1.25e4;
}

// Manipulability m = sqrt(det(J J^T)), gradient and the reference's "gradient time variation"
// (robot_data.cpp:519-573).  J, Jd are 6 x NC blocks (columns col0..col0+NC-1 of the frame Jacobian).
// dJ/dq_i is the closed-form kinematic Hessian (replaces the reference's NC extra J-dot passes):
//   i ancestor of j :  d(col j) = [ (a_i x a_j) x d_j + a_j x (a_i x d_j) ; a_i x a_j ],  d_j = pf - p_j
//   otherwise (i in the support):  d(col j) = [ a_j x (a_i x d_i) ; 0 ]
// (revolute joints; prismatic joints contribute a_i x a_j = 0 and d p = a_i).
template <int NV, int NC, bool CHAIN>
// `route`: 0 = rank-revealing QR only; 1 = Cholesky route or give up (returns false); 2 = Cholesky route, else the QR.
// Cholesky route -- for a well-conditioned J J^T the rank-revealing QR keeps every column, so its
// pseudo-inverse IS the inverse and |det| = prod of the Cholesky pivots.  The route is taken only under a certificate that the QR
// would not truncate: with kappa <= min(tr A, |A|_inf) * min(tr A^-1, |A^-1|_inf) < 1.25e4 the column-pivoted R satisfies
// min|R_kk| / max|R_kk| >= 1 / (sqrt(6) 2^5 kappa) > 1e-6 = the rank threshold.  Returns false (nothing computed) when the certificate
// fails: the caller hands that robot to the exact route (k_robot_job's fix-up launch).
DRC_HD bool manipulability(const DrcModelDev& m, const KinState<NV>& k, int pj, Vec3 pf, const double* J /*6xNV*/,
                           const double* Jd /*6xNV or null*/, int col0, bool with_graddot, double threshold,
                           double& mani, double* grad, double* grad_dot, int route = 0) {
  double A[36], Ainv[36];
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      double s = 0;
#pragma unroll
      for (int j = 0; j < NC; ++j) s += J[r * NV + col0 + j] * J[c * NV + col0 + j];
      A[r * 6 + c] = s;
    }
  bool certified = false;
  if (route != 0) {
    certified = threshold <= 1e-6 && chol6_certified(A, Ainv, &mani);   // mani = sqrt(det(J J^T)) = prod L_kk
    if (!certified && route == 1) return false;
  }
  if (!certified) {
    // the reference takes sqrt(det(JJ^T)) and PinvCOD(JJ^T) (robot_data.cpp:526,539): one rank-revealing
    // QR gives both (|det| = prod |R_kk|), including the rank truncation near singular postures
    double ad;
    pinv_cpqr<6, 6>(A, Ainv, threshold, &ad);
    mani = sqrt(ad);
  }
  // G = J^T Ainv  (NC x 6)
  double G[NC * 6];
#pragma unroll
  for (int j = 0; j < NC; ++j)
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      double s = 0;
#pragma unroll
      for (int r = 0; r < 6; ++r) s += J[r * NV + col0 + j] * Ainv[r * 6 + c];
      G[j * 6 + c] = s;
    }
  double H[NC * 6];
  double mani_dot = 0;
  if (with_graddot) {
    // mani_dot = m tr(Jd G);  Ainv_dot = -Ainv (2 Jd J^T) Ainv;  H = Jd^T Ainv + J^T Ainv_dot
    double tr = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int j = 0; j < NC; ++j) tr += Jd[r * NV + col0 + j] * G[j * 6 + r];
    mani_dot = mani * tr;
    double T1[36], T2[36];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) {  // T1 = 2 Jd J^T
        double s = 0;
#pragma unroll
        for (int j = 0; j < NC; ++j) s += Jd[r * NV + col0 + j] * J[c * NV + col0 + j];
        T1[r * 6 + c] = 2.0 * s;
      }
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) {  // T2 = Ainv T1
        double s = 0;
#pragma unroll
        for (int l = 0; l < 6; ++l) s += Ainv[r * 6 + l] * T1[l * 6 + c];
        T2[r * 6 + c] = s;
      }
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) {  // T1 = -(T2 Ainv) = Ainv_dot
        double s = 0;
#pragma unroll
        for (int l = 0; l < 6; ++l) s += T2[r * 6 + l] * Ainv[l * 6 + c];
        T1[r * 6 + c] = -s;
      }
#pragma unroll
    for (int j = 0; j < NC; ++j)
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double s = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r) s += Jd[r * NV + col0 + j] * Ainv[r * 6 + c] + J[r * NV + col0 + j] * T1[r * 6 + c];
        H[j * 6 + c] = s;
      }
  }
#pragma unroll
  for (int ii = 0; ii < NC; ++ii) {
    const int i = col0 + ii;
    double trG = 0, trH = 0;
    const bool i_in_support = pj >= 0 && is_anc<CHAIN>(m, pj, i);
    if (i_in_support) {
      const bool i_rev = m.jtype[i] == kRevolute;
      const Vec3 ai = k.a[i];
      const Vec3 di = pf - k.p[i];
      const Vec3 dpf = i_rev ? cross(ai, di) : ai;  // d pf / d q_i
#pragma unroll
      for (int jj = 0; jj < NC; ++jj) {
        const int j = col0 + jj;
        if (!is_anc<CHAIN>(m, pj, j)) continue;
        Vec3 dl, dw;
        const bool j_rev = m.jtype[j] == kRevolute;
        if (i != j && is_anc<CHAIN>(m, j, i)) {  // i strict ancestor of j
          const Vec3 dj = pf - k.p[j];
          const Vec3 daj = i_rev ? cross(ai, k.a[j]) : v3(0, 0, 0);
          const Vec3 ddj = i_rev ? cross(ai, dj) : v3(0, 0, 0);  // d(pf - p_j)/dq_i
          if (j_rev) { dl = cross(daj, dj) + cross(k.a[j], ddj); dw = daj; }
          else { dl = daj; dw = v3(0, 0, 0); }
        } else {  // i == j or i below j: only pf moves
          dl = j_rev ? cross(k.a[j], dpf) : v3(0, 0, 0);
          dw = v3(0, 0, 0);
        }
        trG += dl.x * G[jj * 6 + 0] + dl.y * G[jj * 6 + 1] + dl.z * G[jj * 6 + 2] + dw.x * G[jj * 6 + 3] + dw.y * G[jj * 6 + 4] + dw.z * G[jj * 6 + 5];
        if (with_graddot)
          trH += dl.x * H[jj * 6 + 0] + dl.y * H[jj * 6 + 1] + dl.z * H[jj * 6 + 2] + dw.x * H[jj * 6 + 3] + dw.y * H[jj * 6 + 4] + dw.z * H[jj * 6 + 5];
      }
    }
    grad[ii] = mani * trG;
    if (with_graddot) grad_dot[ii] = mani_dot * trG + mani * trH;
  }
  return true;
}

// DyrosMath::getPhi(target, current) orientation error (math_type_define.h:283-298, call at :642)
DRC_HD Vec3 orientation_error(const Mat3& Rt, const Mat3& R) {
  Vec3 s = cross(col(Rt, 0), col(R, 0)) + cross(col(Rt, 1), col(R, 1)) + cross(col(Rt, 2), col(R, 2));
  return -0.5 * s;
}

}  // namespace drc
