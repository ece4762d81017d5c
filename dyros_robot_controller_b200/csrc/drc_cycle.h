// drc_b200 -- per-robot stages of the control cycle (one robot per thread).
//
// Stage "robot_job": updateState (reference robot_data.cpp:91-124) and everything that only needs
// the robot's own kinematics: frame getters (:378-422), manipulability (:519-573), task-space error
// (math_type_define.h:633-645), CLIK / OSF (robot_controller.cpp:156-171,208-247), joint PD torque
// (:115-125) and the QPIK / QPID problem records (QP_IK.cpp:69-131, QP_ID.cpp:92-193).
// Stage "collision_job": min self-distance + gradients (robot_data.cpp:424-517) -> dense row 1.
// The template FLAGS select which parts a kernel instantiation contains.
#pragma once
#include "drc_geom.h"
#include "drc_kin.h"
#include "drc_mobile.h"
#include "drc_qp.h"
#include <type_traits>

namespace drc {

enum JobFlags : unsigned {
  F_DYN = 1u << 0,        // M, Minv, g, nle -> cache
  F_STORE = 1u << 1,      // q, qd, oMi -> cache
  F_FROM_CACHE = 1u << 2, // read q, qd from the cache instead of the inputs
  F_FRAME_OUT = 1u << 3,  // pose / J / Jdot / velocity outputs
  F_MANIP_OUT = 1u << 4,  // manipulability outputs
  F_QPIK = 1u << 5,
  F_QPID = 1u << 6,
  F_STEP = 1u << 7,       // desired task signal from (x_target, xdot_target) and the gains
  F_CLIK = 1u << 8,
  F_OSF = 1u << 9,
  F_TORQUE = 1u << 10,    // moveJointTorqueStep(q_target, qdot_target)
  F_GRADDOT = 1u << 11,   // manipulability gradient time variation
  F_MOMA = 1u << 12,      // mobile manipulator: actuated quantities through the selection matrix S, whole-body QPs
  F_DYN_LIGHT = 1u << 13, // with F_DYN: only what a QPID record reads (M, g; whole-body: M~, g~) and nothing but g~ stored -- a
                          // dynamics-only launch (F_DYN | F_FROM_CACHE) behind the solver completes the cache (inverses, nle)
};

struct JobIO {
  int B;
  // inputs (any layout)
  const double* q; Strided sq;
  const double* qd; Strided sqd;
  const double* x_target; Strided sxt;      // 12 per robot: top 3 rows of the homogeneous matrix
  const double* xdot_target; Strided sxd;   // 6 per robot (target velocity, or the desired xdot / xddot)
  const double* aux; Strided saux;          // n per robot: null-space vector (CLIK/OSF) or q_target (TORQUE)
  const double* aux2; Strided saux2;        // n per robot: qdot_target (TORQUE)
  // state cache (SoA, component stride Bc)
  double *c_q, *c_qd, *c_oMi, *c_M, *c_Minv, *c_g, *c_nle;
  double *c_Mact, *c_Minvact, *c_gact, *c_nleact;  // mobile manipulator: actuated-space dynamics (act x act, act)
  long long Bc;
  // outputs (any layout; null = skip)
  double* pose; Strided spose;
  double* J; Strided sJ;
  double* Jdot; Strided sJd;
  double* vel; Strided svel;
  double* mani; double* mani_grad; Strided smg; double* mani_graddot; Strided smgd;
  double* out; Strided sout;                // n per robot (CLIK qdot / OSF torque / PD torque)
  double* out2;                             // OSF torque when CLIK and OSF run in one launch (same layout)
  double* qp;                               // QP records (AoS, Cfg::STRIDE doubles per robot)
  // priority sub-batch (robots predicted to need many ADMM iterations, see run_qp): slot b of the cache / QP scratch
  // holds robot ids[b] of the INPUT arrays; *count slots are in use (null = identity / io.B)
  const int* ids; const int* count;
  // two-route manipulability (device, main pipeline of the fused QPIK cycle): robots whose J J^T fails the conditioning
  // certificate of the Cholesky route are appended here and redone by the exact (rank-revealing) route in a small follow-up
  // launch of the same job over this list (`redo`: the launch IS that follow-up; slot index = list entry).  null = exact route.
  int* manip_list; int* manip_count;
  bool redo;
  // PinvCOD(M) of the full model (robot_data.cpp:118): robots whose mass matrix fails the Cholesky guard of spd_pinv (whole-body
  // models: every robot) are appended here and k_pinv_list takes their rank-revealing route with 16 lanes per robot.  null = inline.
  int* pinv_list; int* pinv_count;
};

template <int NV>
using QpikCfg = QpCfg<NV, 2, 2, 0, true, true>;
template <int NV>
using QpidCfg = QpCfg<NV, 4, 2, NV, true, true>;
// MobileManipulator::QPIK / QPID (mobile_manipulator/QP_IK.cpp:14-28, QP_ID.cpp:14-36): hard CBF rows (no slacks);
// QPIK keeps its (free) bound rows, QPID has none.  ACT = wheels + manipulator joints.
template <int ACT>
using MomaIkCfg = QpCfg<ACT, 2, 2, 0, false, true>;
template <int ACT>
using MomaIdCfg = QpCfg<ACT, 4, 2, ACT, false, false>;

// (row vector r of the full model, length NV) * S  ->  ACT entries.  S is the reference's selection matrix
// (mobile_manipulator/robot_data.cpp:22-25,115-120): identity blocks for the wheel and manipulator joints plus
// S[virtual, mobile] = Rz(yaw) * J_mobile (Sm, 3 x W row-major).
template <int NV, int W>
DRC_HD void row_times_S(const DrcModelDev& m, const double* Sm, const double* r, int rs, double* out, int os) {
  constexpr int MANI = NV - 3 - W;
#pragma unroll
  for (int i = 0; i < MANI; ++i) out[(m.act_mani_start + i) * os] = r[(m.mani_start + i) * rs];
#pragma unroll
  for (int k = 0; k < W; ++k) {
    double s = r[(m.mobi_start + k) * rs];
#pragma unroll
    for (int a = 0; a < 3; ++a) s += r[(m.virtual_start + a) * rs] * Sm[a * W + k];
    out[(m.act_mobi_start + k) * os] = s;
  }
}

template <int NV, bool CHAIN, unsigned FLAGS, int W = 0>
DRC_HD void robot_job(const DrcModelDev& m, const DrcParams& prm, const DrcFrame& frame, const JobIO& io, int b) {
  constexpr bool MOMA = (FLAGS & F_MOMA) != 0;
  constexpr int ACT = MOMA ? NV - 3 : NV;        // actuated dof = wheels + manipulator joints
  constexpr int MANI = MOMA ? NV - 3 - W : NV;   // manipulator dof (mobile_manipulator/robot_data.cpp:19)
  static_assert(!(FLAGS & F_QPID) || !MOMA || (FLAGS & F_DYN), "whole-body QPID records read the actuated dynamics");
  constexpr bool LIGHT = (FLAGS & F_DYN_LIGHT) != 0;
  // mobile-manipulator jobs pick the state source and the task signal at run time (fewer heavy instantiations)
  const bool from_cache = (FLAGS & F_FROM_CACHE) || (MOMA && io.q == nullptr);
  const bool step = (FLAGS & F_STEP) || (MOMA && io.x_target != nullptr);
  const long long bi = io.ids ? io.ids[b] : b;  // robot index in the input arrays (b indexes cache / QP scratch)
  double q[NV], qd[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (from_cache) { q[i] = io.c_q[i * io.Bc + b]; qd[i] = io.c_qd[i * io.Bc + b]; }
    else { q[i] = io.q[bi * io.sq.sb + i * io.sq.sk]; qd[i] = io.qd[bi * io.sqd.sb + i * io.sqd.sk]; }
  }
  DRC_PHASE(PH_KIN);
  KinState<NV> k;
  k.origin = v3(0, 0, 0);
  forward_kinematics<NV, CHAIN>(m, q, k);
  Spatial v[NV];
  joint_velocities<NV, CHAIN>(m, k, qd, v);

  if (FLAGS & F_STORE) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      io.c_q[i * io.Bc + b] = q[i];
      io.c_qd[i * io.Bc + b] = qd[i];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) io.c_oMi[(12 * i + 4 * r + c) * io.Bc + b] = k.R[i].m[3 * r + c];
        io.c_oMi[(12 * i + 4 * r + 3) * io.Bc + b] = comp(k.p[i], r);
      }
    }
  }

  DRC_PHASE(PH_DYN);
  double M[NV * NV], g[NV], Minv[NV * NV];
  constexpr bool need_dyn_vals = (FLAGS & (F_QPID | F_OSF | F_TORQUE)) != 0;
  if (FLAGS & F_DYN) {
    mass_matrix<NV, CHAIN>(m, k, M);
    rnea_bias<NV, CHAIN>(m, k, v, false, qd, g);
    if (!LIGHT) {
      double nle[NV];
      rnea_bias<NV, CHAIN>(m, k, v, true, qd, nle);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        io.c_g[i * io.Bc + b] = g[i];
        io.c_nle[i * io.Bc + b] = nle[i];
#pragma unroll
        for (int j = 0; j < NV; ++j) io.c_M[(i * NV + j) * io.Bc + b] = M[i * NV + j];
      }
      bool have_inv = true;
#if defined(__CUDA_ARCH__)
      if (io.pinv_list) {
        have_inv = spd_pinv<NV>(M, Minv, prm.pinv_threshold, true);
        if (!have_inv) io.pinv_list[atomicAdd(io.pinv_count, 1)] = b;
      } else
#endif
        spd_pinv<NV>(M, Minv, prm.pinv_threshold);
      if (have_inv) {
#pragma unroll
        for (int i = 0; i < NV * NV; ++i) io.c_Minv[i * io.Bc + b] = Minv[i];
      }
    }
  } else if (need_dyn_vals) {
    if (FLAGS & F_OSF) {
#pragma unroll
      for (int i = 0; i < NV * NV; ++i) Minv[i] = io.c_Minv[i * io.Bc + b];
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      g[i] = io.c_g[i * io.Bc + b];
#pragma unroll
      for (int j = 0; j < NV; ++j) M[i * NV + j] = io.c_M[(i * NV + j) * io.Bc + b];
    }
  }

  // mobile manipulator: S, actuated dynamics M~ = S'MS, g~ = S'g, nle~ = S'nle (mobile_manipulator/robot_data.cpp:104-144)
  double Sm[MOMA ? 3 * W : 1], Mact[MOMA ? ACT * ACT : 1], gact[MOMA ? ACT : 1], qd_act[MOMA ? ACT : 1];
  if (MOMA) {
    double sy, cy;
    sincos(q[m.virtual_start + 2], &sy, &cy);
    // J_mobile: constant for differential / mecanum bases, a function of the steering angles for powered casters
    // (mobile_manipulator/robot_data.cpp:112 -> Mobile::RobotData::computeFKJacobian, mobile/robot_data.cpp:122-204)
    double Jm[3 * (W > 0 ? W : 1)];
    if (m.drive_type == kCaster) {
      double wp[W > 0 ? W : 1];
#pragma unroll
      for (int k = 0; k < W; ++k) wp[k] = q[m.mobi_start + k];
      caster_fk_jacobian<(W > 1 ? W : 2)>(m.wheel_radius, m.wheel_offset, m.b2w_x, m.b2w_y, wp, W, Jm, W);
    } else {
#pragma unroll
      for (int k = 0; k < W; ++k) { Jm[0 * W + k] = m.J_mobile[0][k]; Jm[1 * W + k] = m.J_mobile[1][k]; Jm[2 * W + k] = m.J_mobile[2][k]; }
    }
#pragma unroll
    for (int k = 0; k < W; ++k) {
      Sm[0 * W + k] = cy * Jm[0 * W + k] - sy * Jm[1 * W + k];
      Sm[1 * W + k] = sy * Jm[0 * W + k] + cy * Jm[1 * W + k];
      Sm[2 * W + k] = Jm[2 * W + k];
    }
#pragma unroll
    for (int i = 0; i < MANI; ++i) qd_act[m.act_mani_start + i] = qd[m.mani_start + i];
#pragma unroll
    for (int k = 0; k < W; ++k) qd_act[m.act_mobi_start + k] = qd[m.mobi_start + k];
    if (FLAGS & F_DYN) {
      double T[NV * ACT];
#pragma unroll
      for (int i = 0; i < NV; ++i) row_times_S<NV, W>(m, Sm, M + i * NV, 1, T + i * ACT, 1);
#pragma unroll
      for (int a = 0; a < ACT; ++a) row_times_S<NV, W>(m, Sm, T + a, ACT, Mact + a, ACT);
      row_times_S<NV, W>(m, Sm, g, 1, gact, 1);
#pragma unroll
      for (int i = 0; i < ACT; ++i) io.c_gact[i * io.Bc + b] = gact[i];   // the QPID fallback of the solver launch reads it
      if (!LIGHT) {
        double nleact[ACT], nle_full[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) nle_full[i] = io.c_nle[i * io.Bc + b];
        row_times_S<NV, W>(m, Sm, nle_full, 1, nleact, 1);
        if (io.c_Mact) {
          double Mi[ACT * ACT];
          spd_pinv<ACT>(Mact, Mi, prm.pinv_threshold);
#pragma unroll
          for (int i = 0; i < ACT * ACT; ++i) { io.c_Mact[i * io.Bc + b] = Mact[i]; io.c_Minvact[i * io.Bc + b] = Mi[i]; }
        }
#pragma unroll
        for (int i = 0; i < ACT; ++i) io.c_nleact[i * io.Bc + b] = nleact[i];
      }
    }
  }

  DRC_PHASE(PH_BUILD);
  if (FLAGS & F_TORQUE) {  // tau = M (Kp (q_t - q) + Kv (qd_t - qd)) + g
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i)
      acc[i] = prm.Kp_joint[i] * (io.aux[b * io.saux.sb + i * io.saux.sk] - q[i]) +
               prm.Kv_joint[i] * (io.aux2[b * io.saux2.sb + i * io.saux2.sk] - qd[i]);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double s = g[i];
#pragma unroll
      for (int j = 0; j < NV; ++j) s += M[i * NV + j] * acc[j];
      io.out[b * io.sout.sb + i * io.sout.sk] = s;
    }
  }

  constexpr bool need_frame = (FLAGS & (F_FRAME_OUT | F_MANIP_OUT | F_QPIK | F_QPID | F_CLIK | F_OSF)) != 0;
  if (!need_frame) return;

  DRC_PHASE(PH_KIN);
  Mat3 Rf;
  Vec3 pf;
  frame_pose<NV>(k, frame, Rf, pf);
  double J[6 * NV];
  point_jacobian<NV, CHAIN>(m, k, frame.parent, pf, J);
  double Jd[6 * NV];
  constexpr bool need_jdot = (FLAGS & (F_FRAME_OUT | F_QPID | F_GRADDOT)) != 0;
  if (need_jdot) point_jacobian_dot<NV, CHAIN>(m, k, v, frame.parent, pf, Jd);
  double xdot[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    double s = 0;
#pragma unroll
    for (int j = 0; j < NV; ++j) s += J[r * NV + j] * qd[j];
    xdot[r] = s;
  }

  // actuated Jacobians J~ = J S, J~dot = Jdot S (S-dot neglected, mobile_manipulator/robot_data.cpp:407-415)
  double Jt[MOMA ? 6 * ACT : 1], Jtd[MOMA ? 6 * ACT : 1];
  if (MOMA) {
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      row_times_S<NV, W>(m, Sm, J + r * NV, 1, Jt + r * ACT, 1);
      if (need_jdot) row_times_S<NV, W>(m, Sm, Jd + r * NV, 1, Jtd + r * ACT, 1);
    }
  }

  if (FLAGS & F_FRAME_OUT) {
    if (io.pose) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) io.pose[b * io.spose.sb + (4 * r + c) * io.spose.sk] = Rf.m[3 * r + c];
        io.pose[b * io.spose.sb + (4 * r + 3) * io.spose.sk] = comp(pf, r);
      }
    }
    if (MOMA) {  // the actuated Jacobians (getJacobianActuated / ...TimeVariation), 6 x ACT
#pragma unroll
      for (int i = 0; i < 6 * ACT; ++i) {
        if (io.J) io.J[b * io.sJ.sb + i * io.sJ.sk] = Jt[i];
        if (io.Jdot) io.Jdot[b * io.sJd.sb + i * io.sJd.sk] = Jtd[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 6 * NV; ++i) {
        if (io.J) io.J[b * io.sJ.sb + i * io.sJ.sk] = J[i];
        if (io.Jdot) io.Jdot[b * io.sJd.sb + i * io.sJd.sk] = Jd[i];
      }
    }
    if (io.vel) {
#pragma unroll
      for (int r = 0; r < 6; ++r) io.vel[b * io.svel.sb + r * io.svel.sk] = xdot[r];
    }
  }

  // manipulability
  DRC_PHASE(PH_MANI);
  double mani = 0, mgrad[NV], mgraddot[NV];
  if (FLAGS & (F_MANIP_OUT | F_QPIK | F_QPID)) {
    constexpr bool gd = (FLAGS & (F_QPID | F_GRADDOT)) != 0;
    // Two routes, chosen by the robot's own data only (so a robot's result does not depend on the launch it sits in): Cholesky when
    // the conditioning certificate holds (drc_kin.h), else the rank-revealing QR.  Where the caller provides a list (main pipeline of
    // the fused QPIK cycle) the uncertified robots are left to a follow-up launch over that list instead of stalling their warps here.
    int route = MOMA ? 0 : 2;   // whole-body kernels: the inline fallback would buy nothing (some lane of every warp takes it)
#ifdef DRC_FORCE_EXACT_MANIP   // test infrastructure: the rank-revealing route for every robot
    route = 0;
#endif
#if defined(__CUDA_ARCH__)
    if (io.redo) route = 0;                 // the follow-up launch: these robots failed the certificate already
    else if (io.manip_list) route = 1;
#endif
    if (!manipulability<NV, MANI, CHAIN>(m, k, frame.parent, pf, J, Jd, MOMA ? m.mani_start : 0, gd, prm.pinv_threshold, mani, mgrad, mgraddot, route)) {
#if defined(__CUDA_ARCH__)
      io.manip_list[atomicAdd(io.manip_count, 1)] = b;
#endif
      return;
    }
    if ((FLAGS & F_MANIP_OUT) && io.mani) {
      io.mani[b] = mani;
#pragma unroll
      for (int i = 0; i < MANI; ++i) {
        if (io.mani_grad) io.mani_grad[b * io.smg.sb + i * io.smg.sk] = mgrad[i];
        if (gd && io.mani_graddot) io.mani_graddot[b * io.smgd.sb + i * io.smgd.sk] = mgraddot[i];
      }
    }
  }

  // desired task-space signal
  DRC_PHASE(PH_BUILD);
  double des[6];
  // fused CLIKStep + OSFStep (one job, two outputs): the two controllers form different desired signals from the same error
  constexpr bool CLIK_AND_OSF = (FLAGS & F_CLIK) != 0 && (FLAGS & F_OSF) != 0;
  double des_osf[6];
  if (FLAGS & (F_QPIK | F_QPID | F_CLIK | F_OSF)) {
    double xd_t[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) xd_t[r] = io.xdot_target[bi * io.sxd.sb + r * io.sxd.sk];
    if (step) {
      Mat3 Rt;
      Vec3 pt;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) Rt.m[3 * r + c] = io.x_target[bi * io.sxt.sb + (4 * r + c) * io.sxt.sk];
      }
      pt = v3(io.x_target[bi * io.sxt.sb + 3 * io.sxt.sk], io.x_target[bi * io.sxt.sb + 7 * io.sxt.sk],
              io.x_target[bi * io.sxt.sb + 11 * io.sxt.sk]);
      const Vec3 ep = pt - pf, eo = orientation_error(Rt, Rf);
      const double xe[6] = {ep.x, ep.y, ep.z, eo.x, eo.y, eo.z};
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        // CLIKStep: Kp e + xdot_target (robot_controller.cpp:169); QPIK/QPID/OSF Step: Kp e + Kv edot (:238,:299,:342)
        // MobileManipulator QPIKStep: Kp e + xdot_target as well (mobile_manipulator/robot_controller.cpp:181)
        constexpr bool no_kv = (FLAGS & F_CLIK) != 0 || (MOMA && (FLAGS & F_QPIK) != 0);
        des[r] = prm.Kp_task[r] * xe[r] + (no_kv ? xd_t[r] : prm.Kv_task[r] * (xd_t[r] - xdot[r]));
        if (CLIK_AND_OSF) des_osf[r] = prm.Kp_task[r] * xe[r] + prm.Kv_task[r] * (xd_t[r] - xdot[r]);   // OSFStep (:238)
      }
    } else {
#pragma unroll
      for (int r = 0; r < 6; ++r) { des[r] = xd_t[r]; if (CLIK_AND_OSF) des_osf[r] = xd_t[r]; }
    }
  }

  if (FLAGS & F_CLIK) {
    // qdot = J+ des + (I - J+ J) null,  J+ = PinvCOD(J)
    double Jp[NV * 6];
    pinv_cpqr<6, NV>(J, Jp, prm.pinv_threshold);
    double y[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      double s = des[r];
      if (io.aux) {
#pragma unroll
        for (int j = 0; j < NV; ++j) s -= J[r * NV + j] * io.aux[b * io.saux.sb + j * io.saux.sk];
      }
      y[r] = s;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double s = io.aux ? io.aux[b * io.saux.sb + i * io.saux.sk] : 0.0;
#pragma unroll
      for (int r = 0; r < 6; ++r) s += Jp[i * 6 + r] * y[r];
      io.out[b * io.sout.sb + i * io.sout.sk] = s;
    }
  }

  if (FLAGS & F_OSF) {
    // Lambda = pinv(J Minv J^T); tau = J^T Lambda xddot + (I - J^T Lambda J Minv) tau_null + g
    double JMi[6 * NV], Li[36], Lam[36];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        double s = 0;
#pragma unroll
        for (int l = 0; l < NV; ++l) s += J[r * NV + l] * Minv[l * NV + j];
        JMi[r * NV + j] = s;
      }
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double s = 0;
#pragma unroll
        for (int j = 0; j < NV; ++j) s += JMi[r * NV + j] * J[c * NV + j];
        Li[r * 6 + c] = s;
      }
    pinv_cpqr<6, 6>(Li, Lam, prm.pinv_threshold);
    // F = Lambda (xddot - J Minv tau_null)
    double y[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      double s = CLIK_AND_OSF ? des_osf[r] : des[r];
      if (io.aux) {
#pragma unroll
        for (int j = 0; j < NV; ++j) s -= JMi[r * NV + j] * io.aux[b * io.saux.sb + j * io.saux.sk];
      }
      y[r] = s;
    }
    double F[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      double s = 0;
#pragma unroll
      for (int c = 0; c < 6; ++c) s += Lam[r * 6 + c] * y[c];
      F[r] = s;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double s = g[i] + (io.aux ? io.aux[b * io.saux.sb + i * io.saux.sk] : 0.0);
#pragma unroll
      for (int r = 0; r < 6; ++r) s += J[r * NV + i] * F[r];
      double* o = (FLAGS & F_CLIK) ? io.out2 : io.out;  // fused CLIK+OSF: second output array
      o[b * io.sout.sb + i * io.sout.sk] = s;
    }
  }

  if (FLAGS & (F_QPIK | F_QPID)) {
    constexpr bool ID = (FLAGS & F_QPID) != 0;
    typedef typename std::conditional<MOMA, typename std::conditional<ID, MomaIdCfg<ACT>, MomaIkCfg<ACT>>::type,
                                      typename std::conditional<ID, QpidCfg<NV>, QpikCfg<NV>>::type>::type Cfg;
    static_assert(Cfg::NC == ACT, "core variables of the QP = actuated joints");
    double* rec = io.qp + (long long)b * Cfg::STRIDE;
    const double al = prm.alpha;
    // core-space views: the manipulator QPs act on qdot / qddot of the arm itself, the whole-body QPs on eta / eta_dot
    const double* Jc = MOMA ? Jt : J;
    const double* Jdc = MOMA ? Jtd : Jd;
    const double* vc = MOMA ? qd_act : qd;
    const int am = MOMA ? m.act_mani_start : 0, ms = MOMA ? m.mani_start : 0;
    double rhs[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      double s = des[r];
      if (ID) {
#pragma unroll
        for (int j = 0; j < ACT; ++j) s -= Jdc[r * ACT + j] * vc[j];
      }
      rhs[r] = s;
    }
    const double reg = ID ? 0.0 : (MOMA ? prm.moma_ik_reg : prm.ik_reg);
#pragma unroll
    for (int i = 0; i < ACT; ++i) {
#pragma unroll
      for (int j = i; j < ACT; ++j) {
        double s = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r) s += Jc[r * ACT + i] * Jc[r * ACT + j];
        rec[Cfg::OFF_P + symidx<ACT>(i, j)] = 2.0 * s + (i == j ? reg : 0.0);
      }
      double s = 0;
#pragma unroll
      for (int r = 0; r < 6; ++r) s += Jc[r * ACT + i] * rhs[r];
      rec[Cfg::OFF_Q + i] = -2.0 * s;
      // bound rows: qdot limits for the manipulator QPIK, free otherwise (QP_ID.cpp / mobile_manipulator/QP_IK.cpp:75-83)
      const bool bounded = !ID && !MOMA;
      rec[Cfg::OFF_LO + i] = bounded ? -m.v_lim[i] : -kOsqpInfty;
      rec[Cfg::OFF_HI + i] = bounded ? m.v_lim[i] : kOsqpInfty;
#pragma unroll
      for (int k = 0; k < Cfg::KU; ++k) rec[Cfg::OFF_UNIT + k * ACT + i] = -kOsqpInfty;  // rows of non-manipulator variables are masked off
    }
#pragma unroll
    for (int i = 0; i < MANI; ++i) {  // CBF rows on the manipulator joints
      const double qi = q[ms + i], qdi = qd[ms + i], lo = m.q_lo[ms + i], hi = m.q_hi[ms + i], vl = m.v_lim[ms + i];
      if (!ID) {
        rec[Cfg::OFF_UNIT + 0 * ACT + am + i] = -al * (qi - lo);
        rec[Cfg::OFF_UNIT + 1 * ACT + am + i] = -al * (hi - qi);
      } else {
        rec[Cfg::OFF_UNIT + 0 * ACT + am + i] = -(al + al) * qdi - al * al * (qi - lo);
        rec[Cfg::OFF_UNIT + 1 * ACT + am + i] = +(al + al) * qdi - al * al * (hi - qi);
        rec[Cfg::OFF_UNIT + 2 * ACT + am + i] = -al * (qdi + vl);
        rec[Cfg::OFF_UNIT + 3 * ACT + am + i] = -al * (vl - qdi);
      }
    }
    // dense row 0: singularity avoidance (manipulator columns only)
    double* row0 = rec + Cfg::OFF_ROW;
    double gq = 0, gdq = 0;
#pragma unroll
    for (int i = 0; i < ACT; ++i) row0[i] = 0.0;
#pragma unroll
    for (int i = 0; i < MANI; ++i) { row0[am + i] = mgrad[i]; gq += mgrad[i] * qd[ms + i]; if (ID) gdq += mgraddot[i] * qd[ms + i]; }
    row0[ACT] = ID ? (-gdq - (al + al) * gq - al * al * (mani - prm.mani_thresh)) : (-al * (mani - prm.mani_thresh));
    // dense row 1 (self-collision) is written by collision_job
    if (ID) {  // equality rows  M qddot - tau = -g  (actuated space for the whole-body QP)
      const double* Mc = MOMA ? Mact : M;
      const double* gc = MOMA ? gact : g;
#pragma unroll
      for (int i = 0; i < ACT; ++i) {
        double* row = rec + Cfg::OFF_ROW + (2 + i) * (ACT + 1);
#pragma unroll
        for (int j = 0; j < ACT; ++j) row[j] = Mc[i * ACT + j];
        row[ACT] = -gc[i];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Self-collision stage
// ------------------------------------------------------------------------------------------------
struct CollisionIO {
  int B;
  const double *c_q, *c_qd, *c_oMi;  // state cache (SoA, stride Bc)
  long long Bc;
  double* qp;                        // QP records; row written: dense row 1
  int qp_stride, qp_row_off;         // Cfg::STRIDE, offset of dense row 1 inside the record
  int row_n, row_col0, src0, nsrc;   // row has row_n coefficients; joints src0..src0+nsrc-1 go to columns row_col0.. (0 = all NV)
  int mode;                          // 0 getter only, 1 QPIK row, 2 QPID row
  double* dist; double* grad; Strided sgrad; double* grad_dot; Strided sgd;  // getter outputs (null = skip)
  int* pair_out; double* witness;    // optional: argmin pair (reference order) and pa|pb (6, AoS)
  // hand-over to the EPA kernel
  int* epa_flag;                     // per robot: number of overlapping GJK pairs still to resolve
  unsigned long long* cand_mask;     // per robot: bit i = GJK-type pair i must be resolved by EPA
  int* epa_list; int* epa_count;     // compacted list of flagged robots (device: atomic append), may be null
  const int* count;                  // number of slots in use (device memory; null = B)
  // hand-over from the closed-form kernel to the GJK kernel (device, split narrow phase): winner's pair index, candidate
  // lower bounds [i][Bc]; the running minimum travels in dist / pair_out / witness, the candidate set in cand_mask
  int* nar_k; float* nar_lb;
};

struct JointFrame {
  Mat3 R;
  Vec3 p;
};
DRC_HD JointFrame load_joint_frame(const double* oMi, long long Bc, int b, int j) {
  JointFrame f;
  if (j < 0) { f.R = identity3(); f.p = v3(0, 0, 0); return f; }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) f.R.m[3 * r + c] = oMi[(12 * j + 4 * r + c) * Bc + b];
  }
  f.p = v3(oMi[(12 * j + 3) * Bc + b], oMi[(12 * j + 7) * Bc + b], oMi[(12 * j + 11) * Bc + b]);
  return f;
}
// geometry g of the model, placed by the transform (R, p) applied to its joint-frame placement
DRC_HD Prim place_prim(const GeomTable& G, int g, const Mat3& R, Vec3 p, bool identity) {
  Prim s;
  s.type = G.type[g];
  s.r = G.prm[g][0]; s.h = G.prm[g][1];
  s.hb = v3(G.prm[g][0], G.prm[g][1], G.prm[g][2]);
  const Vec3 cl = v3(G.p[g][0], G.p[g][1], G.p[g][2]);
  const Vec3 al = v3(G.R[g][2], G.R[g][5], G.R[g][8]);
  if (identity) { s.c = cl; s.a = al; }
  else { s.c = mul(R, cl) + p; s.a = mul(R, al); }
  if (s.type == kBox || s.type == kConvex) s.R = identity ? mat3_from(G.R[g]) : mul(R, mat3_from(G.R[g]));
  else s.R = identity3();
  s.verts = nullptr; s.nvert = 0;
  if (s.type == kConvex) { s.verts = G.hull + 3 * G.vert_off[g]; s.nvert = G.vert_n[g]; s.r = G.brad[g]; }
  return s;
}

struct BestPair {
  double d;
  int id;       // reference pair order (tie-break: smallest id wins, robot_data.cpp:434-443)
  int ja, jb;
  Vec3 pa, pb;  // in joint A's frame
};
DRC_HD void consider(BestPair& best, double d, int id, int ja, int jb, Vec3 pa, Vec3 pb) {
  if (d < best.d || (d == best.d && id < best.id)) { best.d = d; best.id = id; best.ja = ja; best.jb = jb; best.pa = pa; best.pb = pb; }
}

// gradient (+ time variation) of the distance for the winning pair, then outputs / QP row
template <int NV, bool CHAIN>
DRC_HD void collision_finish(const DrcModelDev& m, const DrcParams& prm, const CollisionIO& io, int b, const BestPair& best) {
  // world witness points
  const JointFrame FA = load_joint_frame(io.c_oMi, io.Bc, b, best.ja);
  const Vec3 pA = mul(FA.R, best.pa) + FA.p, pB = mul(FA.R, best.pb) + FA.p;
  const Vec3 dn = pB - pA;
  const double ln = norm(dn);
  const Vec3 n = ln > 1e-12 ? (1.0 / ln) * dn : v3(0, 0, 0);
  double qd[NV];
  Vec3 a[NV], p[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    qd[j] = io.c_qd[j * io.Bc + b];
    const JointFrame Fj = load_joint_frame(io.c_oMi, io.Bc, b, j);
    a[j] = mul(Fj.R, v3(m.axis[j][0], m.axis[j][1], m.axis[j][2]));
    p[j] = Fj.p;
  }
  const bool with_gd = io.mode == 2 || io.grad_dot != nullptr;
  // joint angular velocity w_j and origin velocity pd_j (for the time variation)
  Vec3 w[NV], pd[NV];
  Vec3 pAd = v3(0, 0, 0), pBd = v3(0, 0, 0);
  if (with_gd) {
    Vec3 vl[NV];  // spatial velocity (linear part at the origin)
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int pr = parent_of<CHAIN>(m, j);
      Vec3 sl, sa;
      if (m.jtype[j] == kRevolute) { sl = cross(p[j], a[j]); sa = a[j]; } else { sl = a[j]; sa = v3(0, 0, 0); }
      w[j] = qd[j] * sa; vl[j] = qd[j] * sl;
      if (pr >= 0) { w[j] = w[j] + w[pr]; vl[j] = vl[j] + vl[pr]; }
      pd[j] = vl[j] + cross(w[j], p[j]);
    }
    if (best.ja >= 0) pAd = vl[best.ja] + cross(w[best.ja], pA);
    if (best.jb >= 0) pBd = vl[best.jb] + cross(w[best.jb], pB);
  }
  double grad[NV], gdot[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    Vec3 ja = v3(0, 0, 0), jb = v3(0, 0, 0), jad = v3(0, 0, 0), jbd = v3(0, 0, 0);
    const bool rev = m.jtype[j] == kRevolute;
    const bool onA = best.ja >= 0 && is_anc<CHAIN>(m, best.ja, j), onB = best.jb >= 0 && is_anc<CHAIN>(m, best.jb, j);
    if (onA) ja = rev ? cross(a[j], pA - p[j]) : a[j];
    if (onB) jb = rev ? cross(a[j], pB - p[j]) : a[j];
    double gj = dot(n, jb - ja);
    if (best.d < 0) gj = -gj;
    grad[j] = gj;
    if (with_gd) {
      const Vec3 ad = cross(w[j], a[j]);
      if (onA) jad = rev ? cross(ad, pA - p[j]) + cross(a[j], pAd - pd[j]) : ad;
      if (onB) jbd = rev ? cross(ad, pB - p[j]) + cross(a[j], pBd - pd[j]) : ad;
      gdot[j] = dot(n, jbd - jad);  // n_dot neglected, no sign flip (robot_data.cpp:513)
    } else {
      gdot[j] = 0;
    }
  }
  if (io.dist) io.dist[b] = best.d;
  if (io.pair_out) io.pair_out[b] = best.id;
  if (io.witness) {
    io.witness[6 * b + 0] = pA.x; io.witness[6 * b + 1] = pA.y; io.witness[6 * b + 2] = pA.z;
    io.witness[6 * b + 3] = pB.x; io.witness[6 * b + 4] = pB.y; io.witness[6 * b + 5] = pB.z;
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (io.grad) io.grad[b * io.sgrad.sb + j * io.sgrad.sk] = grad[j];
    if (io.grad_dot) io.grad_dot[b * io.sgd.sb + j * io.sgd.sk] = gdot[j];
  }
  if (io.mode != 0) {
    double* row = io.qp + (long long)b * io.qp_stride + io.qp_row_off;
    const double al = prm.alpha;
    double gq = 0, gdq = 0;
    // whole-body QPs keep the manipulator slice of the gradient only (mobile_manipulator/QP_IK.cpp:121)
    const int rn = io.row_n > 0 ? io.row_n : NV, c0 = io.row_n > 0 ? io.row_col0 : 0, s0 = io.row_n > 0 ? io.src0 : 0,
              ns = io.row_n > 0 ? io.nsrc : NV;
    for (int j = 0; j < rn; ++j) row[j] = 0.0;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (j >= s0 && j < s0 + ns) { row[c0 + j - s0] = grad[j]; gq += grad[j] * qd[j]; gdq += gdot[j] * qd[j]; }
    }
    row[rn] = io.mode == 2 ? (-gdq - (al + al) * gq - al * al * (best.d - prm.dist_thresh)) : (-al * (best.d - prm.dist_thresh));
  }
}

// Main self-distance pass, in stages so that the device can run the GJK stage with the lanes of a block compacted onto the
// robots that need it (drc_kernels.cuh `collision_block`); `collision_job` below is the same sequence for ONE robot.
//   narrow_closed_phase  closed-form pairs exactly, in a loop that is uniform across threads (model indices are warp-uniform ->
//                        constant-bank broadcasts); GJK-type pairs get their certified lower bound there
//   narrow_pick          best first: the live candidate with the smallest lower bound, or -1 when none can still win
//   narrow_gjk_item      GJK on one (robot, candidate) item -- any thread may run it (divergent model indices -> `G` is the
//                        shared-memory copy of the geometry table)
//   narrow_apply         the owner folds the item's result into its running minimum
//   narrow_finish        overlapping pairs -> hand-over to the EPA kernel (cand_mask / epa_flag), else gradients + QP row
struct NarrowState {
  BestPair best;
  unsigned long long cand;      // bit i: i-th GJK-type pair whose bound beats the best exact distance so far
  unsigned long long deferred;  // bit i: that pair overlaps -> EPA
  float lbs[64];                // certified lower bounds (rounded DOWN to float: still lower bounds)
};
struct GjkItemResult {
  double d;
  Vec3 pa, pb;
  int state;  // 0 culled by the exact re-test of the bound, 1 separated (d, pa, pb valid), 2 overlapping
};

template <int NV, bool CHAIN>
DRC_HD void narrow_closed_phase(const DrcModelDev& m, const CollisionIO& io, int b, NarrowState& st, int* best_k_out = nullptr) {
  BestPair& best = st.best;
  best.d = 1e300; best.id = 1 << 30; best.ja = -1; best.jb = -1; best.pa = v3(0, 0, 0); best.pb = v3(0, 0, 0);
  int best_k = -1;  // only (distance, reference order, pair index) of the running minimum travel through the loop; the witness
                    // points of the winner are re-evaluated once afterwards
  st.cand = 0ull; st.deferred = 0ull;
  int gi = 0;
  for (int grp = 0; grp < m.ngroup; ++grp) {
    const int ja = m.group_ja[grp], jb = m.group_jb[grp];
    const JointFrame FA = load_joint_frame(io.c_oMi, io.Bc, b, ja), FB = load_joint_frame(io.c_oMi, io.Bc, b, jb);
    const Mat3 Rab = tmul(FA.R, FB.R);
    const Vec3 pab = tmul(FA.R, FB.p - FA.p);
    int placed = -1;  // pairs are sorted by the second geometry inside a group: it is placed once for all its partners
    Prim Bp;
    for (int t = 0; t < m.group_count[grp]; ++t) {
      const int k = m.group_first[grp] + t;
      const int ga = m.geom.pair_a[k], gb = m.geom.pair_b[k];
      if (gb != placed) { Bp = place_prim(m.geom, gb, Rab, pab, false); placed = gb; }
      const int ta = m.geom.type[ga];
      if (has_closed_form(ta, Bp.type)) {
        const Prim A = place_prim(m.geom, ga, Rab, pab, true);
        const double d = closed_form_distance(A, Bp).d;
        const int id = m.geom.pair_id[k];
        if (d < best.d || (d == best.d && id < best.id)) { best.d = d; best.id = id; best_k = k; }
      } else {
        // bounding spheres first: a pair whose spheres are further apart than the running minimum cannot win, now or later (the
        // minimum only decreases) -- it is no candidate, and the costlier certified bound is skipped
        const Vec3 dc = Bp.c - v3(m.geom.p[ga][0], m.geom.p[ga][1], m.geom.p[ga][2]);
        const double reach = best.d + m.geom.brad[ga] + m.geom.brad[gb];
        float lf = 3.0e38f;
        if (!(reach >= 0.0 && dot(dc, dc) > reach * reach * (1.0 + 1e-12))) {
          const Prim A = place_prim(m.geom, ga, Rab, pab, true);
          const double lb = pair_lower_bound(A, Bp);
          if (lb <= best.d) st.cand |= 1ull << gi;
          lf = (float)lb;
          if ((double)lf > lb) lf = lf - fabsf(lf) * 1.2e-7f - 1e-30f;
        }
        st.lbs[gi] = lf;
        ++gi;
      }
    }
  }
  if (best_k >= 0) {  // witness points of the closed-form winner (same evaluation as in the loop: same bits)
    const int ga = m.geom.pair_a[best_k], gb = m.geom.pair_b[best_k];
    const int ja = m.geom.parent[ga], jb = m.geom.parent[gb];
    const JointFrame FA = load_joint_frame(io.c_oMi, io.Bc, b, ja), FB = load_joint_frame(io.c_oMi, io.Bc, b, jb);
    const Mat3 Rab = tmul(FA.R, FB.R);
    const Vec3 pab = tmul(FA.R, FB.p - FA.p);
    const Prim A = place_prim(m.geom, ga, Rab, pab, true), Bp = place_prim(m.geom, gb, Rab, pab, false);
    const PairResult r = closed_form_distance(A, Bp);
    best.ja = ja; best.jb = jb; best.pa = r.pa; best.pb = r.pb;
  }
  if (best_k_out) *best_k_out = best_k;
}

// the most promising live candidate (smallest lower bound), removed from the candidate set; -1: none can still win
// (ties with the running minimum are resolved by the exact re-test in narrow_gjk_item)
DRC_HD int narrow_pick(NarrowState& st) {
  int bit = -1;
  float blb = 3.0e38f;
  for (unsigned long long rem = st.cand; rem; rem &= rem - 1ull) {
#if defined(__CUDA_ARCH__)
    const int i = __ffsll((long long)rem) - 1;
#else
    int i = 0;
    while (!((rem >> i) & 1ull)) ++i;
#endif
    if (st.lbs[i] < blb) { blb = st.lbs[i]; bit = i; }
  }
  if (bit < 0 || (double)blb > st.best.d) { st.cand = 0ull; return -1; }
  st.cand &= ~(1ull << bit);
  return bit;
}

DRC_HD void narrow_gjk_item(const DrcModelDev& m, const GeomTable& G, const DrcParams& prm, const CollisionIO& io, int b, int bit,
                            double best_d, GjkItemResult& res) {
  const int k = m.gjk_pair[bit];
  const int ga = G.pair_a[k], gb = G.pair_b[k];
  const int ja = G.parent[ga], jb = G.parent[gb];
  const JointFrame FA = load_joint_frame(io.c_oMi, io.Bc, b, ja), FB = load_joint_frame(io.c_oMi, io.Bc, b, jb);
  const Mat3 Rab = tmul(FA.R, FB.R);
  const Vec3 pab = tmul(FA.R, FB.p - FA.p);
  const Prim A = place_prim(G, ga, Rab, pab, true), Bp = place_prim(G, gb, Rab, pab, false);
  res.state = 0;
  if (pair_lower_bound(A, Bp) > best_d) return;
  GjkOut g;
  gjk_distance(A, Bp, prm.gjk_tol, prm.gjk_max_iter, g);
  res.state = g.intersect ? 2 : 1;
  res.d = g.dist; res.pa = g.pa; res.pb = g.pb;
}

DRC_HD void narrow_apply(const DrcModelDev& m, const GeomTable& G, NarrowState& st, int bit, const GjkItemResult& res) {
  if (res.state == 2) st.deferred |= 1ull << bit;
  else if (res.state == 1) {
    const int k = m.gjk_pair[bit];
    const int ga = G.pair_a[k], gb = G.pair_b[k];
    consider(st.best, res.d, G.pair_id[k], G.parent[ga], G.parent[gb], res.pa, res.pb);
  }
}

template <int NV, bool CHAIN>
DRC_HD void narrow_finish(const DrcModelDev& m, const DrcParams& prm, const CollisionIO& io, int b, const NarrowState& st) {
  const BestPair& best = st.best;
  if (st.deferred) {
    io.cand_mask[b] = st.deferred;
    io.epa_flag[b] = 1;
    if (io.epa_list) {
#if defined(__CUDA_ARCH__)
      io.epa_list[atomicAdd(io.epa_count, 1)] = b;
#else
      io.epa_list[(*io.epa_count)++] = b;
#endif
    }
    // provisional result so that the EPA kernel can resume from the best separated pair
    io.witness[6 * b + 0] = best.pa.x; io.witness[6 * b + 1] = best.pa.y; io.witness[6 * b + 2] = best.pa.z;
    io.witness[6 * b + 3] = best.pb.x; io.witness[6 * b + 4] = best.pb.y; io.witness[6 * b + 5] = best.pb.z;
    io.dist[b] = best.d;
    io.pair_out[b] = best.id;
    return;
  }
  io.epa_flag[b] = 0;
  collision_finish<NV, CHAIN>(m, prm, io, b, best);
}

template <int NV, bool CHAIN>
DRC_HD void collision_job(const DrcModelDev& m, const GeomTable& G, const DrcParams& prm, const CollisionIO& io, int b) {
  NarrowState st;
  narrow_closed_phase<NV, CHAIN>(m, io, b, st);
  for (int bit = narrow_pick(st); bit >= 0; bit = narrow_pick(st)) {
    GjkItemResult res;
    narrow_gjk_item(m, G, prm, io, b, bit, st.best.d, res);
    narrow_apply(m, G, st, bit, res);
  }
  narrow_finish<NV, CHAIN>(m, prm, io, b, st);
}

// EPA pass: only robots flagged by collision_job.
template <int NV, bool CHAIN>
DRC_HD void collision_epa_job(const DrcModelDev& m, const DrcParams& prm, const CollisionIO& io, int b) {
  if (!io.epa_flag[b]) return;
  unsigned long long deferred = io.cand_mask[b];
  BestPair best;
  best.d = io.dist[b]; best.id = io.pair_out[b];
  best.pa = v3(io.witness[6 * b + 0], io.witness[6 * b + 1], io.witness[6 * b + 2]);
  best.pb = v3(io.witness[6 * b + 3], io.witness[6 * b + 4], io.witness[6 * b + 5]);
  best.ja = -1; best.jb = -1;
  // recover the joints of the provisional best pair
  for (int k = 0; k < m.npair; ++k)
    if (m.geom.pair_id[k] == best.id) { best.ja = m.geom.parent[m.geom.pair_a[k]]; best.jb = m.geom.parent[m.geom.pair_b[k]]; }
  while (deferred) {
    int bit = 0;
    while (!((deferred >> bit) & 1ull)) ++bit;
    deferred &= deferred - 1ull;
    const int k = m.gjk_pair[bit];
    const int ga = m.geom.pair_a[k], gb = m.geom.pair_b[k];
    const int ja = m.geom.parent[ga], jb = m.geom.parent[gb];
    const JointFrame FA = load_joint_frame(io.c_oMi, io.Bc, b, ja), FB = load_joint_frame(io.c_oMi, io.Bc, b, jb);
    const Mat3 Rab = tmul(FA.R, FB.R);
    const Vec3 pab = tmul(FA.R, FB.p - FA.p);
    const Prim A = place_prim(m.geom, ga, Rab, pab, true), Bp = place_prim(m.geom, gb, Rab, pab, false);
    GjkOut g;
    gjk_distance(A, Bp, prm.gjk_tol, prm.gjk_max_iter, g);
    PairResult r;
    r.d = g.dist; r.pa = g.pa; r.pb = g.pb;
    if (g.intersect) epa_penetration(A, Bp, g, prm.epa_tol, prm.epa_max_iter, r);
    consider(best, r.d, m.geom.pair_id[k], ja, jb, r.pa, r.pb);
  }
  collision_finish<NV, CHAIN>(m, prm, io, b, best);
}

// ------------------------------------------------------------------------------------------------
// QP solve + reference fallbacks (robot_controller.cpp:277-290: zeros on failure; :319-333: gravity)
// ------------------------------------------------------------------------------------------------
struct SolveIO {
  int B;
  const double* qp;
  double* out; Strided sout;    // QPIK: qdot* (NC) | QPID: tau* (NC)
  double* out2; Strided sout2;  // QPID: qddot* (optional)
  int* status; int* iters;      // per robot (optional)
  const double* c_g; long long Bc;  // gravity cache for the QPID fallback
  double* qp_x;                 // optional debug: [core x (NC) | unit slacks (KU*NC) | row singletons (NR)] per robot
  double* qp_y;                 // optional debug: unscaled duals [core bound rows (NC) | unit rows (KU*NC) | unit-slack bound rows (KU*NC) | rows (NR) | row-singleton bound rows (NR)]
  const int* order;             // optional schedule: the i-th solver group takes robot order[i] (null = identity)
  int* iters_hint;              // optional: iteration count of every robot, kept by the context for the next tick's schedule
  const int* order_off;         // optional (device): skip the first *order_off entries of `order` (they run in the priority launch)
  const int* out_ids;           // optional: outputs / hints of QP slot i go to robot out_ids[i] (priority launch: records are compact)
  const int* count;             // optional (device): number of slots (null = B)
  const int* skip;              // optional: robots (QP slots) with skip[slot] != 0 are left to another launch (EPA-pending robots)
  // closed-loop rollout (drc_batch_rollout_qpik): the solver launch also performs the caller's integrate step
  // q += dt * qdot*, qdot = qdot* (examples/C++/src/fr3_controller.cpp:129-131, ideal tracking) and the per-robot tallies
  double* roll_q; double* roll_qd; Strided sroll; double roll_dt;
  int* fail_ticks; int* iters_total;
  int* hist_next; int* offs_next; int* sched_ticket;   // next tick's schedule (device only, see k_admm)
  // optional warm start (rollouts with drc_params_t::rollout_warm_start): a previous solve's qp_x / qp_y, indexed by robot
  const double* warm_x; const double* warm_y;
};

template <class Cfg, bool ID, class W>
DRC_HD void solve_and_emit(W& w, const int* robots, const SolveIO& io, const QpOptions& o) {
  WarmStart ws;
  ws.x = io.warm_x; ws.y = io.warm_y; ws.ids = io.out_ids;
  admm_solve<Cfg>(w, io.qp, robots, o, ws);
  DRC_PHASE(PH_QP_EMIT);
  w.each([&](Lane<Cfg>& L, GroupShared<Cfg>& S) {
    const int slot = S.robot;  // index of the QP record / state-cache entry
    if (slot < 0) return;
    const long long b = io.out_ids ? io.out_ids[slot] : slot;  // robot index in the output arrays
    const bool ok = S.status == kQpSolved;
    // warm-started rollouts: an infeasible / non-convex solve leaves zeros, i.e. the next tick of that robot starts cold
    const double wk = (!io.warm_x || S.status == kQpSolved || S.status == kQpMaxIter || S.status == kQpSolvedInaccurate) ? 1.0 : 0.0;
    if (L.gl == 0) {
      if (io.status) io.status[b] = S.status;
      if (io.iters) io.iters[b] = S.iters;
      if (io.iters_hint) io.iters_hint[b] = S.iters;
      if (io.fail_ticks) io.fail_ticks[b] += ok ? 0 : 1;
      if (io.iters_total) io.iters_total[b] += S.iters;
    }
    const ColdLane<Cfg>& C = S.cold[L.gl];
    constexpr int NY = Cfg::NC * (1 + 2 * Cfg::KU) + 2 * Cfg::NR;
    double* yr = io.qp_y ? io.qp_y + (long long)b * NY : nullptr;
    // OSQP unscale_solution: y = E y_scaled / c, y_scaled = rho (v - Proj(v))
    auto dual_row = [&](int k) { return wk * S.cinv * C.E[k] * (L.rho_r * (L.b[k].v - proj_row(L, L.b[k].v, L.b[k].l))); };
    auto dual_sb = [&](int k) { return wk * S.cinv * C.Eb[k] * (L.rho_b * (L.b[k].vb - proj_sb(L, L.b[k].vb))); };
    if (L.is_core) {
      const int j = L.gl;
      const double xc = C.D * L.x;
      if (!ID) {
        const double cmd = ok ? xc : 0.0;
        if (io.out) io.out[b * io.sout.sb + j * io.sout.sk] = cmd;
        if (io.roll_q) {   // rollout: integrate in place
          io.roll_q[b * io.sroll.sb + j * io.sroll.sk] += io.roll_dt * cmd;
          io.roll_qd[b * io.sroll.sb + j * io.sroll.sk] = cmd;
        }
      } else if (io.out2) io.out2[b * io.sout2.sb + j * io.sout2.sk] = ok ? xc : 0.0;
      if (io.qp_x) {
        double* xr = io.qp_x + (long long)b * (Cfg::NC * (1 + Cfg::KU) + Cfg::NR);
        xr[j] = wk * xc;
#pragma unroll
        for (int k = 0; k < Cfg::KU; ++k) xr[Cfg::NC * (1 + k) + j] = C.has_sing[k] ? wk * (C.Dd[k] * L.b[k].xd) : 0.0;
      }
      if (yr) {
        yr[j] = Cfg::BOUNDS ? wk * S.cinv * C.Ecb * (L.rhoc * (L.vc - clampd(L.vc, L.lc, L.uc))) : 0.0;
#pragma unroll
        for (int k = 0; k < Cfg::KU; ++k) {
          yr[Cfg::NC * (1 + k) + j] = C.active[k] ? dual_row(k) : 0.0;
          yr[Cfg::NC * (1 + Cfg::KU + k) + j] = C.has_sb[k] ? dual_sb(k) : 0.0;
        }
      }
    } else {
      const int r = L.gl - Cfg::NC;
      if (ID && r >= Cfg::ND) {
        const int j = r - Cfg::ND;
        io.out[b * io.sout.sb + j * io.sout.sk] = ok ? C.Dd[0] * L.b[0].xd : io.c_g[j * io.Bc + slot];
      }
      if (io.qp_x) io.qp_x[(long long)b * (Cfg::NC * (1 + Cfg::KU) + Cfg::NR) + Cfg::NC * (1 + Cfg::KU) + r] = C.has_sing[0] ? wk * (C.Dd[0] * L.b[0].xd) : 0.0;
      if (yr) {
        yr[Cfg::NC * (1 + 2 * Cfg::KU) + r] = dual_row(0);
        yr[Cfg::NC * (1 + 2 * Cfg::KU) + Cfg::NR + r] = C.has_sb[0] ? dual_sb(0) : 0.0;
      }
    }
  });
}

// lane -> (group, lane-in-group) assignment shared by the device kernel and the emulator
template <class Cfg>
DRC_HD void lane_assign(Lane<Cfg>& L, int lane) {
  const int g = lane / Cfg::GL;
  if (g < Cfg::NG) { L.grp = g; L.gl = lane - g * Cfg::GL; }
  else { L.grp = 0; L.gl = -1; }
  L.is_core = L.gl >= 0 && L.gl < Cfg::NC;
}

}  // namespace drc
