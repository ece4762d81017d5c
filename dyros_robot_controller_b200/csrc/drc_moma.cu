// drc_b200 -- mobile base + mobile manipulator entry points of the C ABI (include/drc_b200.h, "mobile manipulator").
// Replaces Mobile::RobotData (reference src/mobile/robot_data.cpp), MobileManipulator::RobotData
// (src/mobile_manipulator/robot_data.cpp), MobileManipulator::QPIK / QPID (QP_IK.cpp, QP_ID.cpp) and the
// RobotController wrappers (robot_controller.cpp:147-250).  Same kernels as the manipulator path (drc_kernels.cuh),
// instantiated for the tree-shaped whole-body models: NV = 3 virtual + W wheels + 7 arm joints.
#include "drc_host.h"

// whole-body robot shapes with a kernel instantiation: (NV, W) = (12, 2) Husky-FR3 class, (14, 4) XLS-FR3 class
#define DRC_DISPATCH_MOMA(d, CALL)                                                                  \
  if ((d).nv == 12 && (d).wheel_num == 2) { constexpr int NV = 12, W = 2; CALL; }                   \
  else if ((d).nv == 14 && (d).wheel_num == 4) { constexpr int NV = 14, W = 4; CALL; }              \
  else return fail(DRC_E_UNSUPPORTED, "no kernel instantiation for this mobile manipulator (dof / wheel count)");

static int check_moma(const drc_ctx* c, int B) {
  int rc = check_batch(c, B); if (rc) return rc;
  if (c->model->hm.dev.drive_type == kNoBase) return fail(DRC_E_INVALID, "the model has no mobile base (drc_model_attach_mobile_base)");
  if (!c->c_Mact) return fail(DRC_E_INVALID, "the context was created before the mobile base was attached");
  return DRC_OK;
}

constexpr unsigned K_STATE = F_DYN | F_STORE | F_FRAME_OUT | F_MANIP_OUT | F_GRADDOT | F_MOMA;
// QP-build jobs read the state the F_STORE job (or an earlier updateState) cached.  The whole-body QPIK record is kinematic; the QPID
// record reads M~ and g~ only.  Everything else updateState owes the cache (nle, the inverses -- the full-model M^-1 goes through the
// rank-revealing route for these models: 77 % of the old single kernel, ncu round 2) is left to the dynamics-only job K_DYN, which
// fused calls enqueue behind the solver launch on the low-priority stream.
constexpr unsigned K_IK = F_FROM_CACHE | F_QPIK | F_MOMA;
constexpr unsigned K_ID = F_DYN | F_DYN_LIGHT | F_FROM_CACHE | F_QPID | F_MOMA;
constexpr unsigned K_DYN = F_DYN | F_FROM_CACHE | F_MOMA;

// state update (+ optional getter outputs); q == null: from the cached state
template <int NV, int W>
static int moma_state(drc_ctx* c, int B, const double* q, const double* qd, int frame, double* pose, double* Jt, double* Jtd, double* vel,
                      double* mani, double* mgrad, double* mgraddot, int layout, cudaStream_t s) {
  constexpr int ACT = NV - 3, MANI = NV - 3 - W;
  JobIO io;
  std::memset(&io, 0, sizeof io);
  io.B = B; io.q = q; io.qd = qd; io.sq = lay(layout, NV, B); io.sqd = io.sq;
  bind_cache(c, io);
  io.pose = pose; io.spose = lay(layout, 12, B); io.J = Jt; io.sJ = lay(layout, 6 * ACT, B); io.Jdot = Jtd; io.sJd = io.sJ;
  io.vel = vel; io.svel = lay(layout, 6, B);
  io.mani = mani; io.mani_grad = mgrad; io.smg = lay(layout, MANI, B); io.mani_graddot = mgraddot; io.smgd = io.smg;
  DrcFrame fr; std::memset(&fr, 0, sizeof fr); fr.parent = -1;
  if (frame >= 0) fr = frame_of(c->model, frame);
  return launch_dyn_job<NV, K_STATE, W>(c, fr, io, s);
}

template <int NV, int W>
static int moma_qp(drc_ctx* c, int B, bool id, const double* q, const double* qd, const double* x_target, const double* xdot, int frame,
                   double* out, double* out2, int* status, int* iters, int layout, cudaStream_t s) {
  constexpr int ACT = NV - 3, MANI = NV - 3 - W;
  const DrcModelDev& d = c->model->hm.dev;
  JobIO io;
  std::memset(&io, 0, sizeof io);
  io.B = B; io.q = q; io.qd = qd; io.sq = lay(layout, NV, B); io.sqd = io.sq;
  io.x_target = x_target; io.sxt = lay(layout, 12, B); io.xdot_target = xdot; io.sxd = lay(layout, 6, B);
  io.qp = c->qp;
  bind_cache(c, io);
  const DrcFrame fr = frame_of(c->model, frame);
  if (c->timing) { cudaEventRecord(c->ev[0], s); c->tr_n = 0; mark(c, "start", s); }
  int rc = DRC_OK;
  const bool sched = c->prm.schedule_hint != 0 && B >= 64;
  if (sched) { rc = launch_schedule(c, B, s); if (rc) return rc; mark(c, "schedule", s); }
  // stage 1: joint placements -> cache (fused calls), self-collision narrow phase; EPA pass on the side stream
  if (q) { rc = launch_job<NV, false, F_STORE>(c, fr, io, s); if (rc) return rc; mark(c, "fk", s); }
  CU(cudaEventRecord(c->ev_store, s));
  // stage 2 NEXT TO the narrow phase on its own stream (both only read the cached state): whole-body kinematics, manipulability and
  // the QP record except the self-collision row.  QPIK: manipulability in two routes as in the manipulator pipeline (Cholesky under
  // the conditioning certificate, the uncertified robots redone by the rank-revealing route in a follow-up launch over their list).
  // QPID keeps the rank-revealing route for every robot: its hard-constrained QP (KKT condition ~1e8) turns the last-digit
  // difference of the two routes into another ADMM path on single robots, and parity with the reference's path comes first.
  {
    cudaStream_t bs = c->build_stream;
    CU(cudaStreamWaitEvent(bs, c->ev_store, 0));
    if (c->late_pending) { CU(cudaStreamWaitEvent(bs, c->ev_late, 0)); c->late_pending = false; }   // targets still uploading (host entry points)
    JobIO bio = io;
    if (id) {
      rc = launch_job<NV, false, K_ID, W>(c, fr, bio, bs); if (rc) return rc;
    } else {
      bio.manip_list = c->manip_list; bio.manip_count = c->manip_count;
      CU(cudaMemsetAsync(c->manip_count, 0, sizeof(int), bs));
      rc = launch_job<NV, false, K_IK, W>(c, fr, bio, bs); if (rc) return rc;
      bio.redo = true;
      rc = launch_job<NV, false, K_IK, W>(c, fr, bio, bs); if (rc) return rc;
    }
    mark(c, "build", bs);
    CU(cudaEventRecord(c->ev_mbuild, bs));
  }
  CollisionIO cio;
  std::memset(&cio, 0, sizeof cio);
  cio.B = B; cio.mode = id ? 2 : 1; cio.qp = c->qp;
  cio.qp_stride = id ? MomaIdCfg<ACT>::STRIDE : MomaIkCfg<ACT>::STRIDE;
  cio.qp_row_off = (id ? MomaIdCfg<ACT>::OFF_ROW : MomaIkCfg<ACT>::OFF_ROW) + (ACT + 1);
  cio.row_n = ACT; cio.row_col0 = d.act_mani_start; cio.src0 = d.mani_start; cio.nsrc = MANI;
  rc = launch_collision<NV, false>(c, cio, s, true);
  if (rc) return rc;
  if (c->timing) { cudaEventRecord(c->ev[1], s); mark(c, "collision", s); }
  CU(cudaStreamWaitEvent(s, c->ev_mbuild, 0));
  if (c->timing) cudaEventRecord(c->ev[2], s);
  if (q) CU(cudaEventRecord(c->ev_solve, s));   // "the solver launch may start": the dynamics-only job becomes eligible with it
  SolveIO sio;
  std::memset(&sio, 0, sizeof sio);
  sio.B = B; sio.out = out; sio.sout = lay(layout, ACT, B); sio.out2 = out2; sio.sout2 = sio.sout; sio.status = status; sio.iters = iters;
  const unsigned mani_mask = ((1u << MANI) - 1u) << d.act_mani_start;   // CBF unit rows exist on manipulator joints only
  // robots whose self-collision row is still in the EPA pass are solved by a small launch behind that pass (side stream)
  rc = id ? solve_epa_robots<MomaIdCfg<ACT>, true>(c, sio, s, mani_mask, c->c_gact) : solve_epa_robots<MomaIkCfg<ACT>, false>(c, sio, s, mani_mask, c->c_gact);
  if (rc) return rc;
  sio.skip = c->epa_flag;
  if (sched) sio.order = c->order;
  mark(c, "admm_begin", s);   // admm - admm_begin = the main solver launch in situ (next to the EPA pass and the dynamics job)
  rc = id ? launch_admm<MomaIdCfg<ACT>, true>(c, sio, s, mani_mask, c->c_gact) : launch_admm<MomaIkCfg<ACT>, false>(c, sio, s, mani_mask, c->c_gact);
  if (rc) return rc;
  mark(c, "admm", s);
  if (q) {  // updateState's dynamics -> cache, behind the solver launch: its blocks fill the SMs the convergence tail leaves idle
    // (eligible together with the solver launch, not earlier: its long-lived 255-register blocks would otherwise sit on the SMs
    // while the narrow phase and the QP build -- the critical path -- wait for room)
    CU(cudaStreamWaitEvent(c->dyn_stream, c->ev_solve, 0));
    rc = launch_dyn_job<NV, K_DYN, W>(c, fr, io, c->dyn_stream); if (rc) return rc;
    mark(c, "dynamics", c->dyn_stream);
    CU(cudaEventRecord(c->ev_dyn, c->dyn_stream));
    CU(cudaStreamWaitEvent(s, c->ev_dyn, 0));
  }
  rc = join_epa(c, s);
  mark(c, "epa_robots", s);
  if (c->timing) { cudaEventRecord(c->ev[3], s); mark(c, "end", s); }
  return rc;
}

// Full-dof getters of a whole-body model (MobileManipulator::RobotData inherits Manipulator::RobotData's getJacobian /
// getJacobianTimeVariation / getPose / getVelocity / getMinDistance, mobile_manipulator/robot_data.h:42): the generic entry points
// drc_batch_get_frame / drc_batch_get_min_distance forward here when the model has a mobile base.  J, Jdot are 6 x dof.
int moma_get_frame_full(drc_ctx* c, int B, int frame, double* pose12, double* J, double* Jdot, double* vel, int layout, cudaStream_t s) {
  const DrcModelDev& d = c->model->hm.dev;
  JobIO io; std::memset(&io, 0, sizeof io);
  io.B = B; bind_cache(c, io);
  io.pose = pose12; io.spose = lay(layout, 12, B); io.J = J; io.sJ = lay(layout, 6 * d.nv, B); io.Jdot = Jdot; io.sJd = io.sJ;
  io.vel = vel; io.svel = lay(layout, 6, B);
  const DrcFrame fr = frame_of(c->model, frame);
  DRC_DISPATCH_MOMA(d, return (launch_job<NV, false, F_FROM_CACHE | F_FRAME_OUT>(c, fr, io, s)));
}
int moma_get_min_distance(drc_ctx* c, int B, int with_graddot, double* dist, double* grad, double* grad_dot, int* pair, int layout, cudaStream_t s) {
  const DrcModelDev& d = c->model->hm.dev;
  CollisionIO io; std::memset(&io, 0, sizeof io);
  io.B = B; io.mode = 0; io.dist = dist; io.grad = grad; io.sgrad = lay(layout, d.nv, B);
  io.grad_dot = with_graddot ? grad_dot : nullptr; io.sgd = io.sgrad; io.pair_out = pair;
  DRC_DISPATCH_MOMA(d, return (launch_collision<NV, false>(c, io, s)));
}

extern "C" {

// ---- Mobile::RobotData / MobileManipulator::RobotData constructors (mobile/robot_data.cpp:7-32, mobile_manipulator/robot_data.cpp:7-44)
int drc_model_attach_mobile_base(drc_model_t* m, int drive_type, double wheel_radius, double base_width, double wheel_offset, int n_wheels,
                                 const double* roller_angles, const double* b2w_x, const double* b2w_y, const double* b2w_angles,
                                 int virtual_start, int mani_start, int mobi_start, int act_mani_start, int act_mobi_start) {
  if (!m) return fail(DRC_E_INVALID, "null model");
  try {
    MobileParam p;
    p.drive_type = drive_type; p.wheel_radius = wheel_radius; p.base_width = base_width; p.wheel_offset = wheel_offset;
    const int np = drive_type == kCaster ? n_wheels / 2 : n_wheels;
    for (int i = 0; i < n_wheels && roller_angles; ++i) p.roller_angles.push_back(roller_angles[i]);
    for (int i = 0; i < np && b2w_x && b2w_y; ++i) { p.b2w_x.push_back(b2w_x[i]); p.b2w_y.push_back(b2w_y[i]); }
    for (int i = 0; i < n_wheels && b2w_angles; ++i) p.b2w_angles.push_back(b2w_angles[i]);
    attach_mobile_base(m->hm, p, virtual_start, mani_start, mobi_start, act_mani_start, act_mobi_start);
    return DRC_OK;
  } catch (const std::exception& e) {
    return fail(DRC_E_INVALID, e.what());
  }
}
int drc_model_moma_info(const drc_model_t* m, int* s) {
  if (!m || !s) return fail(DRC_E_INVALID, "null argument");
  const DrcModelDev& d = m->hm.dev;
  s[0] = d.drive_type; s[1] = d.wheel_num; s[2] = d.mani_dof; s[3] = d.drive_type == kNoBase ? d.nv : d.wheel_num + d.mani_dof;
  return DRC_OK;
}
// Mobile::RobotData::getFKJacobian for differential / mecanum drives (configuration independent, mobile/robot_data.cpp:138-177)
int drc_model_base_jacobian(const drc_model_t* m, double* J) {
  if (!m || !J) return fail(DRC_E_INVALID, "null argument");
  const DrcModelDev& d = m->hm.dev;
  if (d.drive_type == kNoBase) return fail(DRC_E_INVALID, "the model has no mobile base");
  if (d.drive_type == kCaster) return fail(DRC_E_UNSUPPORTED, "caster base Jacobian depends on the steering angles: use drc_batch_mobile_fk");
  for (int r = 0; r < 3; ++r) for (int k = 0; k < d.wheel_num; ++k) J[r * d.wheel_num + k] = d.J_mobile[r][k];
  return DRC_OK;
}

// ---- device entry points
int drc_batch_moma_update_state(drc_ctx_t* c, int B, const double* q, const double* qdot, int layout, void* stream) {
  int rc = check_moma(c, B); if (rc) return rc;
  if (!q || !qdot) return fail(DRC_E_INVALID, "null state pointer");
  CU(cudaSetDevice(c->device));
  DRC_DISPATCH_MOMA(c->model->hm.dev, return (moma_state<NV, W>(c, B, q, qdot, -1, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, layout, pick(c, stream))));
}
int drc_batch_moma_get_state(drc_ctx_t* c, int B, int frame, double* pose12, double* J_act, double* Jdot_act, double* vel, double* M_act,
                             double* Minv_act, double* g_act, double* nle_act, double* mani, double* mani_grad, double* mani_graddot,
                             int layout, void* stream) {
  int rc = check_moma(c, B); if (rc) return rc;
  rc = check_frame(c, frame); if (rc) return rc;
  CU(cudaSetDevice(c->device));
  cudaStream_t s = pick(c, stream);
  const DrcModelDev& d = c->model->hm.dev;
  const int a = d.wheel_num + d.mani_dof;
  {
    DRC_DISPATCH_MOMA(d, rc = (moma_state<NV, W>(c, B, nullptr, nullptr, frame, pose12, J_act, Jdot_act, vel, mani, mani_grad, mani_graddot, layout, s)));
  }
  if (rc) return rc;
  auto cp = [&](const double* src, int K, double* dst) -> int {
    if (!dst) return DRC_OK;
    const long long tot = (long long)K * B;
    k_copy_cache<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(src, nullptr, c->cap, K, B, dst, lay(layout, K, B));
    c->launches++;
    CU(cudaGetLastError());
    return DRC_OK;
  };
  if ((rc = cp(c->c_Mact, a * a, M_act))) return rc;
  if ((rc = cp(c->c_Minvact, a * a, Minv_act))) return rc;
  if ((rc = cp(c->c_gact, a, g_act))) return rc;
  return cp(c->c_nleact, a, nle_act);
}

#define MOMA_QP_ENTRY(ID_, Q_, QD_, XT_, XD_, OUT_, OUT2_)                                                          \
  int rc = check_moma(c, B); if (rc) return rc;                                                                     \
  rc = check_frame(c, frame); if (rc) return rc;                                                                    \
  if (!(XD_) || !(OUT_)) return fail(DRC_E_INVALID, "null argument");                                               \
  CU(cudaSetDevice(c->device));                                                                                     \
  DRC_DISPATCH_MOMA(c->model->hm.dev, return (moma_qp<NV, W>(c, B, ID_, Q_, QD_, XT_, XD_, frame, OUT_, OUT2_, status, iters, layout, pick(c, stream))));

int drc_batch_moma_qpik(drc_ctx_t* c, int B, const double* xdot_des, int frame, double* eta_out, int* status, int* iters, int layout, void* stream) {
  MOMA_QP_ENTRY(false, nullptr, nullptr, nullptr, xdot_des, eta_out, nullptr)
}
int drc_batch_moma_qpik_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, int frame, double* eta_out, int* status, int* iters, int layout, void* stream) {
  if (!x_target) return fail(DRC_E_INVALID, "null target pose");
  MOMA_QP_ENTRY(false, nullptr, nullptr, x_target, xdot_target, eta_out, nullptr)
}
int drc_batch_moma_qpid(drc_ctx_t* c, int B, const double* xddot_des, int frame, double* tau_out, double* etadot_out, int* status, int* iters, int layout, void* stream) {
  MOMA_QP_ENTRY(true, nullptr, nullptr, nullptr, xddot_des, tau_out, etadot_out)
}
int drc_batch_moma_qpid_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, int frame, double* tau_out, double* etadot_out, int* status, int* iters, int layout, void* stream) {
  if (!x_target) return fail(DRC_E_INVALID, "null target pose");
  MOMA_QP_ENTRY(true, nullptr, nullptr, x_target, xdot_target, tau_out, etadot_out)
}
int drc_batch_moma_cycle_qpik_step(drc_ctx_t* c, int B, const double* q, const double* qdot, const double* x_target, const double* xdot_target, int frame, double* eta_out, int* status, int* iters, int layout, void* stream) {
  if (!q || !qdot || !x_target) return fail(DRC_E_INVALID, "null argument");
  MOMA_QP_ENTRY(false, q, qdot, x_target, xdot_target, eta_out, nullptr)
}
int drc_batch_moma_cycle_qpid_step(drc_ctx_t* c, int B, const double* q, const double* qdot, const double* x_target, const double* xdot_target, int frame, double* tau_out, double* etadot_out, int* status, int* iters, int layout, void* stream) {
  if (!q || !qdot || !x_target) return fail(DRC_E_INVALID, "null argument");
  MOMA_QP_ENTRY(true, q, qdot, x_target, xdot_target, tau_out, etadot_out)
}

// ---- host-buffer variants (AoS, synchronous)
#define MOMA_HOST_PRELUDE                                                        \
  HOST_PRELUDE                                                                   \
  const int a = c->model->hm.dev.wheel_num + c->model->hm.dev.mani_dof; (void)a; \
  const int k = c->model->hm.dev.mani_dof; (void)k;

int drc_host_moma_update_state(drc_ctx_t* c, int B, const double* q, const double* qdot) {
  MOMA_HOST_PRELUDE
  const double *dq = st.in(q, Bz * n), *dqd = st.in(qdot, Bz * n);
  return st.finish(st.err ? st.err : drc_batch_moma_update_state(c, B, dq, dqd, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_moma_get_state(drc_ctx_t* c, int B, int frame, double* pose12, double* J_act, double* Jdot_act, double* vel, double* M_act,
                            double* Minv_act, double* g_act, double* nle_act, double* mani, double* mani_grad, double* mani_graddot) {
  MOMA_HOST_PRELUDE
  double *dp = st.out(pose12, Bz * 12), *dJ = st.out(J_act, Bz * 6 * a), *dJd = st.out(Jdot_act, Bz * 6 * a), *dv = st.out(vel, Bz * 6);
  double *dM = st.out(M_act, Bz * a * a), *dMi = st.out(Minv_act, Bz * a * a), *dg = st.out(g_act, Bz * a), *dn = st.out(nle_act, Bz * a);
  double *dm = st.out(mani, Bz), *dmg = st.out(mani_grad, Bz * k), *dmgd = st.out(mani_graddot, Bz * k);
  return st.finish(st.err ? st.err : drc_batch_moma_get_state(c, B, frame, dp, dJ, dJd, dv, dM, dMi, dg, dn, dm, dmg, dmgd, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_moma_qpik(drc_ctx_t* c, int B, const double* xdot_des, int frame, double* eta_out, int* status, int* iters) {
  MOMA_HOST_PRELUDE
  const double* dx = st.in(xdot_des, Bz * 6);
  double* dout = st.out(eta_out, Bz * a);
  int *ds = st.out_i(status, Bz), *di = st.out_i(iters, Bz);
  return st.finish(st.err ? st.err : drc_batch_moma_qpik(c, B, dx, frame, dout, ds, di, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_moma_qpik_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, int frame, double* eta_out, int* status, int* iters) {
  MOMA_HOST_PRELUDE
  const double *dxt = st.in(x_target, Bz * 12), *dx = st.in(xdot_target, Bz * 6);
  double* dout = st.out(eta_out, Bz * a);
  int *ds = st.out_i(status, Bz), *di = st.out_i(iters, Bz);
  return st.finish(st.err ? st.err : drc_batch_moma_qpik_step(c, B, dxt, dx, frame, dout, ds, di, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_moma_qpid(drc_ctx_t* c, int B, const double* xddot_des, int frame, double* tau_out, double* etadot_out, int* status, int* iters) {
  MOMA_HOST_PRELUDE
  const double* dx = st.in(xddot_des, Bz * 6);
  double *dout = st.out(tau_out, Bz * a), *dout2 = st.out(etadot_out, Bz * a);
  int *ds = st.out_i(status, Bz), *di = st.out_i(iters, Bz);
  return st.finish(st.err ? st.err : drc_batch_moma_qpid(c, B, dx, frame, dout, dout2, ds, di, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_moma_qpid_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, int frame, double* tau_out, double* etadot_out, int* status, int* iters) {
  MOMA_HOST_PRELUDE
  const double *dxt = st.in(x_target, Bz * 12), *dx = st.in(xdot_target, Bz * 6);
  double *dout = st.out(tau_out, Bz * a), *dout2 = st.out(etadot_out, Bz * a);
  int *ds = st.out_i(status, Bz), *di = st.out_i(iters, Bz);
  return st.finish(st.err ? st.err : drc_batch_moma_qpid_step(c, B, dxt, dx, frame, dout, dout2, ds, di, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_moma_cycle_qpik_step(drc_ctx_t* c, int B, const double* q, const double* qdot, const double* x_target, const double* xdot_target, int frame, double* eta_out, int* status, int* iters) {
  MOMA_HOST_PRELUDE
  // q, qdot feed stage 1 (joint placements, narrow phase); the targets are only read by the QP-build job: upload them behind stage 1
  const double *dq = st.in(q, Bz * n), *dqd = st.in(qdot, Bz * n), *dxt = st.in_late(x_target, Bz * 12), *dx = st.in_late(xdot_target, Bz * 6);
  st.late_done();
  double* dout = st.out_direct(eta_out, Bz * a);
  int *ds = st.out_i_direct(status, Bz), *di = st.out_i_direct(iters, Bz);
  return st.finish(st.err ? st.err : drc_batch_moma_cycle_qpik_step(c, B, dq, dqd, dxt, dx, frame, dout, ds, di, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_moma_cycle_qpid_step(drc_ctx_t* c, int B, const double* q, const double* qdot, const double* x_target, const double* xdot_target, int frame, double* tau_out, double* etadot_out, int* status, int* iters) {
  MOMA_HOST_PRELUDE
  const double *dq = st.in(q, Bz * n), *dqd = st.in(qdot, Bz * n), *dxt = st.in_late(x_target, Bz * 12), *dx = st.in_late(xdot_target, Bz * 6);
  st.late_done();
  double *dout = st.out_direct(tau_out, Bz * a), *dout2 = st.out_direct(etadot_out, Bz * a);
  int *ds = st.out_i_direct(status, Bz), *di = st.out_i_direct(iters, Bz);
  return st.finish(st.err ? st.err : drc_batch_moma_cycle_qpid_step(c, B, dq, dqd, dxt, dx, frame, dout, dout2, ds, di, DRC_LAYOUT_AOS, nullptr));
}

}  // extern "C"
