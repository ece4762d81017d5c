// drc_b200 -- CUDA kernels (sm_100a) and the C ABI declared in include/drc_b200.h.
//
// Kernels (one launch each, all fp64):
//   k_robot_job<NV,CHAIN,FLAGS>   one robot per thread: state update, frame quantities, QP records
//   k_collision / k_collision_epa one robot per thread: min self-distance, gradients, QP row
//   k_admm<Cfg,ID>                GL lanes per robot (NG robots per warp): OSQP-algorithm ADMM
//   k_task_cubic                  one robot per thread: cubic task-space trajectory
//   k_copy_cache                  cache (SoA) -> user layout
// The robot model travels as a __grid_constant__ kernel parameter (constant bank, broadcast reads).
// There is NO CPU fallback: every entry point that computes requires a CUDA device.
#include "drc_host.h"

static thread_local std::string g_err;
int drc_set_error(int code, const std::string& msg) { g_err = msg; return code; }

// One QP controller call on the cached state (or fused with the state update when q != null).
//
// Pipeline (main stream s):  [schedule] -> FK store -> narrow phase -> state / QP build (EPA pass next to it on the side
// stream) -> ADMM over the robots in schedule order.  Fused calls on large batches also run a PRIORITY pipeline on a
// high-priority stream: the robots whose previous tick needed >= kPrioIters ADMM iterations (a per-cent of the batch,
// but their 1000-4000 serial iterations are the critical path of the whole call) go through the same stages in a small
// compact scratch, so that their ADMM starts ~0.5 ms after the call instead of behind the full batch's stages.
template <int NV, bool CHAIN>
static int run_qp(drc_ctx* c, int B, bool id, bool step, const double* q, const double* qd, const double* x_target,
                  const double* xdot, int frame, double* out, double* out2, int* status, int* iters, int layout, cudaStream_t s) {
  JobIO io;
  std::memset(&io, 0, sizeof io);
  io.B = B; io.q = q; io.qd = qd; io.sq = lay(layout, NV, B); io.sqd = io.sq;
  io.x_target = x_target; io.sxt = lay(layout, 12, B); io.xdot_target = xdot; io.sxd = lay(layout, 6, B);
  io.qp = c->qp;
  bind_cache(c, io);
  const DrcFrame fr = frame_of(c->model, frame);
  const bool fused = q != nullptr;
  // Fused QPIK cycles run on the context's own streams, whose priorities order the block dispatch: priority pipeline >
  // main pipeline > dynamics-only kernel.  A caller stream is joined by events on both sides.
  cudaStream_t caller = s;
  if (fused && !id && s != c->stream) {
    CU(cudaEventRecord(c->ev_in, caller));
    CU(cudaStreamWaitEvent(c->stream, c->ev_in, 0));
    s = c->stream;
  }
  if (c->timing) { cudaEventRecord(c->ev[0], s); c->tr_n = 0; mark(c, "start", s); }
  int rc;
  const int qp_stride = id ? QpidCfg<NV>::STRIDE : QpikCfg<NV>::STRIDE;
  const int qp_row_off = (id ? QpidCfg<NV>::OFF_ROW : QpikCfg<NV>::OFF_ROW) + (NV + 1);
  const bool sched = c->prm.schedule_hint != 0 && B >= 64;
  const bool prio = sched && fused && B >= kPrioMinBatch;
  // The schedule is needed by the priority pipeline (its robot list) and by the ADMM launches, not by the main pipeline's front
  // stages: with a priority pipeline it is built on the priority stream, off the main pipeline's critical path.
  if (sched && !prio) { rc = launch_schedule(c, B, s); if (rc) return rc; mark(c, "schedule", s); }

#define J(FL, IO, ST) launch_job<NV, CHAIN, FL>(c, fr, IO, ST)
  // QPIK does not read the dynamics (M, M^-1, g, nle of updateState): in fused calls they leave the critical path -- the QP
  // record is built from the state stage 1 has just cached, and a dynamics-only kernel runs on its own stream behind the
  // ADMM launch, where it fills the SMs the convergence tail leaves idle (the call still ends with the full cache).
  const bool split_dyn = fused && !id;
  auto build = [&](const JobIO& jio, cudaStream_t st) -> int {  // state update, manipulability, QP record except the collision row
    if (!id) return step ? J(F_FROM_CACHE | F_QPIK | F_STEP, jio, st) : J(F_FROM_CACHE | F_QPIK, jio, st);
    // (fused: from the state stage 1a has just cached -- the job runs next to the narrow phase, which reads those placements)
    if (fused) return step ? J(F_DYN | F_FROM_CACHE | F_QPID | F_STEP, jio, st) : J(F_DYN | F_FROM_CACHE | F_QPID, jio, st);
    return step ? J(F_FROM_CACHE | F_QPID | F_STEP, jio, st) : J(F_FROM_CACHE | F_QPID, jio, st);
  };

  if (prio) CU(cudaEventRecord(c->ev_sched, s));
  // stage 1a (main pipeline): joint placements -> cache.  Enqueued before the priority pipeline so that every read of the INPUT
  // state arrays by the main pipeline precedes the priority solver launch, which may update them in place (rollouts)
  if (fused) { rc = J(F_STORE, io, s); if (rc) return rc; mark(c, "fk", s); }
  if (fused) CU(cudaEventRecord(c->ev_store, s));
  if (prio) {
    // ---- priority pipeline: slots [0, *slow_count) of the compact scratch hold robots order[slot]
    CU(cudaStreamWaitEvent(c->prio_stream, c->ev_sched, 0));
    cudaStream_t ps = c->prio_stream;
    rc = launch_schedule(c, B, ps); if (rc) return rc;
    mark(c, "schedule", ps);
    CU(cudaEventRecord(c->ev_order, ps));
    JobIO pio = io;
    pio.B = kPrioSlots; pio.ids = c->order; pio.count = c->slow_count;
    bind_scratch(c->prio, pio);
    rc = J(F_STORE, pio, ps); if (rc) return rc;
    mark(c, "prio_fk", ps);
    // the QP-build kernel only needs the cached joint placements: it runs NEXT TO the narrow phase on a second high-priority
    // stream (both are tiny, latency-bound launches here), and the solver waits for both
    CU(cudaEventRecord(c->ev_prio_fk, ps));
    CU(cudaStreamWaitEvent(c->prio_side, c->ev_prio_fk, 0));
    if (c->late_pending) CU(cudaStreamWaitEvent(c->prio_side, c->ev_late, 0));   // x_target / xdot_target still uploading (host entry points)
    rc = build(pio, c->prio_side); if (rc) return rc;
    mark(c, "prio_build", c->prio_side);
    CU(cudaEventRecord(c->ev_prio_build, c->prio_side));
    CollisionIO pc;
    std::memset(&pc, 0, sizeof pc);
    pc.B = kPrioSlots; pc.mode = id ? 2 : 1; pc.qp = c->prio.qp; pc.qp_stride = qp_stride; pc.qp_row_off = qp_row_off;
    pc.count = c->slow_count;
    rc = launch_collision<NV, CHAIN>(c, pc, ps, false, &c->prio); if (rc) return rc;
    mark(c, "prio_collision", ps);
    CU(cudaStreamWaitEvent(ps, c->ev_prio_build, 0));
    CU(cudaStreamWaitEvent(ps, c->ev_store, 0));
    SolveIO ps_io;
    std::memset(&ps_io, 0, sizeof ps_io);
    ps_io.B = kPrioSlots; ps_io.out = out; ps_io.sout = lay(layout, NV, B); ps_io.out2 = out2; ps_io.sout2 = ps_io.sout;
    ps_io.status = status; ps_io.iters = iters; ps_io.out_ids = c->order; ps_io.count = c->slow_count;
    rc = id ? launch_admm<QpidCfg<NV>, true>(c, ps_io, ps, (1u << NV) - 1u, nullptr, &c->prio)
            : launch_admm<QpikCfg<NV>, false>(c, ps_io, ps, (1u << NV) - 1u, nullptr, &c->prio);
    if (rc) return rc;
    mark(c, "prio_admm", ps);
    CU(cudaEventRecord(c->ev_prio, ps));
  }

  // ---- main pipeline
  const bool par_build = c->par_build && fused && prio;
  if (par_build) {
    // the QP build only needs the cached joint placements: it runs NEXT TO the narrow phase on its own stream.  (Round 1 rejected this
    // with the monolithic 255-register collision kernel; with the closed-form stage in a 128-register kernel and the GJK stage waiting
    // at its round barriers most of the time the two now share the SMs: front stages 1.19 -> 1.07 ms, same box A/B.)
    CU(cudaStreamWaitEvent(c->build_stream, c->ev_store, 0));
    if (c->late_pending) { CU(cudaStreamWaitEvent(c->build_stream, c->ev_late, 0)); c->late_pending = false; }
    // two-route manipulability: Cholesky under a conditioning certificate, the rest (~10-15 % of random states) redone by the
    // rank-revealing route in a follow-up launch over their list -- the 6 x 6 column-pivoted QR was 73 % of this kernel
    JobIO bio = io;
    if (id) {   // QPID: state update + dynamics + record in one job (inline manipulability fallback), still next to the narrow phase
      rc = build(bio, c->build_stream); if (rc) return rc;
    } else {
      bio.manip_list = c->manip_list; bio.manip_count = c->manip_count;
      CU(cudaMemsetAsync(c->manip_count, 0, sizeof(int), c->build_stream));
      rc = build(bio, c->build_stream); if (rc) return rc;
      bio.redo = true;
      rc = build(bio, c->build_stream); if (rc) return rc;
    }
    mark(c, "build", c->build_stream);
    CU(cudaEventRecord(c->ev_mbuild, c->build_stream));
  }
  // stage 1b: self-collision narrow phase; its EPA pass goes to the side stream
  CollisionIO cio;
  std::memset(&cio, 0, sizeof cio);
  cio.B = B; cio.mode = id ? 2 : 1; cio.qp = c->qp; cio.qp_stride = qp_stride; cio.qp_row_off = qp_row_off;
  rc = launch_collision<NV, CHAIN>(c, cio, s, true);
  if (rc) return rc;
  if (c->timing) { cudaEventRecord(c->ev[1], s); mark(c, "collision", s); }
  // stage 2 (next to the EPA pass)
  if (par_build) {
    CU(cudaStreamWaitEvent(s, c->ev_mbuild, 0));
  } else {
    if (c->late_pending) { CU(cudaStreamWaitEvent(s, c->ev_late, 0)); c->late_pending = false; }
    rc = build(io, s);
    if (rc) return rc;
    mark(c, "build", s);
  }
#undef J
  if (c->timing) cudaEventRecord(c->ev[2], s);
  SolveIO sio;
  std::memset(&sio, 0, sizeof sio);
  sio.B = B; sio.out = out; sio.sout = lay(layout, NV, B); sio.out2 = out2; sio.sout2 = sio.sout; sio.status = status; sio.iters = iters;
  // robots whose self-collision row is still in the EPA pass (side stream) are solved by a small launch behind that pass
  rc = id ? solve_epa_robots<QpidCfg<NV>, true>(c, sio, s, (1u << NV) - 1u, nullptr) : solve_epa_robots<QpikCfg<NV>, false>(c, sio, s, (1u << NV) - 1u, nullptr);
  if (rc) return rc;
  sio.skip = c->epa_flag;
  if (sched) sio.order = c->order;
  if (prio) { sio.order_off = c->slow_count; CU(cudaStreamWaitEvent(s, c->ev_order, 0)); }   // those robots are solved by the priority pipeline
  if (split_dyn) CU(cudaEventRecord(c->ev_solve, s));
  mark(c, "admm_begin", s);   // admm - admm_begin = the main solver launch in situ
  rc = id ? launch_admm<QpidCfg<NV>, true>(c, sio, s) : launch_admm<QpikCfg<NV>, false>(c, sio, s);
  if (rc) return rc;
  mark(c, "admm", s);
  if (split_dyn) {
    // On the low-priority stream, eligible TOGETHER WITH the main solver launch (ev_solve), not with the front stages: the host
    // enqueues everything long before the GPU gets there, so a job that only waited for the cached placements would start next to
    // the narrow phase and take SM room from the critical path (same-box A/B, profiles/: 2.264 -> 2.300e7 cycles/s); behind the
    // solver's dispatch its blocks fill what the convergence tail leaves idle.
    CU(cudaStreamWaitEvent(c->dyn_stream, c->ev_solve, 0));
    rc = launch_job<NV, CHAIN, F_DYN | F_FROM_CACHE>(c, fr, io, c->dyn_stream); if (rc) return rc;
    mark(c, "dynamics", c->dyn_stream);
    CU(cudaEventRecord(c->ev_dyn, c->dyn_stream));
    CU(cudaStreamWaitEvent(s, c->ev_dyn, 0));
  }
  rc = join_epa(c, s);   // the EPA-pending robots' launch on the side stream
  if (rc) return rc;
  mark(c, "epa_robots", s);
  if (prio) CU(cudaStreamWaitEvent(s, c->ev_prio, 0));
  if (c->timing) { cudaEventRecord(c->ev[3], s); mark(c, "end", s); }
  if (caller != s) {
    CU(cudaEventRecord(c->ev_out, s));
    CU(cudaStreamWaitEvent(caller, c->ev_out, 0));
  }
  return rc;
}

// Closed-loop rollout: T control ticks of updateState + QPIKCubic / QPIKStep + the caller's integrate step
// (examples/C++/src/fr3_controller.cpp:116-131), TWO launches per tick:
//   k_tick_front  [schedule scatter] -> [cubic profile] -> joint placements -> narrow phase (+ EPA for the block's flagged robots)
//                 -> QPIK record
//   k_admm        solve -> command -> q += dt qdot*, qdot = qdot* in place -> tallies -> histogram + offsets of the next tick's schedule
// The robot -> warp schedule of tick k is tick k-1's iteration counts (tick 0: the context's previous call), results unaffected.
template <int NV, bool CHAIN>
static int run_rollout(drc_ctx* c, int B, int T, double dt, double* q, double* qdot, const double* x_target, const double* xdot_target,
                       const double* x_init, const double* xdot_init, double t_start, double t0, double duration, int frame,
                       int* fail_ticks, int* iters_total, int layout, cudaStream_t s) {
  const bool cubic = duration > 0;
  if (!c->roll) {
    CU(cudaMalloc((void**)&c->roll, (size_t)c->cap * (12 + 6) * sizeof(double)));
    CU(cudaMalloc((void**)&c->roll_i, (size_t)(2 * kSchedBuckets + 1) * sizeof(int)));
    CU(cudaMemset(c->roll_i, 0, (size_t)(2 * kSchedBuckets + 1) * sizeof(int)));
  }
  double *x_des = c->roll, *xd_des = c->roll + (size_t)c->cap * 12;
  int *hist_next = c->roll_i, *offs_next = c->roll_i + kSchedBuckets, *ticket = c->roll_i + 2 * kSchedBuckets;
  if (fail_ticks) CU(cudaMemsetAsync(fail_ticks, 0, (size_t)B * sizeof(int), s));
  if (iters_total) CU(cudaMemsetAsync(iters_total, 0, (size_t)B * sizeof(int), s));
  // warm start (extension, drc_params_t::rollout_warm_start): the solver launches read the previous tick's (x, y) of every robot and
  // leave this tick's there; zeros = cold start, which is how tick 0 begins
  constexpr int NX = NV * 3 + 2, NY = NV * 5 + 4;   // QpikCfg<NV>: NC (1 + KU) + NR, NC (1 + 2 KU) + 2 NR
  struct WarmGuard { drc_ctx* c; ~WarmGuard() { c->warm_on = false; } } warm_guard{c};
  if (c->prm.rollout_warm_start) {
    if (!c->ws_x) {
      CU(cudaMalloc((void**)&c->ws_x, (size_t)c->cap * NX * sizeof(double)));
      CU(cudaMalloc((void**)&c->ws_y, (size_t)c->cap * NY * sizeof(double)));
    }
    CU(cudaMemsetAsync(c->ws_x, 0, (size_t)B * NX * sizeof(double), s));
    CU(cudaMemsetAsync(c->ws_y, 0, (size_t)B * NY * sizeof(double), s));
    c->warm_on = true;
  }
  if (!c->prm.rollout_fused) {
    // pipeline variant: per tick [cubic profile kernel] + the multi-stream pipeline of the fused cycle (run_qp), whose solver launches
    // (priority, main, EPA-pending) integrate the state in place -- the slow robots' solves overlap the other robots' front stages
    c->roll_q = q; c->roll_qd = qdot; c->sroll = lay(layout, NV, B); c->roll_dt = dt; c->roll_fail = fail_ticks; c->roll_iters = iters_total;
    int rc = DRC_OK;
    for (int k = 0; k < T && rc == DRC_OK; ++k) {
      const double *xt = x_target, *xd = xdot_target;
      if (cubic) {
        rc = drc_batch_task_space_cubic(c, B, x_target, xdot_target, x_init, xdot_init, t_start + k * dt, t0, duration, x_des, xd_des, layout, s);
        xt = x_des; xd = xd_des;
      }
      if (rc == DRC_OK) rc = run_qp<NV, CHAIN>(c, B, false, true, q, qdot, xt, xd, frame, nullptr, nullptr, nullptr, nullptr, layout, s);
    }
    c->roll_q = c->roll_qd = nullptr; c->roll_fail = c->roll_iters = nullptr;
    return rc;
  }
  const bool sched = c->prm.schedule_hint != 0 && B >= 64;
  if (sched) { int rc = launch_schedule(c, B, s); if (rc) return rc; }
  const DrcFrame fr = frame_of(c->model, frame);
  const Scratch sc = main_scratch(c);
  for (int k = 0; k < T; ++k) {
    TickIO io;
    std::memset(&io, 0, sizeof io);
    JobIO& j = io.job;
    j.B = B; j.q = q; j.qd = qdot; j.sq = lay(layout, NV, B); j.sqd = j.sq;
    j.x_target = cubic ? x_des : x_target; j.sxt = lay(layout, 12, B);
    j.xdot_target = cubic ? xd_des : xdot_target; j.sxd = lay(layout, 6, B);
    j.qp = c->qp;
    bind_cache(c, j);
    CollisionIO& col = io.col;
    col.B = B; col.mode = 1; col.qp = c->qp; col.qp_stride = QpikCfg<NV>::STRIDE; col.qp_row_off = QpikCfg<NV>::OFF_ROW + (NV + 1);
    col.c_q = sc.c_q; col.c_qd = sc.c_qd; col.c_oMi = sc.c_oMi; col.Bc = sc.Bc;
    col.epa_flag = sc.epa_flag; col.cand_mask = sc.cand_mask; col.dist = sc.col_dist; col.pair_out = sc.col_pair; col.witness = sc.col_wit;
    if (cubic) {  // DyrosMath::getTaskSpaceCubic at the tick's time (QPIKCubic, robot_controller.cpp:302-317)
      CubicIO& cb = io.cubic;
      cb.B = B; cb.x_target = x_target; cb.xdot_target = xdot_target; cb.x_init = x_init; cb.xdot_init = xdot_init;
      cb.s12 = lay(layout, 12, B); cb.s6 = lay(layout, 6, B); cb.t = t_start + k * dt; cb.t0 = t0; cb.dur = duration; cb.x_des = x_des; cb.xdot_des = xd_des;
    }
    if (sched) { io.order = c->order; io.offs = offs_next; io.prev = k > 0 ? c->prev_iters : nullptr; }
    k_tick_front<NV, CHAIN><<<(B + kTickThreads - 1) / kTickThreads, kTickThreads, 0, s>>>(c->mdev, c->prm, fr, io);
    c->launches++;
    CU(cudaGetLastError());
    SolveIO sio;
    std::memset(&sio, 0, sizeof sio);
    sio.B = B; sio.sout = lay(layout, NV, B);
    sio.roll_q = q; sio.roll_qd = qdot; sio.sroll = lay(layout, NV, B); sio.roll_dt = dt; sio.fail_ticks = fail_ticks; sio.iters_total = iters_total;
    if (sched) { sio.order = c->order; sio.hist_next = hist_next; sio.offs_next = offs_next; sio.sched_ticket = ticket; }
    int rc = launch_admm<QpikCfg<NV>, false>(c, sio, s);
    if (rc) return rc;
  }
  return DRC_OK;
}

extern "C" {

const char* drc_last_error(void) { return g_err.c_str(); }
int drc_version(void) { return 100; }
int drc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// ------------------------------------------------------------------------------------------------ model
static int finish_model(std::unique_ptr<drc_model>& m, drc_model_t** out) {
  const DrcModelDev& d = m->hm.dev;
  char line[160];
  std::string v = "Total nq = " + std::to_string(d.nv) + "\nTotal nv = " + std::to_string(d.nv) + "\n\n";
  v += " id | name                 | nq | nv | idx_q | idx_v\n----+----------------------+----+----+-------+------\n";
  for (int i = 0; i < d.nv; ++i) {
    std::snprintf(line, sizeof line, "%3d | %20s | %2d | %2d | %5d | %4d\n", i + 1, m->hm.joint_names[i].c_str(), 1, 1, i, i);
    v += line;
  }
  m->verbose = v;
  m->hm.bind_hull();
  *out = m.release();
  return DRC_OK;
}
int drc_model_create_from_text(const char* urdf_text, const char* srdf_text, drc_model_t** out) {
  if (!urdf_text || !out) return fail(DRC_E_INVALID, "null argument");
  try {
    std::unique_ptr<drc_model> m(new drc_model);
    m->hm = compile_model(urdf_text, srdf_text ? srdf_text : "");
    return finish_model(m, out);
  } catch (const std::exception& e) {
    return fail(DRC_E_PARSE, e.what());
  }
}
static std::string dir_of(const std::string& path) {
  const size_t sl = path.find_last_of('/');
  return sl == std::string::npos ? std::string(".") : path.substr(0, sl);
}
int drc_model_create_from_urdf(const char* urdf_path, const char* srdf_path, const char* packages_path, drc_model_t** out) {
  if (!urdf_path || !out) return fail(DRC_E_INVALID, "null argument");
  std::string urdf, srdf;
  try {
    urdf = read_text_file(urdf_path);
  } catch (const std::exception& e) {
    return fail(DRC_E_IO, std::string("URDF file does not exist: ") + urdf_path);
  }
  if (srdf_path && srdf_path[0]) {
    try { srdf = read_text_file(srdf_path); } catch (const std::exception&) { srdf.clear(); }  // all pairs stay enabled
  }
  try {
    // mesh collision elements: package:// under packages_path (robot_data.cpp:24-28), relative names next to the URDF
    MeshSource ms;
    ms.urdf_dir = dir_of(urdf_path);
    ms.packages_path = packages_path ? packages_path : "";
    std::unique_ptr<drc_model> m(new drc_model);
    m->hm = compile_model(urdf, srdf, ms);
    return finish_model(m, out);
  } catch (const std::exception& e) {
    return fail(DRC_E_PARSE, e.what());
  }
}
void drc_model_destroy(drc_model_t* m) { delete m; }
int drc_model_dof(const drc_model_t* m) { return m ? m->hm.dev.nv : DRC_E_INVALID; }
int drc_model_frame_id(const drc_model_t* m, const char* link) { return (m && link) ? m->hm.frame_id(link) : -1; }
int drc_model_num_frames(const drc_model_t* m) { return m ? (int)m->hm.frames.size() : DRC_E_INVALID; }
const char* drc_model_frame_name(const drc_model_t* m, int f) {
  return (m && f >= 0 && f < (int)m->hm.frames.size()) ? m->hm.frames[f].name.c_str() : "";
}
const char* drc_model_joint_name(const drc_model_t* m, int j) {
  return (m && j >= 0 && j < m->hm.dev.nv) ? m->hm.joint_names[j].c_str() : "";
}
int drc_model_limits(const drc_model_t* m, double* q_lo, double* q_hi, double* v_lim, double* effort) {
  if (!m) return fail(DRC_E_INVALID, "null model");
  for (int i = 0; i < m->hm.dev.nv; ++i) {
    if (q_lo) q_lo[i] = m->hm.dev.q_lo[i];
    if (q_hi) q_hi[i] = m->hm.dev.q_hi[i];
    if (v_lim) v_lim[i] = m->hm.dev.v_lim[i];
    if (effort) effort[i] = m->hm.effort[i];
  }
  return DRC_OK;
}
int drc_model_info(const drc_model_t* m, int* s) {
  if (!m || !s) return fail(DRC_E_INVALID, "null argument");
  const DrcModelDev& d = m->hm.dev;
  s[0] = d.nv; s[1] = d.ngeom; s[2] = d.npair; s[3] = d.ngroup; s[4] = (int)m->hm.frames.size(); s[5] = m->hm.skipped_geoms;
  return DRC_OK;
}
int drc_model_mesh_info(const drc_model_t* m, int* s) {
  if (!m || !s) return fail(DRC_E_INVALID, "null argument");
  s[0] = m->hm.mesh_geoms; s[1] = (int)m->hm.hull.size() / 3;
  return DRC_OK;
}
const char* drc_model_verbose(const drc_model_t* m) { return m ? m->verbose.c_str() : ""; }

// ------------------------------------------------------------------------------------------------ context
// allocations of drc_ctx_create; on any failure the caller releases whatever was created so far (drc_ctx_destroy skips nulls)
static int ctx_create_impl(drc_ctx* c, const drc_model_t* m, int device, int max_batch) {
  c->model = m; c->device = device; c->cap = max_batch;
  c->mdev = m->hm.dev;
  c->mdev.geom.hull = nullptr;
  if (!m->hm.hull.empty()) {  // mesh hull vertices (kConvex geometry) live in device memory of this context's GPU
    CU(cudaMalloc((void**)&c->hull_dev, m->hm.hull.size() * sizeof(double)));
    CU(cudaMemcpy(c->hull_dev, m->hm.hull.data(), m->hm.hull.size() * sizeof(double), cudaMemcpyHostToDevice));
    c->mdev.geom.hull = c->hull_dev;
  }
  c->prm = DrcParams();
  for (int i = 0; i < kMaxV; ++i) { c->prm.Kp_joint[i] = 400; c->prm.Kv_joint[i] = 40; }
  const int n = m->hm.dev.nv;
  const size_t B = (size_t)max_batch;
  int prio_lo = 0, prio_hi = 0;   // numerically lower = higher priority; lo is the default (lowest) level
  CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  const int prio_main = prio_hi < prio_lo ? prio_lo - 1 : prio_lo;   // one level above the dynamics-only stream
  CU(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_main));
  CU(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, prio_main));
  CU(cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_last, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_out, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_col, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_epa, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_build, cudaEventDisableTiming));
  auto dalloc = [&](double** p, size_t cnt) { return cudaMalloc((void**)p, cnt * sizeof(double)); };
  CU(dalloc(&c->c_q, n * B)); CU(dalloc(&c->c_qd, n * B)); CU(dalloc(&c->c_oMi, 12 * n * B));
  CU(dalloc(&c->c_M, n * n * B)); CU(dalloc(&c->c_Minv, n * n * B)); CU(dalloc(&c->c_g, n * B)); CU(dalloc(&c->c_nle, n * B));
  if (m->hm.dev.drive_type != kNoBase) {  // mobile manipulator: actuated-space dynamics (attach the base BEFORE creating contexts)
    const int a = n - 3;
    CU(dalloc(&c->c_Mact, a * a * B)); CU(dalloc(&c->c_Minvact, a * a * B)); CU(dalloc(&c->c_gact, a * B)); CU(dalloc(&c->c_nleact, a * B));
  }
  // largest QP record among the four formulations for this dof
  const int nc = n;
  const int stride_id = nc * (nc + 1) / 2 + 3 * nc + 4 * nc + (2 + nc) * (nc + 1);
  c->qp_stride_max = stride_id;
  CU(dalloc(&c->qp, (size_t)stride_id * B));
  CU(cudaMalloc((void**)&c->epa_flag, B * sizeof(int)));
  CU(cudaMalloc((void**)&c->cand_mask, B * sizeof(unsigned long long)));
  CU(dalloc(&c->col_dist, B));
  CU(cudaMalloc((void**)&c->col_pair, B * sizeof(int)));
  CU(dalloc(&c->col_wit, 6 * B));
  CU(cudaMalloc((void**)&c->pinv_list, B * sizeof(int)));
  CU(cudaMalloc((void**)&c->pinv_count, sizeof(int)));
  CU(cudaMalloc((void**)&c->manip_list, B * sizeof(int)));
  CU(cudaMalloc((void**)&c->manip_count, sizeof(int)));
  CU(cudaMalloc((void**)&c->nar_k, B * sizeof(int)));
  CU(cudaMalloc((void**)&c->nar_lb, 64 * B * sizeof(float)));
  CU(cudaMalloc((void**)&c->epa_list, B * sizeof(int)));
  CU(cudaMalloc((void**)&c->epa_count, sizeof(int)));
  CU(cudaMalloc((void**)&c->prev_iters, B * sizeof(int)));
  CU(cudaMemset(c->prev_iters, 0, B * sizeof(int)));
  CU(cudaMalloc((void**)&c->order, B * sizeof(int)));
  CU(cudaMalloc((void**)&c->sched_hist, kSchedBuckets * sizeof(int)));
  CU(cudaMalloc((void**)&c->slow_count, sizeof(int)));
  CU(cudaMemset(c->slow_count, 0, sizeof(int)));
  {  // compact scratch of the priority pipeline and of the single-kernel cycle of the slowest robots
    auto alloc_scratch = [&](Scratch& p, size_t P) -> int {
      p.Bc = (long long)P;
      CU(dalloc(&p.c_q, n * P)); CU(dalloc(&p.c_qd, n * P)); CU(dalloc(&p.c_oMi, 12 * n * P)); CU(dalloc(&p.c_M, n * n * P));
      CU(dalloc(&p.c_Minv, n * n * P)); CU(dalloc(&p.c_g, n * P)); CU(dalloc(&p.c_nle, n * P));
      CU(dalloc(&p.qp, (size_t)stride_id * P));
      CU(cudaMalloc((void**)&p.epa_flag, P * sizeof(int)));
      CU(cudaMalloc((void**)&p.cand_mask, P * sizeof(unsigned long long)));
      CU(dalloc(&p.col_dist, P));
      CU(cudaMalloc((void**)&p.col_pair, P * sizeof(int)));
      CU(dalloc(&p.col_wit, 6 * P));
      CU(cudaMalloc((void**)&p.nar_k, P * sizeof(int)));
      CU(cudaMalloc((void**)&p.nar_lb, 64 * P * sizeof(float)));
      CU(cudaMalloc((void**)&p.epa_list, P * sizeof(int)));
      CU(cudaMalloc((void**)&p.epa_count, sizeof(int)));
      return DRC_OK;
    };
    int arc = alloc_scratch(c->prio, kPrioSlots); if (arc) return arc;
    int lo = 0, hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CU(cudaStreamCreateWithPriority(&c->prio_stream, cudaStreamNonBlocking, hi));
    CU(cudaStreamCreateWithPriority(&c->prio_side, cudaStreamNonBlocking, hi));
    CU(cudaStreamCreateWithPriority(&c->build_stream, cudaStreamNonBlocking, prio_main));
    CU(cudaEventCreateWithFlags(&c->ev_mbuild, cudaEventDisableTiming));
    c->par_build = true;
    CU(cudaEventCreateWithFlags(&c->ev_prio_fk, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_prio_build, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_sched, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_order, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_prio, cudaEventDisableTiming));
    CU(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithPriority(&c->dyn_stream, cudaStreamNonBlocking, prio_lo));
    CU(cudaEventCreateWithFlags(&c->ev_store, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_dyn, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_solve, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_late, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_early, cudaEventDisableTiming));
  }
  {
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
  }
  // staging: enough for the largest host call (q, qd, pose, xdot in; J/Jdot/M outs)
  c->stage_doubles = B * (size_t)(2 * n + 12 + 6 + 4 * n + 12 * n + 2 * n * n + 64);
  CU(dalloc(&c->stage, c->stage_doubles));
  c->stage_ints = 2 * B;
  CU(cudaMalloc((void**)&c->stage_i, c->stage_ints * sizeof(int)));
  for (int i = 0; i < 4; ++i) CU(cudaEventCreate(&c->ev[i]));
  return DRC_OK;
}
int drc_ctx_create(const drc_model_t* m, int device, int max_batch, drc_ctx_t** out) {
  if (!m || !out || max_batch <= 0) return fail(DRC_E_INVALID, "bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(DRC_E_CUDA, "no CUDA device: drc_b200 has no CPU fallback");
  }
  if (device < 0 || device >= ndev) return fail(DRC_E_INVALID, "device index out of range");
  CU(cudaSetDevice(device));
  drc_ctx* c = new drc_ctx;
  std::memset(c, 0, sizeof(drc_ctx));
  const int rc = ctx_create_impl(c, m, device, max_batch);
  if (rc != DRC_OK) {  // e.g. out of memory for a large batch: free the partial context so that the caller can retry smaller
    const std::string msg = g_err;
    cudaGetLastError();
    drc_ctx_destroy(c);
    g_err = msg;
    return rc;
  }
  *out = c;
  return DRC_OK;
}
void drc_ctx_destroy(drc_ctx_t* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  double* ds[] = {c->c_q, c->c_qd, c->c_oMi, c->c_M, c->c_Minv, c->c_g, c->c_nle, c->qp, c->col_dist, c->col_wit, c->stage,
                  c->c_Mact, c->c_Minvact, c->c_gact, c->c_nleact};
  for (double* p : ds) if (p) cudaFree(p);
  if (c->hull_dev) cudaFree(c->hull_dev);
  if (c->epa_flag) cudaFree(c->epa_flag);
  if (c->cand_mask) cudaFree(c->cand_mask);
  if (c->col_pair) cudaFree(c->col_pair);
  if (c->pinv_list) cudaFree(c->pinv_list);
  if (c->pinv_count) cudaFree(c->pinv_count);
  if (c->manip_list) cudaFree(c->manip_list);
  if (c->manip_count) cudaFree(c->manip_count);
  if (c->nar_k) cudaFree(c->nar_k);
  if (c->nar_lb) cudaFree(c->nar_lb);
  if (c->epa_list) cudaFree(c->epa_list);
  if (c->epa_count) cudaFree(c->epa_count);
  if (c->prev_iters) cudaFree(c->prev_iters);
  if (c->roll) cudaFree(c->roll);
  if (c->ws_x) cudaFree(c->ws_x);
  if (c->ws_y) cudaFree(c->ws_y);
  if (c->roll_i) cudaFree(c->roll_i);
  if (c->order) cudaFree(c->order);
  if (c->sched_hist) cudaFree(c->sched_hist);
  if (c->slow_count) cudaFree(c->slow_count);
  {
    for (Scratch* sp : {&c->prio}) {
      Scratch& p = *sp;
      void* ps[] = {p.c_q, p.c_qd, p.c_oMi, p.c_M, p.c_Minv, p.c_g, p.c_nle, p.qp, p.epa_flag, p.cand_mask, p.col_dist, p.col_pair, p.col_wit,
                    p.epa_list, p.epa_count, p.nar_k, p.nar_lb};
      for (void* q : ps) if (q) cudaFree(q);
    }
    if (c->prio_stream) { cudaStreamSynchronize(c->prio_stream); cudaStreamDestroy(c->prio_stream); }
    if (c->prio_side) { cudaStreamSynchronize(c->prio_side); cudaStreamDestroy(c->prio_side); }
    if (c->build_stream) { cudaStreamSynchronize(c->build_stream); cudaStreamDestroy(c->build_stream); }
    if (c->ev_mbuild) cudaEventDestroy(c->ev_mbuild);
    if (c->ev_prio_fk) cudaEventDestroy(c->ev_prio_fk);
    if (c->ev_prio_build) cudaEventDestroy(c->ev_prio_build);
    if (c->ev_build) cudaEventDestroy(c->ev_build);
    if (c->ev_sched) cudaEventDestroy(c->ev_sched);
    if (c->ev_order) cudaEventDestroy(c->ev_order);
    if (c->ev_prio) cudaEventDestroy(c->ev_prio);
    if (c->copy) { cudaStreamSynchronize(c->copy); cudaStreamDestroy(c->copy); }
    if (c->dyn_stream) { cudaStreamSynchronize(c->dyn_stream); cudaStreamDestroy(c->dyn_stream); }
    if (c->ev_store) cudaEventDestroy(c->ev_store);
    if (c->ev_in) cudaEventDestroy(c->ev_in);
    if (c->ev_last) cudaEventDestroy(c->ev_last);
    if (c->ev_out) cudaEventDestroy(c->ev_out);
    if (c->ev_dyn) cudaEventDestroy(c->ev_dyn);
    if (c->ev_solve) cudaEventDestroy(c->ev_solve);
    if (c->ev_late) cudaEventDestroy(c->ev_late);
    if (c->ev_early) cudaEventDestroy(c->ev_early);
  }
  if (c->stage_i) cudaFree(c->stage_i);
  for (int i = 0; i < 4; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  for (int i = 0; i < 32; ++i) if (c->tr_ev[i]) cudaEventDestroy(c->tr_ev[i]);
  if (c->dbg_x) cudaFree(c->dbg_x);
  if (c->dbg_y) cudaFree(c->dbg_y);
  if (c->ev_col) cudaEventDestroy(c->ev_col);
  if (c->ev_epa) cudaEventDestroy(c->ev_epa);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}
int drc_ctx_get_params(const drc_ctx_t* c, drc_params_t* p) {
  if (!c || !p) return fail(DRC_E_INVALID, "null argument");
  const DrcParams& s = c->prm;
  p->alpha = s.alpha; p->slack_weight = s.slack_weight; p->ik_reg = s.ik_reg; p->moma_ik_reg = s.moma_ik_reg;
  p->mani_thresh = s.mani_thresh; p->dist_thresh = s.dist_thresh;
  for (int i = 0; i < 6; ++i) { p->Kp_task[i] = s.Kp_task[i]; p->Kv_task[i] = s.Kv_task[i]; }
  for (int i = 0; i < 16; ++i) { p->Kp_joint[i] = s.Kp_joint[i]; p->Kv_joint[i] = s.Kv_joint[i]; }
  p->rho = s.rho; p->sigma = s.sigma; p->osqp_alpha = s.osqp_alpha; p->eps_abs = s.eps_abs; p->eps_rel = s.eps_rel;
  p->eps_prim_inf = s.eps_prim_inf; p->eps_dual_inf = s.eps_dual_inf; p->max_iter = s.max_iter;
  p->check_termination = s.check_termination; p->scaling = s.scaling; p->adaptive_rho = s.adaptive_rho;
  p->adaptive_rho_interval = s.adaptive_rho_interval; p->adaptive_rho_tolerance = s.adaptive_rho_tolerance;
  p->gjk_tol = s.gjk_tol; p->epa_tol = s.epa_tol; p->gjk_max_iter = s.gjk_max_iter; p->epa_max_iter = s.epa_max_iter;
  p->pinv_threshold = s.pinv_threshold; p->schedule_hint = s.schedule_hint; p->rollout_fused = s.rollout_fused; p->rollout_warm_start = s.rollout_warm_start;
  return DRC_OK;
}
int drc_ctx_set_params(drc_ctx_t* c, const drc_params_t* p) {
  if (!c || !p) return fail(DRC_E_INVALID, "null argument");
  if (p->check_termination <= 0 || p->max_iter <= 0 || p->scaling < 0 || p->rho <= 0 || p->sigma <= 0)
    return fail(DRC_E_INVALID, "invalid solver settings");
  if (p->epa_max_iter > kEpaMaxVert - 4) return fail(DRC_E_INVALID, "epa_max_iter exceeds the polytope capacity (100)");
  DrcParams& s = c->prm;
  s.alpha = p->alpha; s.slack_weight = p->slack_weight; s.ik_reg = p->ik_reg; s.moma_ik_reg = p->moma_ik_reg;
  s.mani_thresh = p->mani_thresh; s.dist_thresh = p->dist_thresh;
  for (int i = 0; i < 6; ++i) { s.Kp_task[i] = p->Kp_task[i]; s.Kv_task[i] = p->Kv_task[i]; }
  for (int i = 0; i < 16; ++i) { s.Kp_joint[i] = p->Kp_joint[i]; s.Kv_joint[i] = p->Kv_joint[i]; }
  s.rho = p->rho; s.sigma = p->sigma; s.osqp_alpha = p->osqp_alpha; s.eps_abs = p->eps_abs; s.eps_rel = p->eps_rel;
  s.eps_prim_inf = p->eps_prim_inf; s.eps_dual_inf = p->eps_dual_inf; s.max_iter = p->max_iter;
  s.check_termination = p->check_termination; s.scaling = p->scaling; s.adaptive_rho = p->adaptive_rho;
  s.adaptive_rho_interval = p->adaptive_rho_interval; s.adaptive_rho_tolerance = p->adaptive_rho_tolerance;
  s.gjk_tol = p->gjk_tol; s.epa_tol = p->epa_tol; s.gjk_max_iter = p->gjk_max_iter; s.epa_max_iter = p->epa_max_iter;
  s.pinv_threshold = p->pinv_threshold; s.schedule_hint = p->schedule_hint; s.rollout_fused = p->rollout_fused; s.rollout_warm_start = p->rollout_warm_start;
  return DRC_OK;
}
int drc_ctx_max_batch(const drc_ctx_t* c) { return c ? c->cap : DRC_E_INVALID; }
int drc_ctx_synchronize(drc_ctx_t* c) {
  if (!c) return fail(DRC_E_INVALID, "null context");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaStreamSynchronize(c->side));
  CU(cudaStreamSynchronize(c->prio_stream));
  CU(cudaStreamSynchronize(c->prio_side));
  CU(cudaStreamSynchronize(c->build_stream));
  CU(cudaStreamSynchronize(c->dyn_stream));
  return DRC_OK;
}
void* drc_ctx_stream(drc_ctx_t* c) { return c ? (void*)c->stream : nullptr; }
int drc_ctx_enable_timing(drc_ctx_t* c, int on) { if (!c) return fail(DRC_E_INVALID, "null context"); c->timing = on != 0; return DRC_OK; }
int drc_ctx_last_timing(drc_ctx_t* c, float* ms) {
  if (!c || !ms) return fail(DRC_E_INVALID, "null argument");
  if (!c->timing) return fail(DRC_E_INVALID, "timing is not enabled on this context");
  CU(cudaEventSynchronize(c->ev[3]));
  CU(cudaEventElapsedTime(&ms[0], c->ev[0], c->ev[1]));
  CU(cudaEventElapsedTime(&ms[1], c->ev[1], c->ev[2]));
  CU(cudaEventElapsedTime(&ms[2], c->ev[2], c->ev[3]));
  CU(cudaEventElapsedTime(&ms[3], c->ev[0], c->ev[3]));
  return DRC_OK;
}
long long drc_ctx_launch_count(const drc_ctx_t* c) { return c ? c->launches : 0; }
int drc_ctx_last_trace(drc_ctx_t* c, int max_marks, float* ms, char* names, int names_len) {
  if (!c || !ms || !names || names_len <= 0) return fail(DRC_E_INVALID, "null argument");
  if (!c->timing) return fail(DRC_E_INVALID, "timing is not enabled on this context");
  std::string all;
  int n = 0;
  for (int i = 0; i < c->tr_n && n < max_marks; ++i) {
    CU(cudaEventSynchronize(c->tr_ev[i]));
    CU(cudaEventElapsedTime(&ms[n], c->tr_ev[0], c->tr_ev[i]));
    all += c->tr_name[i]; all += ';';
    ++n;
  }
  std::snprintf(names, (size_t)names_len, "%s", all.c_str());
  return n;
}
int drc_ctx_enable_qp_debug(drc_ctx_t* c, int on) {
  if (!c) return fail(DRC_E_INVALID, "null context");
  CU(cudaSetDevice(c->device));
  if (on && !c->dbg_x) {
    const int n = c->model->hm.dev.nv;
    CU(cudaMalloc((void**)&c->dbg_x, (size_t)c->cap * (6 * n + 2) * sizeof(double)));
    CU(cudaMalloc((void**)&c->dbg_y, (size_t)c->cap * (11 * n + 4) * sizeof(double)));
  } else if (!on && c->dbg_x) {
    CU(cudaDeviceSynchronize());
    cudaFree(c->dbg_x); cudaFree(c->dbg_y);
    c->dbg_x = c->dbg_y = nullptr;
  }
  return DRC_OK;
}
int drc_host_get_qp_debug(drc_ctx_t* c, int B, int x_per_robot, int y_per_robot, double* x, double* y) {
  int rc = check_batch(c, B); if (rc) return rc;
  if (!c->dbg_x) return fail(DRC_E_INVALID, "QP debug outputs are not enabled on this context");
  const int n = c->model->hm.dev.nv;
  if (x_per_robot < 0 || x_per_robot > 6 * n + 2 || y_per_robot < 0 || y_per_robot > 11 * n + 4) return fail(DRC_E_INVALID, "debug vector size out of range");
  CU(cudaSetDevice(c->device));
  CU(cudaDeviceSynchronize());
  if (x) CU(cudaMemcpy(x, c->dbg_x, (size_t)B * x_per_robot * sizeof(double), cudaMemcpyDeviceToHost));
  if (y) CU(cudaMemcpy(y, c->dbg_y, (size_t)B * y_per_robot * sizeof(double), cudaMemcpyDeviceToHost));
  return DRC_OK;
}

// ------------------------------------------------------------------------------------------------ device entry points
int drc_batch_update_state(drc_ctx_t* c, int B, const double* q, const double* qdot, int layout, void* stream) {
  int rc = check_batch(c, B); if (rc) return rc;
  if (!q || !qdot) return fail(DRC_E_INVALID, "null state pointer");
  CU(cudaSetDevice(c->device));
  const int n = c->model->hm.dev.nv;
  JobIO io; std::memset(&io, 0, sizeof io);
  io.B = B; io.q = q; io.qd = qdot; io.sq = lay(layout, n, B); io.sqd = io.sq;
  bind_cache(c, io);
  DrcFrame fr; std::memset(&fr, 0, sizeof fr); fr.parent = -1;
  DRC_DISPATCH_NV(n, c->model->hm.chain, return (launch_job<NV, CHAIN, F_DYN | F_STORE>(c, fr, io, pick(c, stream))));
}
int drc_batch_get_frame(drc_ctx_t* c, int B, int frame, double* pose12, double* J, double* Jdot, double* vel, int layout, void* stream) {
  int rc = check_batch(c, B); if (rc) return rc;
  rc = check_frame(c, frame); if (rc) return rc;
  CU(cudaSetDevice(c->device));
  if (c->model->hm.dev.drive_type != kNoBase) return moma_get_frame_full(c, B, frame, pose12, J, Jdot, vel, layout, pick(c, stream));
  const int n = c->model->hm.dev.nv;
  JobIO io; std::memset(&io, 0, sizeof io);
  io.B = B; bind_cache(c, io);
  io.pose = pose12; io.spose = lay(layout, 12, B); io.J = J; io.sJ = lay(layout, 6 * n, B); io.Jdot = Jdot; io.sJd = io.sJ;
  io.vel = vel; io.svel = lay(layout, 6, B);
  const DrcFrame fr = frame_of(c->model, frame);
  DRC_DISPATCH_NV(n, c->model->hm.chain, return (launch_job<NV, CHAIN, F_FROM_CACHE | F_FRAME_OUT>(c, fr, io, pick(c, stream))));
}
int drc_batch_get_dynamics(drc_ctx_t* c, int B, double* M, double* Minv, double* g, double* coriolis, double* nle, int layout, void* stream) {
  int rc = check_batch(c, B); if (rc) return rc;
  CU(cudaSetDevice(c->device));
  const int n = c->model->hm.dev.nv;
  cudaStream_t s = pick(c, stream);
  auto cp = [&](const double* src, const double* sub, int K, double* dst) -> int {
    if (!dst) return DRC_OK;
    const long long tot = (long long)K * B;
    k_copy_cache<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(src, sub, c->cap, K, B, dst, lay(layout, K, B));
    c->launches++;
    CU(cudaGetLastError());
    return DRC_OK;
  };
  if ((rc = cp(c->c_M, nullptr, n * n, M))) return rc;
  if ((rc = cp(c->c_Minv, nullptr, n * n, Minv))) return rc;
  if ((rc = cp(c->c_g, nullptr, n, g))) return rc;
  if ((rc = cp(c->c_nle, c->c_g, n, coriolis))) return rc;
  return cp(c->c_nle, nullptr, n, nle);
}
int drc_batch_get_manipulability(drc_ctx_t* c, int B, int frame, int with_graddot, double* mani, double* grad, double* grad_dot, int layout, void* stream) {
  int rc = check_batch(c, B); if (rc) return rc;
  rc = check_frame(c, frame); if (rc) return rc;
  if (!mani) return fail(DRC_E_INVALID, "null output");
  CU(cudaSetDevice(c->device));
  const int n = c->model->hm.dev.nv;
  JobIO io; std::memset(&io, 0, sizeof io);
  io.B = B; bind_cache(c, io);
  io.mani = mani; io.mani_grad = grad; io.smg = lay(layout, n, B); io.mani_graddot = grad_dot; io.smgd = io.smg;
  const DrcFrame fr = frame_of(c->model, frame);
  if (with_graddot) {
    DRC_DISPATCH_NV(n, c->model->hm.chain, return (launch_job<NV, CHAIN, F_FROM_CACHE | F_MANIP_OUT | F_GRADDOT>(c, fr, io, pick(c, stream))));
  }
  DRC_DISPATCH_NV(n, c->model->hm.chain, return (launch_job<NV, CHAIN, F_FROM_CACHE | F_MANIP_OUT>(c, fr, io, pick(c, stream))));
}
int drc_batch_get_min_distance(drc_ctx_t* c, int B, int with_graddot, double* dist, double* grad, double* grad_dot, int* pair, int layout, void* stream) {
  int rc = check_batch(c, B); if (rc) return rc;
  CU(cudaSetDevice(c->device));
  if (c->model->hm.dev.drive_type != kNoBase) return moma_get_min_distance(c, B, with_graddot, dist, grad, grad_dot, pair, layout, pick(c, stream));
  const int n = c->model->hm.dev.nv;
  CollisionIO io; std::memset(&io, 0, sizeof io);
  io.B = B; io.mode = 0; io.dist = dist; io.grad = grad; io.sgrad = lay(layout, n, B);
  io.grad_dot = with_graddot ? grad_dot : nullptr; io.sgd = io.sgrad; io.pair_out = pair;
  DRC_DISPATCH_NV(n, c->model->hm.chain, return (launch_collision<NV, CHAIN>(c, io, pick(c, stream))));
}

#define QP_ENTRY(ID_, STEP_, Q_, QD_, XT_, XD_, OUT_, OUT2_)                                                        \
  int rc = check_batch(c, B); if (rc) return rc;                                                                    \
  rc = check_frame(c, frame); if (rc) return rc;                                                                    \
  if (!(XD_) || !(OUT_)) return fail(DRC_E_INVALID, "null argument");                                               \
  CU(cudaSetDevice(c->device));                                                                                     \
  DRC_DISPATCH_NV(c->model->hm.dev.nv, c->model->hm.chain,                                                          \
                  return (run_qp<NV, CHAIN>(c, B, ID_, STEP_, Q_, QD_, XT_, XD_, frame, OUT_, OUT2_, status, iters, layout, pick(c, stream))));

int drc_batch_qpik(drc_ctx_t* c, int B, const double* xdot_des, int frame, double* qdot_out, int* status, int* iters, int layout, void* stream) {
  QP_ENTRY(false, false, nullptr, nullptr, nullptr, xdot_des, qdot_out, nullptr)
}
int drc_batch_qpik_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, int frame, double* qdot_out, int* status, int* iters, int layout, void* stream) {
  if (!x_target) return fail(DRC_E_INVALID, "null target pose");
  QP_ENTRY(false, true, nullptr, nullptr, x_target, xdot_target, qdot_out, nullptr)
}
int drc_batch_qpid(drc_ctx_t* c, int B, const double* xddot_des, int frame, double* tau_out, double* qddot_out, int* status, int* iters, int layout, void* stream) {
  QP_ENTRY(true, false, nullptr, nullptr, nullptr, xddot_des, tau_out, qddot_out)
}
int drc_batch_qpid_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, int frame, double* tau_out, double* qddot_out, int* status, int* iters, int layout, void* stream) {
  if (!x_target) return fail(DRC_E_INVALID, "null target pose");
  QP_ENTRY(true, true, nullptr, nullptr, x_target, xdot_target, tau_out, qddot_out)
}
int drc_batch_cycle_qpik_step(drc_ctx_t* c, int B, const double* q, const double* qdot, const double* x_target, const double* xdot_target, int frame, double* qdot_out, int* status, int* iters, int layout, void* stream) {
  if (!q || !qdot || !x_target) return fail(DRC_E_INVALID, "null argument");
  QP_ENTRY(false, true, q, qdot, x_target, xdot_target, qdot_out, nullptr)
}
int drc_batch_cycle_qpid_step(drc_ctx_t* c, int B, const double* q, const double* qdot, const double* x_target, const double* xdot_target, int frame, double* tau_out, int* status, int* iters, int layout, void* stream) {
  if (!q || !qdot || !x_target) return fail(DRC_E_INVALID, "null argument");
  QP_ENTRY(true, true, q, qdot, x_target, xdot_target, tau_out, nullptr)
}

// BASELINE config 2 as ONE job: updateState (state + dynamics -> cache) + CLIKStep (qdot*) + OSFStep (tau*) from the values the job
// already holds in registers, instead of three launches that each redo the forward kinematics.  No null-space vectors (use the
// separate entry points for those).
int drc_batch_cycle_clik_osf_step(drc_ctx_t* c, int B, const double* q, const double* qdot, const double* x_target, const double* xdot_target,
                                  int frame, double* qdot_out, double* tau_out, int layout, void* stream) {
  int rc = check_batch(c, B); if (rc) return rc;
  rc = check_frame(c, frame); if (rc) return rc;
  if (!q || !qdot || !x_target || !xdot_target || !qdot_out || !tau_out) return fail(DRC_E_INVALID, "null argument");
  CU(cudaSetDevice(c->device));
  if (c->model->hm.dev.drive_type != kNoBase) return fail(DRC_E_UNSUPPORTED, "CLIK / OSF are manipulator controllers");
  const int n = c->model->hm.dev.nv;
  JobIO io; std::memset(&io, 0, sizeof io);
  io.B = B; io.q = q; io.qd = qdot; io.sq = lay(layout, n, B); io.sqd = io.sq;
  io.x_target = x_target; io.sxt = lay(layout, 12, B); io.xdot_target = xdot_target; io.sxd = lay(layout, 6, B);
  io.out = qdot_out; io.out2 = tau_out; io.sout = lay(layout, n, B); io.saux = io.sout; io.saux2 = io.sout;
  bind_cache(c, io);
  const DrcFrame fr = frame_of(c->model, frame);
  DRC_DISPATCH_NV(n, c->model->hm.chain, return (launch_job<NV, CHAIN, F_DYN | F_STORE | F_CLIK | F_OSF | F_STEP>(c, fr, io, pick(c, stream))));
}

// T control ticks of QPIKCubic / QPIKStep + the integrate step, without a host round trip between ticks
int drc_batch_rollout_qpik(drc_ctx_t* c, int B, int T, double dt, double* q, double* qdot, const double* x_target,
                           const double* xdot_target, const double* x_init, const double* xdot_init, double t_start, double t0,
                           double duration, int frame, int* fail_ticks, int* iters_total, int layout, void* stream) {
  int rc = check_batch(c, B); if (rc) return rc;
  rc = check_frame(c, frame); if (rc) return rc;
  if (!q || !qdot || !x_target || !xdot_target) return fail(DRC_E_INVALID, "null argument");
  if (T <= 0 || !(dt > 0)) return fail(DRC_E_INVALID, "rollout needs T > 0 ticks and dt > 0");
  if (duration > 0 && (!x_init || !xdot_init)) return fail(DRC_E_INVALID, "a cubic profile needs x_init and xdot_init");
  CU(cudaSetDevice(c->device));
  cudaStream_t s = pick(c, stream);
  DRC_DISPATCH_NV(c->model->hm.dev.nv, c->model->hm.chain,
                  return (run_rollout<NV, CHAIN>(c, B, T, dt, q, qdot, x_target, xdot_target, x_init, xdot_init, t_start, t0, duration, frame,
                                                 fail_ticks, iters_total, layout, s)));
}

static int taskspace(drc_ctx_t* c, int B, int kind, const double* x_target, const double* xdot, const double* aux, const double* aux2,
                     int frame, double* out, int layout, void* stream) {
  int rc = check_batch(c, B); if (rc) return rc;
  if (kind != 3) { rc = check_frame(c, frame); if (rc) return rc; }
  if (!out) return fail(DRC_E_INVALID, "null output");
  CU(cudaSetDevice(c->device));
  const int n = c->model->hm.dev.nv;
  JobIO io; std::memset(&io, 0, sizeof io);
  io.B = B; bind_cache(c, io);
  io.x_target = x_target; io.sxt = lay(layout, 12, B); io.xdot_target = xdot; io.sxd = lay(layout, 6, B);
  io.aux = aux; io.saux = lay(layout, n, B); io.aux2 = aux2; io.saux2 = io.saux; io.out = out; io.sout = io.saux;
  DrcFrame fr; std::memset(&fr, 0, sizeof fr); fr.parent = -1;
  if (kind != 3) fr = frame_of(c->model, frame);
  cudaStream_t s = pick(c, stream);
  if (kind == 0) { DRC_DISPATCH_NV(n, c->model->hm.chain, return (launch_job<NV, CHAIN, F_FROM_CACHE | F_CLIK | F_STEP>(c, fr, io, s))); }
  if (kind == 1) { DRC_DISPATCH_NV(n, c->model->hm.chain, return (launch_job<NV, CHAIN, F_FROM_CACHE | F_OSF | F_STEP>(c, fr, io, s))); }
  if (kind == 2) { DRC_DISPATCH_NV(n, c->model->hm.chain, return (launch_job<NV, CHAIN, F_FROM_CACHE | F_OSF>(c, fr, io, s))); }
  DRC_DISPATCH_NV(n, c->model->hm.chain, return (launch_job<NV, CHAIN, F_FROM_CACHE | F_TORQUE>(c, fr, io, s)));
}
int drc_batch_clik_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, const double* null_qdot, int frame, double* qdot_out, int layout, void* stream) {
  if (!x_target || !xdot_target) return fail(DRC_E_INVALID, "null argument");
  return taskspace(c, B, 0, x_target, xdot_target, null_qdot, nullptr, frame, qdot_out, layout, stream);
}
int drc_batch_osf_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, const double* null_torque, int frame, double* tau_out, int layout, void* stream) {
  if (!x_target || !xdot_target) return fail(DRC_E_INVALID, "null argument");
  return taskspace(c, B, 1, x_target, xdot_target, null_torque, nullptr, frame, tau_out, layout, stream);
}
int drc_batch_osf(drc_ctx_t* c, int B, const double* xddot_target, const double* null_torque, int frame, double* tau_out, int layout, void* stream) {
  if (!xddot_target) return fail(DRC_E_INVALID, "null argument");
  return taskspace(c, B, 2, nullptr, xddot_target, null_torque, nullptr, frame, tau_out, layout, stream);
}
int drc_batch_joint_torque_step(drc_ctx_t* c, int B, const double* q_target, const double* qdot_target, double* tau_out, int layout, void* stream) {
  if (!q_target || !qdot_target) return fail(DRC_E_INVALID, "null argument");
  return taskspace(c, B, 3, nullptr, nullptr, q_target, qdot_target, 0, tau_out, layout, stream);
}
int drc_batch_task_space_cubic(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, const double* x_init, const double* xdot_init,
                               double t, double t0, double duration, double* x_des, double* xdot_des, int layout, void* stream) {
  int rc = check_batch(c, B); if (rc) return rc;
  if (!x_target || !xdot_target || !x_init || !xdot_init || !x_des || !xdot_des) return fail(DRC_E_INVALID, "null argument");
  CU(cudaSetDevice(c->device));
  CubicIO io;
  io.B = B; io.x_target = x_target; io.xdot_target = xdot_target; io.x_init = x_init; io.xdot_init = xdot_init;
  io.s12 = lay(layout, 12, B); io.s6 = lay(layout, 6, B); io.t = t; io.t0 = t0; io.dur = duration; io.x_des = x_des; io.xdot_des = xdot_des;
  k_task_cubic<<<(B + 127) / 128, 128, 0, pick(c, stream)>>>(io);
  c->launches++;
  CU(cudaGetLastError());
  return DRC_OK;
}

// ------------------------------------------------------------------------------------------------ host entry points
int drc_host_update_state(drc_ctx_t* c, int B, const double* q, const double* qdot) {
  HOST_PRELUDE
  const double *dq = st.in(q, Bz * n), *dqd = st.in(qdot, Bz * n);
  return st.finish(st.err ? st.err : drc_batch_update_state(c, B, dq, dqd, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_get_frame(drc_ctx_t* c, int B, int frame, double* pose12, double* J, double* Jdot, double* vel) {
  HOST_PRELUDE
  double *dp = st.out(pose12, Bz * 12), *dJ = st.out(J, Bz * 6 * n), *dJd = st.out(Jdot, Bz * 6 * n), *dv = st.out(vel, Bz * 6);
  return st.finish(st.err ? st.err : drc_batch_get_frame(c, B, frame, dp, dJ, dJd, dv, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_get_dynamics(drc_ctx_t* c, int B, double* M, double* Minv, double* g, double* coriolis, double* nle) {
  HOST_PRELUDE
  double *dM = st.out(M, Bz * n * n), *dMi = st.out(Minv, Bz * n * n), *dg = st.out(g, Bz * n), *dc = st.out(coriolis, Bz * n), *dn = st.out(nle, Bz * n);
  return st.finish(st.err ? st.err : drc_batch_get_dynamics(c, B, dM, dMi, dg, dc, dn, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_get_manipulability(drc_ctx_t* c, int B, int frame, int with_graddot, double* mani, double* grad, double* grad_dot) {
  HOST_PRELUDE
  double *dm = st.out(mani, Bz), *dg = st.out(grad, Bz * n), *dgd = st.out(with_graddot ? grad_dot : nullptr, Bz * n);
  return st.finish(st.err ? st.err : drc_batch_get_manipulability(c, B, frame, with_graddot, dm, dg, dgd, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_get_min_distance(drc_ctx_t* c, int B, int with_graddot, double* dist, double* grad, double* grad_dot, int* pair) {
  HOST_PRELUDE
  double *dd = st.out(dist, Bz), *dg = st.out(grad, Bz * n), *dgd = st.out(with_graddot ? grad_dot : nullptr, Bz * n);
  int* dp = st.out_i(pair, Bz);
  return st.finish(st.err ? st.err : drc_batch_get_min_distance(c, B, with_graddot, dd, dg, dgd, dp, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_qpik(drc_ctx_t* c, int B, const double* xdot_des, int frame, double* qdot_out, int* status, int* iters) {
  HOST_PRELUDE
  const double* dx = st.in(xdot_des, Bz * 6);
  double* dout = st.out(qdot_out, Bz * n);
  int *ds = st.out_i(status, Bz), *di = st.out_i(iters, Bz);
  return st.finish(st.err ? st.err : drc_batch_qpik(c, B, dx, frame, dout, ds, di, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_qpik_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, int frame, double* qdot_out, int* status, int* iters) {
  HOST_PRELUDE
  const double *dxt = st.in(x_target, Bz * 12), *dx = st.in(xdot_target, Bz * 6);
  double* dout = st.out(qdot_out, Bz * n);
  int *ds = st.out_i(status, Bz), *di = st.out_i(iters, Bz);
  return st.finish(st.err ? st.err : drc_batch_qpik_step(c, B, dxt, dx, frame, dout, ds, di, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_qpid(drc_ctx_t* c, int B, const double* xddot_des, int frame, double* tau_out, double* qddot_out, int* status, int* iters) {
  HOST_PRELUDE
  const double* dx = st.in(xddot_des, Bz * 6);
  double *dout = st.out(tau_out, Bz * n), *dout2 = st.out(qddot_out, Bz * n);
  int *ds = st.out_i(status, Bz), *di = st.out_i(iters, Bz);
  return st.finish(st.err ? st.err : drc_batch_qpid(c, B, dx, frame, dout, dout2, ds, di, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_qpid_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, int frame, double* tau_out, double* qddot_out, int* status, int* iters) {
  HOST_PRELUDE
  const double *dxt = st.in(x_target, Bz * 12), *dx = st.in(xdot_target, Bz * 6);
  double *dout = st.out(tau_out, Bz * n), *dout2 = st.out(qddot_out, Bz * n);
  int *ds = st.out_i(status, Bz), *di = st.out_i(iters, Bz);
  return st.finish(st.err ? st.err : drc_batch_qpid_step(c, B, dxt, dx, frame, dout, dout2, ds, di, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_clik_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, const double* null_qdot, int frame, double* qdot_out) {
  HOST_PRELUDE
  const double *dxt = st.in(x_target, Bz * 12), *dx = st.in(xdot_target, Bz * 6), *dn = st.in(null_qdot, Bz * n);
  double* dout = st.out(qdot_out, Bz * n);
  return st.finish(st.err ? st.err : drc_batch_clik_step(c, B, dxt, dx, dn, frame, dout, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_osf(drc_ctx_t* c, int B, const double* xddot_target, const double* null_torque, int frame, double* tau_out) {
  HOST_PRELUDE
  const double *dx = st.in(xddot_target, Bz * 6), *dn = st.in(null_torque, Bz * n);
  double* dout = st.out(tau_out, Bz * n);
  return st.finish(st.err ? st.err : drc_batch_osf(c, B, dx, dn, frame, dout, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_osf_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, const double* null_torque, int frame, double* tau_out) {
  HOST_PRELUDE
  const double *dxt = st.in(x_target, Bz * 12), *dx = st.in(xdot_target, Bz * 6), *dn = st.in(null_torque, Bz * n);
  double* dout = st.out(tau_out, Bz * n);
  return st.finish(st.err ? st.err : drc_batch_osf_step(c, B, dxt, dx, dn, frame, dout, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_joint_torque_step(drc_ctx_t* c, int B, const double* q_target, const double* qdot_target, double* tau_out) {
  HOST_PRELUDE
  const double *dq = st.in(q_target, Bz * n), *dqd = st.in(qdot_target, Bz * n);
  double* dout = st.out(tau_out, Bz * n);
  return st.finish(st.err ? st.err : drc_batch_joint_torque_step(c, B, dq, dqd, dout, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_task_space_cubic(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, const double* x_init, const double* xdot_init,
                              double t, double t0, double duration, double* x_des, double* xdot_des) {
  HOST_PRELUDE
  const double *a = st.in(x_target, Bz * 12), *b = st.in(xdot_target, Bz * 6), *ci = st.in(x_init, Bz * 12), *d = st.in(xdot_init, Bz * 6);
  double *o1 = st.out(x_des, Bz * 12), *o2 = st.out(xdot_des, Bz * 6);
  return st.finish(st.err ? st.err : drc_batch_task_space_cubic(c, B, a, b, ci, d, t, t0, duration, o1, o2, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_cycle_qpik_step(drc_ctx_t* c, int B, const double* q, const double* qdot, const double* x_target, const double* xdot_target, int frame, double* qdot_out, int* status, int* iters) {
  HOST_PRELUDE
  // q, qdot feed stage 1 (FK, self-collision); the targets are only read by the QP-build kernel: upload them behind stage 1
  const double *dq = st.in(q, Bz * n), *dqd = st.in(qdot, Bz * n), *dxt = st.in_late(x_target, Bz * 12), *dx = st.in_late(xdot_target, Bz * 6);
  st.late_done();
  double* dout = st.out_direct(qdot_out, Bz * n);
  int *ds = st.out_i_direct(status, Bz), *di = st.out_i_direct(iters, Bz);
  return st.finish(st.err ? st.err : drc_batch_cycle_qpik_step(c, B, dq, dqd, dxt, dx, frame, dout, ds, di, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_cycle_qpid_step(drc_ctx_t* c, int B, const double* q, const double* qdot, const double* x_target, const double* xdot_target, int frame, double* tau_out, int* status, int* iters) {
  HOST_PRELUDE
  const double *dq = st.in(q, Bz * n), *dqd = st.in(qdot, Bz * n), *dxt = st.in_late(x_target, Bz * 12), *dx = st.in_late(xdot_target, Bz * 6);
  st.late_done();
  double* dout = st.out_direct(tau_out, Bz * n);
  int *ds = st.out_i_direct(status, Bz), *di = st.out_i_direct(iters, Bz);
  return st.finish(st.err ? st.err : drc_batch_cycle_qpid_step(c, B, dq, dqd, dxt, dx, frame, dout, ds, di, DRC_LAYOUT_AOS, nullptr));
}

int drc_host_cycle_clik_osf_step(drc_ctx_t* c, int B, const double* q, const double* qdot, const double* x_target, const double* xdot_target, int frame, double* qdot_out, double* tau_out) {
  HOST_PRELUDE
  const double *dq = st.in(q, Bz * n), *dqd = st.in(qdot, Bz * n), *dxt = st.in(x_target, Bz * 12), *dx = st.in(xdot_target, Bz * 6);
  double *dout = st.out_direct(qdot_out, Bz * n), *dout2 = st.out_direct(tau_out, Bz * n);
  return st.finish(st.err ? st.err : drc_batch_cycle_clik_osf_step(c, B, dq, dqd, dxt, dx, frame, dout, dout2, DRC_LAYOUT_AOS, nullptr));
}

int drc_host_rollout_qpik(drc_ctx_t* c, int B, int T, double dt, double* q, double* qdot, const double* x_target,
                          const double* xdot_target, const double* x_init, const double* xdot_init, double t_start, double t0,
                          double duration, int frame, int* fail_ticks, int* iters_total) {
  HOST_PRELUDE
  if (!q || !qdot) return fail(DRC_E_INVALID, "null state pointer");
  double *dq = st.in(q, Bz * n), *dqd = st.in(qdot, Bz * n);
  const double *dxt = st.in(x_target, Bz * 12), *dx = st.in(xdot_target, Bz * 6), *dxi = st.in(x_init, Bz * 12), *dxdi = st.in(xdot_init, Bz * 6);
  if (dq) st.outs.push_back({q, dq, Bz * n * sizeof(double)});      // the state arrays are updated in place
  if (dqd) st.outs.push_back({qdot, dqd, Bz * n * sizeof(double)});
  int *df = st.out_i(fail_ticks, Bz), *di = st.out_i(iters_total, Bz);
  return st.finish(st.err ? st.err : drc_batch_rollout_qpik(c, B, T, dt, dq, dqd, dxt, dx, dxi, dxdi, t_start, t0, duration, frame, df, di,
                                                            DRC_LAYOUT_AOS, nullptr));
}

int drc_bench_fp64_peak(int device, double* tflops) {
  if (!tflops) return fail(DRC_E_INVALID, "null output");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return fail(DRC_E_CUDA, "no such CUDA device"); }
  CU(cudaSetDevice(device));
  cudaDeviceProp p;
  CU(cudaGetDeviceProperties(&p, device));
  const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
  double* out;
  CU(cudaMalloc((void**)&out, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t a, b;
  CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
  k_fp64_peak<<<blocks, threads>>>(out, 1024);  // warm-up
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    CU(cudaEventRecord(a));
    k_fp64_peak<<<blocks, threads>>>(out, iters);
    CU(cudaEventRecord(b));
    CU(cudaEventSynchronize(b));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, a, b));
    const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
    const double tf = fl / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
  *tflops = best;
  return DRC_OK;
}

}  // extern "C"
