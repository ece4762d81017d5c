// TEST INFRASTRUCTURE ONLY -- host emulation of the CUDA kernel bodies.
//
// The device code of dyros_robot_controller_b200/csrc (drc_kin.h, drc_geom.h, drc_qp.h, drc_cycle.h)
// is written as host+device inlines; this file instantiates the SAME bodies with g++ and runs
// thread-per-robot stages in a plain loop and warp-per-robot stages through WarpEmu (32 lanes
// emulated sequentially, __syncwarp() == end of a lane loop).  It lets `pytest -m "not gpu"` check
// kernel logic against the oracle in the CPU-only build container.
// It is NOT part of the product: libdrc_b200.so never links it and the package never loads it.
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../dyros_robot_controller_b200/csrc/drc_cycle.h"
#include "../../dyros_robot_controller_b200/csrc/model.h"

using namespace drc;

struct EmuHandle {
  HostModel hm;
  DrcParams prm;
};

static DrcFrame make_frame(const HostModel& hm, int fid) {
  DrcFrame f;
  f.parent = hm.frames[fid].parent;
  std::memcpy(f.R, hm.frames[fid].R, sizeof f.R);
  std::memcpy(f.p, hm.frames[fid].p, sizeof f.p);
  return f;
}

struct Cache {
  std::vector<double> q, qd, oMi, M, Minv, g, nle;
  long long Bc;
  explicit Cache(int nv, int B) : q(nv * B), qd(nv * B), oMi(12 * nv * B), M(nv * nv * B), Minv(nv * nv * B), g(nv * B), nle(nv * B), Bc(B) {}
  void bind(JobIO& io) {
    io.c_q = q.data(); io.c_qd = qd.data(); io.c_oMi = oMi.data(); io.c_M = M.data(); io.c_Minv = Minv.data();
    io.c_g = g.data(); io.c_nle = nle.data(); io.Bc = Bc;
  }
};

template <int NV, bool CHAIN, unsigned FLAGS>
static void run_job(const EmuHandle* h, const DrcFrame& fr, const JobIO& io) {
  for (int b = 0; b < io.B; ++b) robot_job<NV, CHAIN, FLAGS>(h->hm.dev, h->prm, fr, io, b);
}

template <int NV, bool CHAIN>
static void run_collision(const EmuHandle* h, CollisionIO& io, std::vector<int>& flag, std::vector<unsigned long long>& mask,
                          std::vector<double>& dist, std::vector<int>& pair, std::vector<double>& wit) {
  flag.assign(io.B, 0); mask.assign(io.B, 0ull);
  if (!io.dist) { dist.assign(io.B, 0.0); io.dist = dist.data(); }
  if (!io.pair_out) { pair.assign(io.B, 0); io.pair_out = pair.data(); }
  if (!io.witness) { wit.assign(6 * (size_t)io.B, 0.0); io.witness = wit.data(); }
  io.epa_flag = flag.data(); io.cand_mask = mask.data();
  for (int b = 0; b < io.B; ++b) collision_job<NV, CHAIN>(h->hm.dev, h->hm.dev.geom, h->prm, io, b);
  for (int b = 0; b < io.B; ++b) collision_epa_job<NV, CHAIN>(h->hm.dev, h->prm, io, b);
}

template <class Cfg, bool ID>
static void run_solve(const EmuHandle* h, const SolveIO& io, unsigned unit_mask) {
  const QpOptions o = qp_options(h->prm, unit_mask);
  std::vector<GroupShared<Cfg>> sh(Cfg::NG);
  std::unique_ptr<WarpEmu<Cfg>> w(new WarpEmu<Cfg>);
  for (int b0 = 0; b0 < io.B; b0 += Cfg::NG) {
    int robots[8];
    for (int g = 0; g < Cfg::NG; ++g) robots[g] = (b0 + g < io.B) ? b0 + g : -1;
    w->sh = sh.data();
    for (int t = 0; t < 32; ++t) lane_assign<Cfg>(w->Ls[t], t);
    solve_and_emit<Cfg, ID>(*w, robots, io, o);
  }
}

extern "C" {

// urdf_dir / packages_path: where <mesh> collision elements are looked up (may be null)
EmuHandle* emu_create2(const char* urdf_text, const char* srdf_text, const char* urdf_dir, const char* packages_path, char* err, int errlen) {
  try {
    std::unique_ptr<EmuHandle> h(new EmuHandle);
    MeshSource ms;
    ms.urdf_dir = urdf_dir ? urdf_dir : ""; ms.packages_path = packages_path ? packages_path : "";
    h->hm = compile_model(urdf_text, srdf_text ? srdf_text : "", ms);
    h->hm.bind_hull();
    for (int i = 0; i < kMaxV; ++i) { h->prm.Kp_joint[i] = 400; h->prm.Kv_joint[i] = 40; }
    return h.release();
  } catch (const std::exception& e) {
    if (err && errlen > 0) { std::strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; }
    return nullptr;
  }
}
EmuHandle* emu_create(const char* urdf_text, const char* srdf_text, char* err, int errlen) {
  return emu_create2(urdf_text, srdf_text, nullptr, nullptr, err, errlen);
}
void emu_destroy(EmuHandle* h) { delete h; }
int emu_nv(EmuHandle* h) { return h->hm.dev.nv; }
int emu_is_chain(EmuHandle* h) { return h->hm.chain ? 1 : 0; }
int emu_frame_id(EmuHandle* h, const char* name) { return h->hm.frame_id(name); }
int emu_model_sizes(EmuHandle* h, int* out) {
  const DrcModelDev& d = h->hm.dev;
  out[0] = d.nv; out[1] = d.ngeom; out[2] = d.npair; out[3] = d.ngroup; out[4] = (int)h->hm.frames.size(); out[5] = h->hm.skipped_geoms;
  return (int)sizeof(DrcModelDev);
}
// mesh geometry: out[0] = mesh geometries, out[1] = hull vertices in total; per geometry vertex counts in vert_n (ngeom ints);
// hull (optional): the vertices (3 * out[1] doubles, geometry frames about the placement points)
void emu_mesh_info(EmuHandle* h, int* out, int* vert_n, double* hull) {
  out[0] = h->hm.mesh_geoms; out[1] = (int)h->hm.hull.size() / 3;
  for (int g = 0; g < h->hm.dev.ngeom; ++g) vert_n[g] = h->hm.dev.geom.vert_n[g];
  if (hull) std::copy(h->hm.hull.begin(), h->hm.hull.end(), hull);
}
// flat copy of the compiled model for comparison with the oracle's independent loader
void emu_model_arrays(EmuHandle* h, int* parent, int* jtype, double* axis, double* jR, double* jp, double* mass, double* com,
                      double* inertia6, double* q_lo, double* q_hi, double* v_lim, int* geom_type, int* geom_parent,
                      double* geom_prm, double* geom_R, double* geom_p, int* pairs_ref_order, int* frame_parent,
                      double* frame_R, double* frame_p) {
  const DrcModelDev& d = h->hm.dev;
  for (int i = 0; i < d.nv; ++i) {
    parent[i] = d.parent[i]; jtype[i] = d.jtype[i];
    std::memcpy(axis + 3 * i, d.axis[i], 24); std::memcpy(jR + 9 * i, d.jR[i], 72); std::memcpy(jp + 3 * i, d.jp[i], 24);
    mass[i] = d.mass[i]; std::memcpy(com + 3 * i, d.com[i], 24); std::memcpy(inertia6 + 6 * i, d.inertia[i], 48);
    q_lo[i] = d.q_lo[i]; q_hi[i] = d.q_hi[i]; v_lim[i] = d.v_lim[i];
  }
  for (int g = 0; g < d.ngeom; ++g) {
    geom_type[g] = d.geom.type[g]; geom_parent[g] = d.geom.parent[g];
    std::memcpy(geom_prm + 3 * g, d.geom.prm[g], 24); std::memcpy(geom_R + 9 * g, d.geom.R[g], 72); std::memcpy(geom_p + 3 * g, d.geom.p[g], 24);
  }
  for (int k = 0; k < d.npair; ++k) { pairs_ref_order[2 * d.geom.pair_id[k]] = d.geom.pair_a[k]; pairs_ref_order[2 * d.geom.pair_id[k] + 1] = d.geom.pair_b[k]; }
  for (size_t f = 0; f < h->hm.frames.size(); ++f) {
    frame_parent[f] = h->hm.frames[f].parent;
    std::memcpy(frame_R + 9 * f, h->hm.frames[f].R, 72); std::memcpy(frame_p + 3 * f, h->hm.frames[f].p, 24);
  }
}
void emu_set_params(EmuHandle* h, const double* kp_task, const double* kv_task, const double* kp_joint, const double* kv_joint,
                    int adaptive_rho_interval, int max_iter, double gjk_tol) {
  if (kp_task) for (int i = 0; i < 6; ++i) { h->prm.Kp_task[i] = kp_task[i]; h->prm.Kv_task[i] = kv_task[i]; }
  if (kp_joint) for (int i = 0; i < h->hm.dev.nv; ++i) { h->prm.Kp_joint[i] = kp_joint[i]; h->prm.Kv_joint[i] = kv_joint[i]; }
  if (adaptive_rho_interval >= 0) h->prm.adaptive_rho_interval = adaptive_rho_interval;
  if (max_iter > 0) h->prm.max_iter = max_iter;
  if (gjk_tol > 0) { h->prm.gjk_tol = gjk_tol; h->prm.epa_tol = gjk_tol; }
}

// ---- stage: updateState + frame getters + manipulability (all arrays batch-major / AoS)
}  // extern "C"
template <int NV>
static int emu_update_and_get_t(EmuHandle* h, int frame_id, int B, const double* q, const double* qd, double* pose, double* J,
                       double* Jdot, double* vel, double* M, double* Minv, double* g, double* nle, double* oMi,
                       double* mani, double* mgrad, double* mgraddot) {
  const int n = h->hm.dev.nv;
  const DrcFrame fr = make_frame(h->hm, frame_id);
  Cache c(n, B);
  JobIO io;
  std::memset(&io, 0, sizeof io);
  io.B = B; io.q = q; io.sq = aos(n); io.qd = qd; io.sqd = aos(n);
  c.bind(io);
  io.pose = pose; io.spose = aos(12); io.J = J; io.sJ = aos(6 * n); io.Jdot = Jdot; io.sJd = aos(6 * n); io.vel = vel; io.svel = aos(6);
  io.mani = mani; io.mani_grad = mgrad; io.smg = aos(n); io.mani_graddot = mgraddot; io.smgd = aos(n);
  run_job<NV, true, F_DYN | F_STORE | F_FRAME_OUT | F_MANIP_OUT | F_GRADDOT>(h, fr, io);
  for (int b = 0; b < B; ++b) {
    for (int i = 0; i < n * n; ++i) { if (M) M[b * n * n + i] = c.M[i * B + b]; if (Minv) Minv[b * n * n + i] = c.Minv[i * B + b]; }
    for (int i = 0; i < n; ++i) { if (g) g[b * n + i] = c.g[i * B + b]; if (nle) nle[b * n + i] = c.nle[i * B + b]; }
    if (oMi) for (int i = 0; i < 12 * n; ++i) oMi[b * 12 * n + i] = c.oMi[i * B + b];
  }
  return 0;
}
extern "C" {
int emu_update_and_get(EmuHandle* h, int frame_id, int B, const double* q, const double* qd, double* pose, double* J,
                       double* Jdot, double* vel, double* M, double* Minv, double* g, double* nle, double* oMi,
                       double* mani, double* mgrad, double* mgraddot) {
  if (!h->hm.chain) return -1;
  if (h->hm.dev.nv == 7) return emu_update_and_get_t<7>(h, frame_id, B, q, qd, pose, J, Jdot, vel, M, Minv, g, nle, oMi, mani, mgrad, mgraddot);
  if (h->hm.dev.nv == 6) return emu_update_and_get_t<6>(h, frame_id, B, q, qd, pose, J, Jdot, vel, M, Minv, g, nle, oMi, mani, mgrad, mgraddot);
  return -1;
}

// ---- stage: min self-distance with gradients
}  // extern "C"
template <int NV>
static int emu_min_distance_t(EmuHandle* h, int B, const double* q, const double* qd, double* dist, double* grad, double* grad_dot,
                     int* pair, double* witness) {
  const int n = h->hm.dev.nv;
  Cache c(n, B);
  JobIO io;
  std::memset(&io, 0, sizeof io);
  io.B = B; io.q = q; io.sq = aos(n); io.qd = qd; io.sqd = aos(n);
  c.bind(io);
  DrcFrame fr = make_frame(h->hm, 0);
  run_job<NV, true, F_STORE>(h, fr, io);
  CollisionIO cio;
  std::memset(&cio, 0, sizeof cio);
  cio.B = B; cio.c_q = c.q.data(); cio.c_qd = c.qd.data(); cio.c_oMi = c.oMi.data(); cio.Bc = B; cio.mode = 0;
  cio.dist = dist; cio.grad = grad; cio.sgrad = aos(n); cio.grad_dot = grad_dot; cio.sgd = aos(n); cio.pair_out = pair; cio.witness = witness;
  std::vector<int> flag, pr; std::vector<unsigned long long> mask; std::vector<double> ds, wt;
  run_collision<NV, true>(h, cio, flag, mask, ds, pr, wt);
  int nepa = 0;
  for (int b = 0; b < B; ++b) nepa += flag[b];
  return nepa;
}
extern "C" {
int emu_min_distance(EmuHandle* h, int B, const double* q, const double* qd, double* dist, double* grad, double* grad_dot,
                     int* pair, double* witness) {
  if (!h->hm.chain) return -1;
  if (h->hm.dev.nv == 7) return emu_min_distance_t<7>(h, B, q, qd, dist, grad, grad_dot, pair, witness);
  if (h->hm.dev.nv == 6) return emu_min_distance_t<6>(h, B, q, qd, dist, grad, grad_dot, pair, witness);
  return -1;
}

// ---- full control cycle: updateState + QPIKStep / QPIDStep (mode 1 / 3) or QPIK / QPID with the
//      desired task signal given in xdot_target (mode 0 / 2)
}  // extern "C"
template <int NV>
static int emu_cycle_t(EmuHandle* h, int mode, int frame_id, int B, const double* q, const double* qd, const double* x_target,
              const double* xdot_target, double* out, int* status, int* iters, double* qp_x, double* qp_records, double* qp_y,
              bool warm = false) {
  const int n = h->hm.dev.nv;
  const DrcFrame fr = make_frame(h->hm, frame_id);
  Cache c(n, B);
  JobIO io;
  std::memset(&io, 0, sizeof io);
  io.B = B; io.q = q; io.sq = aos(n); io.qd = qd; io.sqd = aos(n);
  io.x_target = x_target; io.sxt = aos(12); io.xdot_target = xdot_target; io.sxd = aos(6);
  c.bind(io);
  const bool ID = mode >= 2;
  const int stride = ID ? QpidCfg<NV>::STRIDE : QpikCfg<NV>::STRIDE;
  std::vector<double> rec((size_t)stride * B, 0.0);
  io.qp = rec.data();
  if (mode == 0) run_job<NV, true, F_DYN | F_STORE | F_QPIK>(h, fr, io);
  else if (mode == 1) run_job<NV, true, F_DYN | F_STORE | F_QPIK | F_STEP>(h, fr, io);
  else if (mode == 2) run_job<NV, true, F_DYN | F_STORE | F_QPID>(h, fr, io);
  else run_job<NV, true, F_DYN | F_STORE | F_QPID | F_STEP>(h, fr, io);
  CollisionIO cio;
  std::memset(&cio, 0, sizeof cio);
  cio.B = B; cio.c_q = c.q.data(); cio.c_qd = c.qd.data(); cio.c_oMi = c.oMi.data(); cio.Bc = B;
  cio.mode = ID ? 2 : 1; cio.qp = rec.data(); cio.qp_stride = stride;
  cio.qp_row_off = (ID ? QpidCfg<NV>::OFF_ROW : QpikCfg<NV>::OFF_ROW) + (n + 1);
  std::vector<int> flag, pr; std::vector<unsigned long long> mask; std::vector<double> ds, wt;
  run_collision<NV, true>(h, cio, flag, mask, ds, pr, wt);
  if (qp_records) std::memcpy(qp_records, rec.data(), rec.size() * sizeof(double));
  SolveIO sio;
  std::memset(&sio, 0, sizeof sio);
  sio.B = B; sio.qp = rec.data(); sio.out = out; sio.sout = aos(n); sio.status = status; sio.iters = iters;
  sio.c_g = c.g.data(); sio.Bc = B; sio.qp_x = qp_x; sio.qp_y = qp_y;
  if (warm) { sio.warm_x = qp_x; sio.warm_y = qp_y; }   // in place: read at the start of the solve, written by the emit stage
  if (ID) run_solve<QpidCfg<NV>, true>(h, sio, (1u << NV) - 1u);
  else run_solve<QpikCfg<NV>, false>(h, sio, (1u << NV) - 1u);
  return 0;
}
extern "C" {
int emu_cycle(EmuHandle* h, int mode, int frame_id, int B, const double* q, const double* qd, const double* x_target,
              const double* xdot_target, double* out, int* status, int* iters, double* qp_x, double* qp_records) {
  if (!h->hm.chain) return -1;
  if (h->hm.dev.nv == 7) return emu_cycle_t<7>(h, mode, frame_id, B, q, qd, x_target, xdot_target, out, status, iters, qp_x, qp_records, nullptr);
  if (h->hm.dev.nv == 6) return emu_cycle_t<6>(h, mode, frame_id, B, q, qd, x_target, xdot_target, out, status, iters, qp_x, qp_records, nullptr);
  return -1;
}
// same, plus the unscaled dual vector in structured order (SolveIO::qp_y)
int emu_cycle_xy(EmuHandle* h, int mode, int frame_id, int B, const double* q, const double* qd, const double* x_target,
                 const double* xdot_target, double* out, int* status, int* iters, double* qp_x, double* qp_y) {
  if (!h->hm.chain) return -1;
  if (h->hm.dev.nv == 7) return emu_cycle_t<7>(h, mode, frame_id, B, q, qd, x_target, xdot_target, out, status, iters, qp_x, nullptr, qp_y);
  if (h->hm.dev.nv == 6) return emu_cycle_t<6>(h, mode, frame_id, B, q, qd, x_target, xdot_target, out, status, iters, qp_x, nullptr, qp_y);
  return -1;
}
// warm-started step cycle: (qp_x, qp_y) hold the previous tick's solution on entry (zeros = cold) and this tick's on return
int emu_cycle_warm(EmuHandle* h, int mode, int frame_id, int B, const double* q, const double* qd, const double* x_target,
                   const double* xdot_target, double* out, int* status, int* iters, double* qp_x, double* qp_y) {
  if (!h->hm.chain || h->hm.dev.nv != 7) return -1;
  return emu_cycle_t<7>(h, mode, frame_id, B, q, qd, x_target, xdot_target, out, status, iters, qp_x, nullptr, qp_y, true);
}
int emu_qp_stride(int mode) { return mode >= 2 ? QpidCfg<7>::STRIDE : QpikCfg<7>::STRIDE; }  // FR3 records

// ---- CLIKStep (mode 0) / OSFStep (mode 1) / OSF (mode 2) / joint PD torque (mode 3: aux=q_t, aux2=qd_t)
}  // extern "C"
template <int NV>
static int emu_taskspace_t(EmuHandle* h, int mode, int frame_id, int B, const double* q, const double* qd, const double* x_target,
                  const double* xdot_target, const double* aux, const double* aux2, double* out) {
  const int n = h->hm.dev.nv;
  const DrcFrame fr = make_frame(h->hm, frame_id);
  Cache c(n, B);
  JobIO io;
  std::memset(&io, 0, sizeof io);
  io.B = B; io.q = q; io.sq = aos(n); io.qd = qd; io.sqd = aos(n);
  io.x_target = x_target; io.sxt = aos(12); io.xdot_target = xdot_target; io.sxd = aos(6);
  io.aux = aux; io.saux = aos(n); io.aux2 = aux2; io.saux2 = aos(n); io.out = out; io.sout = aos(n);
  c.bind(io);
  if (mode == 0) run_job<NV, true, F_CLIK | F_STEP>(h, fr, io);
  else if (mode == 1) run_job<NV, true, F_DYN | F_OSF | F_STEP>(h, fr, io);
  else if (mode == 2) run_job<NV, true, F_DYN | F_OSF>(h, fr, io);
  else run_job<NV, true, F_DYN | F_TORQUE>(h, fr, io);
  return 0;
}
extern "C" {
int emu_taskspace(EmuHandle* h, int mode, int frame_id, int B, const double* q, const double* qd, const double* x_target,
                  const double* xdot_target, const double* aux, const double* aux2, double* out) {
  if (!h->hm.chain) return -1;
  if (h->hm.dev.nv == 7) return emu_taskspace_t<7>(h, mode, frame_id, B, q, qd, x_target, xdot_target, aux, aux2, out);
  if (h->hm.dev.nv == 6) return emu_taskspace_t<6>(h, mode, frame_id, B, q, qd, x_target, xdot_target, aux, aux2, out);
  return -1;
}

// ---- narrow phase of the product on one shape pair (type, prm[3], pose12), world frame
double emu_shape_distance(EmuHandle* h, int ta, const double* pa, const double* Ta, int tb, const double* pb, const double* Tb,
                          double* wa, double* wb, int* info) {
  auto mk = [](int t, const double* prm, const double* T) {
    Prim s;
    s.type = t; s.r = prm[0]; s.h = prm[1]; s.hb = v3(prm[0], prm[1], prm[2]);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) s.R.m[3 * i + j] = T[4 * i + j];
    s.c = v3(T[3], T[7], T[11]);
    s.a = v3(T[2], T[6], T[10]);
    return s;
  };
  const Prim A = mk(ta, pa, Ta), B = mk(tb, pb, Tb);
  PairResult r;
  info[0] = info[1] = 0;
  if (has_closed_form(ta, tb)) r = closed_form_distance(A, B);
  else {
    GjkOut g;
    gjk_distance(A, B, h->prm.gjk_tol, h->prm.gjk_max_iter, g);
    r.d = g.dist; r.pa = g.pa; r.pb = g.pb;
    info[0] = g.iters;
    if (g.intersect) { epa_penetration(A, B, g, h->prm.epa_tol, h->prm.epa_max_iter, r); info[1] = 1; }
  }
  wa[0] = r.pa.x; wa[1] = r.pa.y; wa[2] = r.pa.z; wb[0] = r.pb.x; wb[1] = r.pb.y; wb[2] = r.pb.z;
  return r.d;
}
double emu_pair_lower_bound(int ta, const double* pa, const double* Ta, int tb, const double* pb, const double* Tb) {
  auto mk = [](int t, const double* prm, const double* T) {
    Prim s;
    s.type = t; s.r = prm[0]; s.h = prm[1]; s.hb = v3(prm[0], prm[1], prm[2]);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) s.R.m[3 * i + j] = T[4 * i + j];
    s.c = v3(T[3], T[7], T[11]);
    s.a = v3(T[2], T[6], T[10]);
    return s;
  };
  return pair_lower_bound(mk(ta, pa, Ta), mk(tb, pb, Tb));
}


// ---- mobile manipulator (rows a15-a18): attach a base, state update with actuated quantities, whole-body QP cycles
int emu_moma_attach(EmuHandle* h, int drive_type, double wheel_radius, double base_width, double wheel_offset, int w,
                    const double* roller, const double* b2w_x, const double* b2w_y, const double* b2w_ang, int vs, int ms,
                    int bs, int ams, int abs_, char* err, int errlen) {
  try {
    MobileParam p;
    p.drive_type = drive_type; p.wheel_radius = wheel_radius; p.base_width = base_width; p.wheel_offset = wheel_offset;
    const int np = drive_type == kCaster ? w / 2 : w;
    for (int i = 0; i < w && roller; ++i) p.roller_angles.push_back(roller[i]);
    for (int i = 0; i < np && b2w_x; ++i) { p.b2w_x.push_back(b2w_x[i]); p.b2w_y.push_back(b2w_y[i]); }
    for (int i = 0; i < w && b2w_ang; ++i) p.b2w_angles.push_back(b2w_ang[i]);
    attach_mobile_base(h->hm, p, vs, ms, bs, ams, abs_);
    return 0;
  } catch (const std::exception& e) {
    if (err && errlen > 0) { std::strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; }
    return -1;
  }
}
void emu_moma_base_jacobian(EmuHandle* h, double* J /* 3 x w */) {
  const DrcModelDev& d = h->hm.dev;
  for (int r = 0; r < 3; ++r) for (int k = 0; k < d.wheel_num; ++k) J[r * d.wheel_num + k] = d.J_mobile[r][k];
}

}  // extern "C"

struct MomaCache : Cache {
  std::vector<double> Mact, Minvact, gact, nleact;
  MomaCache(int nv, int act, int B) : Cache(nv, B), Mact(act * act * B), Minvact(act * act * B), gact(act * B), nleact(act * B) {}
  void bind(JobIO& io) {
    Cache::bind(io);
    io.c_Mact = Mact.data(); io.c_Minvact = Minvact.data(); io.c_gact = gact.data(); io.c_nleact = nleact.data();
  }
};

template <int NV, int W>
static int moma_state_t(EmuHandle* h, int frame_id, int B, const double* q, const double* qd, double* pose, double* Jt, double* Jtd,
                        double* vel, double* Mact, double* Minvact, double* gact, double* nleact, double* mani, double* mgrad,
                        double* mgraddot) {
  constexpr int ACT = NV - 3, MANI = NV - 3 - W;
  const DrcFrame fr = make_frame(h->hm, frame_id);
  MomaCache c(NV, ACT, B);
  JobIO io;
  std::memset(&io, 0, sizeof io);
  io.B = B; io.q = q; io.sq = aos(NV); io.qd = qd; io.sqd = aos(NV);
  c.bind(io);
  io.pose = pose; io.spose = aos(12); io.J = Jt; io.sJ = aos(6 * ACT); io.Jdot = Jtd; io.sJd = aos(6 * ACT); io.vel = vel; io.svel = aos(6);
  io.mani = mani; io.mani_grad = mgrad; io.smg = aos(MANI); io.mani_graddot = mgraddot; io.smgd = aos(MANI);
  for (int b = 0; b < B; ++b)
    robot_job<NV, false, F_DYN | F_STORE | F_FRAME_OUT | F_MANIP_OUT | F_GRADDOT | F_MOMA, W>(h->hm.dev, h->prm, fr, io, b);
  for (int b = 0; b < B; ++b) {
    for (int i = 0; i < ACT * ACT; ++i) { if (Mact) Mact[b * ACT * ACT + i] = c.Mact[i * B + b]; if (Minvact) Minvact[b * ACT * ACT + i] = c.Minvact[i * B + b]; }
    for (int i = 0; i < ACT; ++i) { if (gact) gact[b * ACT + i] = c.gact[i * B + b]; if (nleact) nleact[b * ACT + i] = c.nleact[i * B + b]; }
  }
  return 0;
}
extern "C" {
int emu_moma_state(EmuHandle* h, int frame_id, int B, const double* q, const double* qd, double* pose, double* Jt, double* Jtd,
                   double* vel, double* Mact, double* Minvact, double* gact, double* nleact, double* mani, double* mgrad,
                   double* mgraddot) {
  const DrcModelDev& d = h->hm.dev;
  if (d.drive_type == kNoBase) return -1;
  if (d.nv == 12 && d.wheel_num == 2) return moma_state_t<12, 2>(h, frame_id, B, q, qd, pose, Jt, Jtd, vel, Mact, Minvact, gact, nleact, mani, mgrad, mgraddot);
  if (d.nv == 14 && d.wheel_num == 4) return moma_state_t<14, 4>(h, frame_id, B, q, qd, pose, Jt, Jtd, vel, Mact, Minvact, gact, nleact, mani, mgrad, mgraddot);
  return -2;
}

}  // extern "C"

template <int NV, int W>
static int moma_cycle_t(EmuHandle* h, int mode, int frame_id, int B, const double* q, const double* qd, const double* x_target,
                        const double* xdot_target, double* out, double* out2, int* status, int* iters) {
  constexpr int ACT = NV - 3, MANI = NV - 3 - W;
  const DrcModelDev& d = h->hm.dev;
  const DrcFrame fr = make_frame(h->hm, frame_id);
  MomaCache c(NV, ACT, B);
  JobIO io;
  std::memset(&io, 0, sizeof io);
  io.B = B; io.q = q; io.sq = aos(NV); io.qd = qd; io.sqd = aos(NV);
  io.x_target = (mode == 1 || mode == 3) ? x_target : nullptr; io.sxt = aos(12); io.xdot_target = xdot_target; io.sxd = aos(6);
  c.bind(io);
  const bool ID = mode >= 2;
  const int stride = ID ? MomaIdCfg<ACT>::STRIDE : MomaIkCfg<ACT>::STRIDE;
  std::vector<double> rec((size_t)stride * B, 0.0);
  io.qp = rec.data();
  // the product's launch sequence (csrc/drc_moma.cu moma_qp): joint placements -> cache; QP record from the cached state (QPIK:
  // kinematics only, QPID: M~ and g~ only); the dynamics-only job completes the cache
  for (int b = 0; b < B; ++b) {
    robot_job<NV, false, F_STORE>(d, h->prm, fr, io, b);
    if (!ID) robot_job<NV, false, F_FROM_CACHE | F_QPIK | F_MOMA, W>(d, h->prm, fr, io, b);
    else robot_job<NV, false, F_DYN | F_DYN_LIGHT | F_FROM_CACHE | F_QPID | F_MOMA, W>(d, h->prm, fr, io, b);
    robot_job<NV, false, F_DYN | F_FROM_CACHE | F_MOMA, W>(d, h->prm, fr, io, b);
  }
  CollisionIO cio;
  std::memset(&cio, 0, sizeof cio);
  cio.B = B; cio.c_q = c.q.data(); cio.c_qd = c.qd.data(); cio.c_oMi = c.oMi.data(); cio.Bc = B;
  cio.mode = ID ? 2 : 1; cio.qp = rec.data(); cio.qp_stride = stride;
  cio.qp_row_off = (ID ? MomaIdCfg<ACT>::OFF_ROW : MomaIkCfg<ACT>::OFF_ROW) + (ACT + 1);
  cio.row_n = ACT; cio.row_col0 = d.act_mani_start; cio.src0 = d.mani_start; cio.nsrc = MANI;
  std::vector<int> flag, pr; std::vector<unsigned long long> mask; std::vector<double> ds, wt;
  run_collision<NV, false>(h, cio, flag, mask, ds, pr, wt);
  SolveIO sio;
  std::memset(&sio, 0, sizeof sio);
  sio.B = B; sio.qp = rec.data(); sio.out = out; sio.sout = aos(ACT); sio.out2 = out2; sio.sout2 = aos(ACT); sio.status = status; sio.iters = iters;
  sio.c_g = c.gact.data(); sio.Bc = B;
  const unsigned mani_mask = ((1u << MANI) - 1u) << d.act_mani_start;
  if (ID) run_solve<MomaIdCfg<ACT>, true>(h, sio, mani_mask);
  else run_solve<MomaIkCfg<ACT>, false>(h, sio, mani_mask);
  return 0;
}
extern "C" {
// mode 0 QPIK(xdot_des) 1 QPIKStep 2 QPID(xddot_des) 3 QPIDStep; out: eta* (modes 0/1) or tau* (2/3), out2: eta_dot* (2/3)
int emu_moma_cycle(EmuHandle* h, int mode, int frame_id, int B, const double* q, const double* qd, const double* x_target,
                   const double* xdot_target, double* out, double* out2, int* status, int* iters) {
  const DrcModelDev& d = h->hm.dev;
  if (d.drive_type == kNoBase) return -1;
  if (d.nv == 12 && d.wheel_num == 2) return moma_cycle_t<12, 2>(h, mode, frame_id, B, q, qd, x_target, xdot_target, out, out2, status, iters);
  if (d.nv == 14 && d.wheel_num == 4) return moma_cycle_t<14, 4>(h, mode, frame_id, B, q, qd, x_target, xdot_target, out, out2, status, iters);
  return -2;
}

}  // extern "C"

// ---- mobile base alone (Mobile::RobotData / Mobile::RobotController): the bodies of k_mobile_fk / k_mobile_ik
extern "C" int emu_mobile(int drive_type, double wheel_radius, double base_width, double wheel_offset, double max_lin_speed,
                          double max_ang_speed, int w, const double* roller, const double* b2w_x, const double* b2w_y,
                          const double* b2w_ang, int fk, int saturate, int B, const double* wheel_pos, const double* in, double* J,
                          double* out) {
  try {
    MobileParam p;
    p.drive_type = drive_type; p.wheel_radius = wheel_radius; p.base_width = base_width; p.wheel_offset = wheel_offset;
    const int np = drive_type == kCaster ? w / 2 : w;
    for (int i = 0; i < w && roller; ++i) p.roller_angles.push_back(roller[i]);
    for (int i = 0; i < np && b2w_x; ++i) { p.b2w_x.push_back(b2w_x[i]); p.b2w_y.push_back(b2w_y[i]); }
    for (int i = 0; i < w && b2w_ang; ++i) p.b2w_angles.push_back(b2w_ang[i]);
    MobileDev d;
    std::memset(&d, 0, sizeof d);
    d.drive_type = drive_type; d.wheel_num = w; d.wheel_radius = wheel_radius; d.base_width = base_width; d.wheel_offset = wheel_offset;
    d.max_lin_speed = max_lin_speed; d.max_ang_speed = max_ang_speed;
    for (size_t i = 0; i < p.b2w_x.size(); ++i) { d.b2w_x[i] = p.b2w_x[i]; d.b2w_y[i] = p.b2w_y[i]; }
    mobile_constant_jacobians(p, w, d.J_fk, d.J_ik);
    MobileIO io;
    std::memset(&io, 0, sizeof io);
    io.B = B; io.wheel_pos = wheel_pos; io.swp = aos(w); io.J = J; io.sj = aos(3 * w); io.out = out; io.saturate = saturate;
    if (fk) { io.wheel_vel = in; io.swv = aos(w); io.so = aos(3); }
    else { io.base_vel = in; io.sbv = aos(3); io.so = aos(w); }
    for (int b = 0; b < B; ++b) {
      if (fk) mobile_fk_job(d, io, b); else mobile_ik_job(d, io, b);
    }
    return 0;
  } catch (const std::exception&) {
    return -1;
  }
}
