// TEST / MEASUREMENT INFRASTRUCTURE (never linked into libdrc_b200.so): algorithmic fp64 flop counts of the product's kernel
// bodies.  The very same DRC_HD routines that nvcc compiles into the kernels (drc_kin.h, drc_geom.h, drc_qp.h, drc_cycle.h) are
// compiled here with `double` replaced by a COUNTING scalar; every +, -, *, /, sqrt, sin, cos, ... on it increments the counter of
// the current phase (DRC_PHASE markers in the bodies).  Multiplications by / additions of an exact structural zero are NOT
// counted: the kernels are branch-free and multiply absent rows by zero coefficients, which is not algorithmic work.
// tools/count_flops.py runs this on the benchmark inputs and writes bench/flops_model.json, which bench.py loads.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

enum { PH_OTHER = 0, PH_KIN, PH_DYN, PH_MANI, PH_BUILD, PH_COLLISION, PH_QP_LOAD, PH_QP_SCALE, PH_QP_FACTOR, PH_QP_ITER, PH_QP_CHECK, PH_QP_EMIT, PH_N };
static int g_phase = PH_OTHER;
static long long g_flops[PH_N];
static long long g_rho_updates = 0;
#define DRC_PHASE(x) (g_phase = (x))

struct CountReal {
  double v;
  CountReal() = default;
  constexpr CountReal(double x) : v(x) {}
  constexpr CountReal(int x) : v(x) {}
  constexpr CountReal(float x) : v(x) {}
  explicit constexpr operator float() const { return (float)v; }
  explicit constexpr operator int() const { return (int)v; }
  explicit constexpr operator bool() const { return v != 0; }
  explicit constexpr operator long long() const { return (long long)v; }
};
static inline void tick(int n = 1) { g_flops[g_phase] += n; }
#define BINOP(op, ZERO_RULE)                                                                                                  \
  static inline CountReal operator op(CountReal a, CountReal b) { if (!(ZERO_RULE)) tick(); return CountReal(a.v op b.v); }        \
  static inline CountReal operator op(CountReal a, double b) { return a op CountReal(b); }                                          \
  static inline CountReal operator op(double a, CountReal b) { return CountReal(a) op b; }                                          \
  static inline CountReal operator op(CountReal a, int b) { return a op CountReal((double)b); }                                     \
  static inline CountReal operator op(int a, CountReal b) { return CountReal((double)a) op b; }
BINOP(+, a.v == 0.0 || b.v == 0.0)
BINOP(-, b.v == 0.0)
BINOP(*, a.v == 0.0 || b.v == 0.0 || a.v == 1.0 || b.v == 1.0 || a.v == -1.0 || b.v == -1.0)
BINOP(/, a.v == 0.0 || b.v == 1.0)
static inline CountReal operator-(CountReal a) { return CountReal(-a.v); }
static inline CountReal operator+(CountReal a) { return a; }
static inline CountReal& operator+=(CountReal& a, CountReal b) { a = a + b; return a; }
static inline CountReal& operator-=(CountReal& a, CountReal b) { a = a - b; return a; }
static inline CountReal& operator*=(CountReal& a, CountReal b) { a = a * b; return a; }
static inline CountReal& operator/=(CountReal& a, CountReal b) { a = a / b; return a; }
#define CMP(op)                                                                                          \
  static inline bool operator op(CountReal a, CountReal b) { return a.v op b.v; }                           \
  static inline bool operator op(CountReal a, double b) { return a.v op b; }                                \
  static inline bool operator op(double a, CountReal b) { return a op b.v; }
CMP(<) CMP(>) CMP(<=) CMP(>=) CMP(==) CMP(!=)
static inline CountReal sqrt(CountReal a) { tick(); return CountReal(std::sqrt(a.v)); }
static inline CountReal fabs(CountReal a) { return CountReal(std::fabs(a.v)); }
static inline CountReal sin(CountReal a) { tick(); return CountReal(std::sin(a.v)); }
static inline CountReal cos(CountReal a) { tick(); return CountReal(std::cos(a.v)); }
static inline CountReal tan(CountReal a) { tick(); return CountReal(std::tan(a.v)); }
static inline CountReal acos(CountReal a) { tick(); return CountReal(std::acos(a.v)); }
static inline CountReal atan2(CountReal a, CountReal b) { tick(); return CountReal(std::atan2(a.v, b.v)); }
static inline void sincos(CountReal a, CountReal* s, CountReal* c) { tick(2); s->v = std::sin(a.v); c->v = std::cos(a.v); }
static inline bool isfinite(CountReal a) { return std::isfinite(a.v); }
static inline CountReal fmin(CountReal a, CountReal b) { return a.v < b.v ? a : b; }
static inline CountReal fmax(CountReal a, CountReal b) { return a.v > b.v ? a : b; }
static inline CountReal fabsf(CountReal a) { return CountReal(std::fabs(a.v)); }
namespace std {
static inline CountReal sqrt(CountReal a) { return ::sqrt(a); }
static inline CountReal fabs(CountReal a) { return ::fabs(a); }
static inline CountReal sin(CountReal a) { return ::sin(a); }
static inline CountReal cos(CountReal a) { return ::cos(a); }
static inline CountReal tan(CountReal a) { return ::tan(a); }
static inline CountReal max(CountReal a, CountReal b) { return a.v > b.v ? a : b; }
static inline CountReal min(CountReal a, CountReal b) { return a.v < b.v ? a : b; }
}  // namespace std

#define double CountReal
#include "../../dyros_robot_controller_b200/csrc/drc_cycle.h"
#include "../../dyros_robot_controller_b200/csrc/model.h"
#undef double
static_assert(sizeof(CountReal) == sizeof(double) && std::is_trivially_copyable<CountReal>::value, "layout must match the double build of model.cpp");

using namespace drc;
typedef CountReal R;

struct Cache {
  int n, B;
  std::vector<R> q, qd, oMi, M, Minv, g, nle, Mact, Minvact, gact, nleact;
  Cache(int n_, int B_) : n(n_), B(B_), q(n_ * B_), qd(n_ * B_), oMi(12 * n_ * B_), M(n_ * n_ * B_), Minv(n_ * n_ * B_), g(n_ * B_), nle(n_ * B_),
                          Mact(n_ * n_ * B_), Minvact(n_ * n_ * B_), gact(n_ * B_), nleact(n_ * B_) {}
  void bind(JobIO& io) {
    io.c_q = q.data(); io.c_qd = qd.data(); io.c_oMi = oMi.data(); io.c_M = M.data(); io.c_Minv = Minv.data(); io.c_g = g.data(); io.c_nle = nle.data();
    io.c_Mact = Mact.data(); io.c_Minvact = Minvact.data(); io.c_gact = gact.data(); io.c_nleact = nleact.data(); io.Bc = B;
  }
};

template <class Cfg, bool ID>
static void run_solve(const DrcParams& prm, const SolveIO& io, unsigned unit_mask) {
  const QpOptions o = qp_options(prm, unit_mask);
  std::vector<GroupShared<Cfg>> sh(Cfg::NG);
  std::unique_ptr<WarpEmu<Cfg>> w(new WarpEmu<Cfg>);
  for (int b0 = 0; b0 < io.B; b0 += Cfg::NG) {
    int robots[8];
    for (int g = 0; g < Cfg::NG; ++g) robots[g] = (b0 + g < io.B) ? b0 + g : -1;
    w->sh = sh.data();
    for (int t = 0; t < 32; ++t) lane_assign<Cfg>(w->Ls[t], t);
    solve_and_emit<Cfg, ID>(*w, robots, io, o);
    for (int g = 0; g < Cfg::NG; ++g) if (robots[g] >= 0) g_rho_updates += sh[g].rho_updates;
  }
}

static DrcFrame make_frame(const HostModel& hm, int fid) {
  DrcFrame f;
  f.parent = hm.frames[fid].parent;
  for (int i = 0; i < 9; ++i) f.R[i] = hm.frames[fid].R[i];
  for (int i = 0; i < 3; ++i) f.p[i] = hm.frames[fid].p[i];
  return f;
}

// one fused control cycle (updateState + QPIKStep / QPIDStep) of B robots of a serial chain (NV = 7 / 6)
template <int NV>
static void cycle_chain(const HostModel& hm, const DrcParams& prm, bool id, int fid, int B, const double* q, const double* qd, const double* xt, const double* xd,
                        int* iters, int* status) {
  const DrcFrame fr = make_frame(hm, fid);
  Cache c(NV, B);
  std::vector<R> rq(q, q + (size_t)B * NV), rqd(qd, qd + (size_t)B * NV), rxt(xt, xt + (size_t)B * 12), rxd(xd, xd + (size_t)B * 6), out((size_t)B * NV);
  JobIO io;
  std::memset(&io, 0, sizeof io);
  io.B = B; io.q = rq.data(); io.sq = aos(NV); io.qd = rqd.data(); io.sqd = aos(NV);
  io.x_target = rxt.data(); io.sxt = aos(12); io.xdot_target = rxd.data(); io.sxd = aos(6);
  c.bind(io);
  const int stride = id ? QpidCfg<NV>::STRIDE : QpikCfg<NV>::STRIDE;
  std::vector<R> rec((size_t)stride * B, R(0.0));
  io.qp = rec.data();
  for (int b = 0; b < B; ++b) {
    if (id) robot_job<NV, true, F_DYN | F_STORE | F_QPID | F_STEP>(hm.dev, prm, fr, io, b);
    else robot_job<NV, true, F_DYN | F_STORE | F_QPIK | F_STEP>(hm.dev, prm, fr, io, b);
  }
  CollisionIO cio;
  std::memset(&cio, 0, sizeof cio);
  std::vector<int> flag(B, 0), pr(B, 0);
  std::vector<unsigned long long> mask(B, 0ull);
  std::vector<R> ds(B), wt((size_t)6 * B);
  cio.B = B; cio.c_q = c.q.data(); cio.c_qd = c.qd.data(); cio.c_oMi = c.oMi.data(); cio.Bc = B;
  cio.mode = id ? 2 : 1; cio.qp = rec.data(); cio.qp_stride = stride;
  cio.qp_row_off = (id ? QpidCfg<NV>::OFF_ROW : QpikCfg<NV>::OFF_ROW) + (NV + 1);
  cio.epa_flag = flag.data(); cio.cand_mask = mask.data(); cio.dist = ds.data(); cio.pair_out = pr.data(); cio.witness = wt.data();
  g_phase = PH_COLLISION;
  for (int b = 0; b < B; ++b) collision_job<NV, true>(hm.dev, hm.dev.geom, prm, cio, b);
  for (int b = 0; b < B; ++b) collision_epa_job<NV, true>(hm.dev, prm, cio, b);
  SolveIO sio;
  std::memset(&sio, 0, sizeof sio);
  sio.B = B; sio.qp = rec.data(); sio.out = out.data(); sio.sout = aos(NV); sio.status = status; sio.iters = iters; sio.c_g = c.g.data(); sio.Bc = B;
  if (id) run_solve<QpidCfg<NV>, true>(prm, sio, (1u << NV) - 1u);
  else run_solve<QpikCfg<NV>, false>(prm, sio, (1u << NV) - 1u);
  g_phase = PH_OTHER;
}

// CLIKStep + OSFStep (config 2): no QP
template <int NV>
static void taskspace_chain(const HostModel& hm, const DrcParams& prm, int fid, int B, const double* q, const double* qd, const double* xt, const double* xd) {
  const DrcFrame fr = make_frame(hm, fid);
  Cache c(NV, B);
  std::vector<R> rq(q, q + (size_t)B * NV), rqd(qd, qd + (size_t)B * NV), rxt(xt, xt + (size_t)B * 12), rxd(xd, xd + (size_t)B * 6), o1((size_t)B * NV), o2((size_t)B * NV);
  JobIO io;
  std::memset(&io, 0, sizeof io);
  io.B = B; io.q = rq.data(); io.sq = aos(NV); io.qd = rqd.data(); io.sqd = aos(NV);
  io.x_target = rxt.data(); io.sxt = aos(12); io.xdot_target = rxd.data(); io.sxd = aos(6);
  io.out = o1.data(); io.out2 = o2.data(); io.sout = aos(NV);
  c.bind(io);
  for (int b = 0; b < B; ++b) robot_job<NV, true, F_DYN | F_STORE>(hm.dev, prm, fr, io, b);
  for (int b = 0; b < B; ++b) robot_job<NV, true, F_FROM_CACHE | F_CLIK | F_STEP>(hm.dev, prm, fr, io, b);
  io.out = o2.data();
  for (int b = 0; b < B; ++b) robot_job<NV, true, F_FROM_CACHE | F_OSF | F_STEP>(hm.dev, prm, fr, io, b);
  g_phase = PH_OTHER;
}

// whole-body cycle (NV, W) = (12, 2) / (14, 4)
template <int NV, int W>
static void cycle_moma(const HostModel& hm, const DrcParams& prm, bool id, int fid, int B, const double* q, const double* qd, const double* xt, const double* xd,
                       int* iters, int* status) {
  constexpr int ACT = NV - 3, MANI = NV - 3 - W;
  const DrcFrame fr = make_frame(hm, fid);
  Cache c(NV, B);
  std::vector<R> rq(q, q + (size_t)B * NV), rqd(qd, qd + (size_t)B * NV), rxt(xt, xt + (size_t)B * 12), rxd(xd, xd + (size_t)B * 6), out((size_t)B * ACT), out2((size_t)B * ACT);
  JobIO io;
  std::memset(&io, 0, sizeof io);
  io.B = B; io.q = rq.data(); io.sq = aos(NV); io.qd = rqd.data(); io.sqd = aos(NV);
  io.x_target = rxt.data(); io.sxt = aos(12); io.xdot_target = rxd.data(); io.sxd = aos(6);
  c.bind(io);
  const int stride = id ? MomaIdCfg<ACT>::STRIDE : MomaIkCfg<ACT>::STRIDE;
  std::vector<R> rec((size_t)stride * B, R(0.0));
  io.qp = rec.data();
  for (int b = 0; b < B; ++b) {
    if (id) robot_job<NV, false, F_DYN | F_STORE | F_QPID | F_MOMA, W>(hm.dev, prm, fr, io, b);
    else robot_job<NV, false, F_DYN | F_STORE | F_QPIK | F_MOMA, W>(hm.dev, prm, fr, io, b);
  }
  CollisionIO cio;
  std::memset(&cio, 0, sizeof cio);
  std::vector<int> flag(B, 0), pr(B, 0);
  std::vector<unsigned long long> mask(B, 0ull);
  std::vector<R> ds(B), wt((size_t)6 * B);
  cio.B = B; cio.c_q = c.q.data(); cio.c_qd = c.qd.data(); cio.c_oMi = c.oMi.data(); cio.Bc = B;
  cio.mode = id ? 2 : 1; cio.qp = rec.data(); cio.qp_stride = stride;
  cio.qp_row_off = (id ? MomaIdCfg<ACT>::OFF_ROW : MomaIkCfg<ACT>::OFF_ROW) + (ACT + 1);
  cio.row_n = ACT; cio.row_col0 = hm.dev.act_mani_start; cio.src0 = hm.dev.mani_start; cio.nsrc = MANI;
  cio.epa_flag = flag.data(); cio.cand_mask = mask.data(); cio.dist = ds.data(); cio.pair_out = pr.data(); cio.witness = wt.data();
  g_phase = PH_COLLISION;
  for (int b = 0; b < B; ++b) collision_job<NV, false>(hm.dev, hm.dev.geom, prm, cio, b);
  for (int b = 0; b < B; ++b) collision_epa_job<NV, false>(hm.dev, prm, cio, b);
  SolveIO sio;
  std::memset(&sio, 0, sizeof sio);
  sio.B = B; sio.qp = rec.data(); sio.out = out.data(); sio.sout = aos(ACT); sio.out2 = out2.data(); sio.sout2 = aos(ACT); sio.status = status; sio.iters = iters;
  sio.c_g = c.gact.data(); sio.Bc = B;
  const unsigned mani_mask = ((1u << MANI) - 1u) << hm.dev.act_mani_start;
  if (id) run_solve<MomaIdCfg<ACT>, true>(prm, sio, mani_mask);
  else run_solve<MomaIkCfg<ACT>, false>(prm, sio, mani_mask);
  g_phase = PH_OTHER;
}

extern "C" {
struct FcHandle { HostModel hm; DrcParams prm; };
FcHandle* fc_create(const char* urdf_text, const char* srdf_text) {
  try {
    std::unique_ptr<FcHandle> h(new FcHandle);
    h->hm = compile_model(urdf_text, srdf_text ? srdf_text : "");
    h->hm.bind_hull();
    for (int i = 0; i < kMaxV; ++i) { h->prm.Kp_joint[i] = 400; h->prm.Kv_joint[i] = 40; }
    return h.release();
  } catch (const std::exception& e) {
    std::fprintf(stderr, "fc_create: %s\n", e.what());
    return nullptr;
  }
}
void fc_destroy(FcHandle* h) { delete h; }
int fc_frame_id(FcHandle* h, const char* name) { return h->hm.frame_id(name); }
int fc_attach_base(FcHandle* h, int drive_type, double wheel_radius, double base_width, double wheel_offset, int n_wheels, const double* roller_angles,
                   const double* bx, const double* by, const double* ba, int virtual_start, int mani_start, int mobi_start, int act_mani_start, int act_mobi_start) {
  try {
    MobileParam p;
    p.drive_type = drive_type; p.wheel_radius = wheel_radius; p.base_width = base_width; p.wheel_offset = wheel_offset;
    const int np = drive_type == kCaster ? n_wheels / 2 : n_wheels;
    for (int i = 0; i < n_wheels && roller_angles; ++i) p.roller_angles.push_back(roller_angles[i]);
    for (int i = 0; i < np && bx && by; ++i) { p.b2w_x.push_back(bx[i]); p.b2w_y.push_back(by[i]); }
    for (int i = 0; i < n_wheels && ba; ++i) p.b2w_angles.push_back(ba[i]);
    attach_mobile_base(h->hm, p, virtual_start, mani_start, mobi_start, act_mani_start, act_mobi_start);
    return 0;
  } catch (const std::exception& e) { std::fprintf(stderr, "fc_attach_base: %s\n", e.what()); return -1; }
}
void fc_set_solver(FcHandle* h, int max_iter, int check_termination, int scaling, int adaptive_rho) {
  h->prm.max_iter = max_iter; h->prm.check_termination = check_termination; h->prm.scaling = scaling; h->prm.adaptive_rho = adaptive_rho;
}
void fc_set_task_gains(FcHandle* h, double kp, double kv) { for (int i = 0; i < 6; ++i) { h->prm.Kp_task[i] = kp; h->prm.Kv_task[i] = kv; } }
// kind 0 QPIKStep cycle, 1 QPIDStep cycle, 2 CLIKStep + OSFStep; flops (PH_N long longs) are reset on entry
int fc_run(FcHandle* h, int kind, int frame, int B, const double* q, const double* qd, const double* xt, const double* xd, int* iters, int* status,
           long long* flops) {
  std::memset(g_flops, 0, sizeof g_flops);
  g_rho_updates = 0;
  const DrcModelDev& d = h->hm.dev;
  const int nv = (int)d.nv;
  if (d.drive_type == kNoBase) {
    if (kind == 2) {
      if (nv == 6) taskspace_chain<6>(h->hm, h->prm, frame, B, q, qd, xt, xd);
      else if (nv == 7) taskspace_chain<7>(h->hm, h->prm, frame, B, q, qd, xt, xd);
      else return -1;
    } else if (nv == 7) cycle_chain<7>(h->hm, h->prm, kind == 1, frame, B, q, qd, xt, xd, iters, status);
    else if (nv == 6) cycle_chain<6>(h->hm, h->prm, kind == 1, frame, B, q, qd, xt, xd, iters, status);
    else return -1;
  } else if (nv == 12 && d.wheel_num == 2) cycle_moma<12, 2>(h->hm, h->prm, kind == 1, frame, B, q, qd, xt, xd, iters, status);
  else if (nv == 14 && d.wheel_num == 4) cycle_moma<14, 4>(h->hm, h->prm, kind == 1, frame, B, q, qd, xt, xd, iters, status);
  else return -1;
  std::memcpy(flops, g_flops, sizeof g_flops);
  return 0;
}
int fc_num_phases() { return PH_N; }
long long fc_rho_updates() { return g_rho_updates; }
// pose (top three rows) of `frame` at q: targets for the benchmark inputs
int fc_pose(FcHandle* h, int frame, int B, const double* q, double* pose12) {
  const DrcModelDev& d = h->hm.dev;
  const int nv = (int)d.nv;
  const DrcFrame fr = make_frame(h->hm, frame);
  std::vector<R> rq(q, q + (size_t)B * nv), zero((size_t)B * nv, R(0.0)), out((size_t)B * 12);
  JobIO io;
  std::memset(&io, 0, sizeof io);
  io.B = B; io.q = rq.data(); io.sq = aos(nv); io.qd = zero.data(); io.sqd = aos(nv); io.pose = out.data(); io.spose = aos(12);
  for (int b = 0; b < B; ++b) {
    if (nv == 7) robot_job<7, true, F_FRAME_OUT>(d, h->prm, fr, io, b);
    else if (nv == 6) robot_job<6, true, F_FRAME_OUT>(d, h->prm, fr, io, b);
    else if (nv == 12) robot_job<12, false, F_FRAME_OUT>(d, h->prm, fr, io, b);
    else if (nv == 14) robot_job<14, false, F_FRAME_OUT>(d, h->prm, fr, io, b);
    else return -1;
  }
  for (size_t i = 0; i < out.size(); ++i) pose12[i] = out[i].v;
  return 0;
}
}
