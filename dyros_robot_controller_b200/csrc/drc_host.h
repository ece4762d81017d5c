// drc_b200 -- host-side internals shared by the translation units of libdrc_b200.so: the opaque handle types of the C
// ABI, launch helpers and the host staging used by the drc_host_* entry points.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/drc_b200.h"
#include "drc_kernels.cuh"
#include "model.h"

using namespace drc;
using namespace drc_kernels;

// error channel of the C ABI (drc_last_error); defined in drc_lib.cu
int drc_set_error(int code, const std::string& msg);
static inline int fail(int code, const std::string& msg) { return drc_set_error(code, msg); }
#define CU(x)                                                                                      \
  do {                                                                                             \
    cudaError_t e_ = (x);                                                                          \
    if (e_ != cudaSuccess) return fail(DRC_E_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); \
  } while (0)

struct drc_model {
  HostModel hm;
  std::string verbose;
};

// State cache + QP / collision scratch of one pipeline.  The context owns two: the full batch and a small compact one
// for the priority sub-batch (robots predicted to need many ADMM iterations; run_qp in drc_lib.cu).
struct Scratch {
  double *c_q, *c_qd, *c_oMi, *c_M, *c_Minv, *c_g, *c_nle;
  long long Bc;
  double* qp;
  int* epa_flag; unsigned long long* cand_mask; double* col_dist; int* col_pair; double* col_wit;
  int* epa_list; int* epa_count;
  int* nar_k; float* nar_lb;   // split narrow phase: hand-over from the closed-form kernel to the GJK kernel
};
#ifndef DRC_PRIO_SLOTS   // lab builds (python -m dyros_robot_controller_b200.build --dev -DDRC_PRIO_SLOTS=...) measure other settings
#define DRC_PRIO_SLOTS 2048
#endif
#ifndef DRC_PRIO_ITERS
#define DRC_PRIO_ITERS 300
#endif
constexpr int kPrioSlots = DRC_PRIO_SLOTS;   // capacity of the priority sub-batch
constexpr int kPrioMinBatch = 8192;          // batches below this run as one pipeline
constexpr int kPrioIters = DRC_PRIO_ITERS;   // previous-tick iteration count from which a robot is "slow"

struct drc_ctx {
  const drc_model* model;
  DrcModelDev mdev;   // the model as the kernels of THIS context get it: a snapshot taken at creation, mesh hull vertices in device memory
  double* hull_dev;
  int device, cap;
  DrcParams prm;
  cudaStream_t stream;
  cudaStream_t side;      // EPA pass of the self-collision stage runs here, next to the state / QP-build kernel
  cudaEvent_t ev_col, ev_epa, ev_build;
  cudaStream_t prio_side; cudaEvent_t ev_prio_fk, ev_prio_build;
  cudaStream_t build_stream; cudaEvent_t ev_mbuild; bool par_build;   // main pipeline: QP build next to the narrow phase   // priority pipeline: QP build next to the narrow phase
  // state cache (SoA, stride cap)
  double *c_q, *c_qd, *c_oMi, *c_M, *c_Minv, *c_g, *c_nle;
  double *c_Mact, *c_Minvact, *c_gact, *c_nleact;  // mobile manipulator only (actuated-space dynamics)
  // QP + collision scratch
  double* qp;
  int qp_stride_max;
  int* epa_flag; unsigned long long* cand_mask; double* col_dist; int* col_pair; double* col_wit;
  int* epa_list; int* epa_count;
  int* nar_k; float* nar_lb;
  int *pinv_list, *pinv_count;           // whole-body models: robots whose full-model PinvCOD(M) goes to k_pinv_list
  int *manip_list, *manip_count;         // two-route manipulability of the main QPIK build: robots left to the exact route
  int *prev_iters, *order, *sched_hist;  // ADMM schedule: previous tick's iteration counts -> robot order (k_sched_*)
  double* roll; int* roll_i;             // rollout scratch (cubic profile; next tick's schedule histogram / offsets / ticket), allocated on first use
  int* slow_count;                       // device: number of leading entries of `order` that run in the priority pipeline
  Scratch prio;                          // compact scratch of the priority pipeline (kPrioSlots robots)
  cudaStream_t prio_stream;              // high-priority stream of the priority pipeline
  cudaEvent_t ev_sched, ev_prio, ev_order;   // ev_order: the ADMM schedule (built on the priority stream) is ready
  cudaStream_t last_stream; bool last_stream_set; cudaEvent_t ev_last;  // cross-stream ordering of consecutive calls (pick)
  cudaEvent_t ev_in, ev_out;   // joins of a caller stream with the context's prioritised streams (fused QPIK cycles)
  cudaStream_t dyn_stream; cudaEvent_t ev_store, ev_dyn, ev_solve;  // fused QPIK cycles: dynamics-only kernel behind the ADMM launch
  cudaStream_t copy;                     // host entry points: inputs that only stage 2 reads are uploaded here, behind stage 1
  cudaEvent_t ev_late, ev_early; bool late_pending;
  int sm_count;
  // device staging for host entry points
  double* stage; size_t stage_doubles;
  int* stage_i; size_t stage_ints;
  long long launches;
  bool timing;
  cudaEvent_t ev[4];
  float last_ms[4];
  // stage trace of the last fused call (instrumentation only, drc_ctx_last_trace): named events on the streams of the pipelines
  cudaEvent_t tr_ev[32]; const char* tr_name[32]; int tr_n;
  // optional debug outputs of the QP solves (drc_ctx_enable_qp_debug): primal / dual vectors in structured order
  double *dbg_x, *dbg_y;
  // rollout in progress (pipeline variant): every solver launch of the tick integrates the state in place and keeps the tallies
  double *roll_q, *roll_qd; Strided sroll; double roll_dt; int *roll_fail, *roll_iters;
  // warm-started rollout in progress: previous tick's primal / dual solution per robot (structured order of SolveIO::qp_x / qp_y)
  double *ws_x, *ws_y; bool warm_on;
};
static inline void mark(drc_ctx* c, const char* name, cudaStream_t s) {
  if (!c->timing || c->tr_n >= 32) return;
  if (!c->tr_ev[c->tr_n] && cudaEventCreate(&c->tr_ev[c->tr_n]) != cudaSuccess) { cudaGetLastError(); return; }
  c->tr_name[c->tr_n] = name;
  cudaEventRecord(c->tr_ev[c->tr_n++], s);
}

static DrcFrame frame_of(const drc_model* m, int fid) {
  DrcFrame f;
  const HostFrame& hf = m->hm.frames[fid];
  f.parent = hf.parent;
  std::memcpy(f.R, hf.R, sizeof f.R);
  std::memcpy(f.p, hf.p, sizeof f.p);
  return f;
}
static Strided lay(int layout, int K, int B) { return layout == DRC_LAYOUT_SOA ? soa(B) : aos(K); }

static Scratch main_scratch(const drc_ctx* c) {
  Scratch sc;
  sc.c_q = c->c_q; sc.c_qd = c->c_qd; sc.c_oMi = c->c_oMi; sc.c_M = c->c_M; sc.c_Minv = c->c_Minv; sc.c_g = c->c_g; sc.c_nle = c->c_nle;
  sc.Bc = c->cap; sc.qp = c->qp; sc.epa_flag = c->epa_flag; sc.cand_mask = c->cand_mask; sc.col_dist = c->col_dist;
  sc.col_pair = c->col_pair; sc.col_wit = c->col_wit; sc.epa_list = c->epa_list; sc.epa_count = c->epa_count;
  sc.nar_k = c->nar_k; sc.nar_lb = c->nar_lb;
  return sc;
}
static void bind_scratch(const Scratch& sc, JobIO& io) {
  io.c_q = sc.c_q; io.c_qd = sc.c_qd; io.c_oMi = sc.c_oMi; io.c_M = sc.c_M; io.c_Minv = sc.c_Minv; io.c_g = sc.c_g;
  io.c_nle = sc.c_nle; io.Bc = sc.Bc; io.qp = sc.qp;
}
static void bind_cache(const drc_ctx* c, JobIO& io) {
  io.c_q = c->c_q; io.c_qd = c->c_qd; io.c_oMi = c->c_oMi; io.c_M = c->c_M; io.c_Minv = c->c_Minv; io.c_g = c->c_g;
  io.c_nle = c->c_nle; io.Bc = c->cap;
  io.c_Mact = c->c_Mact; io.c_Minvact = c->c_Minvact; io.c_gact = c->c_gact; io.c_nleact = c->c_nleact;
}

// dispatch on the compile-time robot shape; extend the list to add robots (serial chains of 7 / 6 revolute joints:
// FR3 class, UR5e class)
#ifdef DRC_DEV_FR3_ONLY  // kernel development builds (build.py --dev): half the compile time
#define DRC_DISPATCH_NV(nv, chain, CALL)                                         \
  if ((nv) == 7 && (chain)) { constexpr int NV = 7; constexpr bool CHAIN = true; CALL; } \
  else return fail(DRC_E_UNSUPPORTED, "development build: 7-dof serial chains only");
#else
#define DRC_DISPATCH_NV(nv, chain, CALL)                                         \
  if ((nv) == 7 && (chain)) { constexpr int NV = 7; constexpr bool CHAIN = true; CALL; } \
  else if ((nv) == 6 && (chain)) { constexpr int NV = 6; constexpr bool CHAIN = true; CALL; } \
  else return fail(DRC_E_UNSUPPORTED, "no kernel instantiation for this robot (dof / topology)");
#endif

template <int NV, bool CHAIN, unsigned FLAGS, int W = 0>
static int launch_job(drc_ctx* c, const DrcFrame& fr, const JobIO& io, cudaStream_t s) {
  constexpr int threads = 64;
  const int blocks = io.redo ? 2 * c->sm_count : (io.B + threads - 1) / threads;   // redo: grid-stride over JobIO::manip_list
  k_robot_job<NV, CHAIN, FLAGS, W><<<blocks, threads, 0, s>>>(c->mdev, c->prm, fr, io);
  c->launches++;
  CU(cudaGetLastError());
  return DRC_OK;
}

// A job that refreshes the full-model dynamics of a whole-body model: the robots whose mass matrix fails the Cholesky guard (all of
// them for these models) are collected in a list and k_pinv_list takes PinvCOD(M) with 16 lanes per robot, behind the job.
template <int NV, unsigned FLAGS, int W>
static int launch_dyn_job(drc_ctx* c, const DrcFrame& fr, JobIO io, cudaStream_t s) {
  static_assert((FLAGS & F_DYN) && !(FLAGS & F_DYN_LIGHT), "a job that computes M^-1");
  io.pinv_list = c->pinv_list; io.pinv_count = c->pinv_count;
  CU(cudaMemsetAsync(c->pinv_count, 0, sizeof(int), s));
  int rc = launch_job<NV, false, FLAGS, W>(c, fr, io, s);
  if (rc) return rc;
  k_pinv_list<NV><<<4 * c->sm_count, 128, 0, s>>>(c->c_M, c->c_Minv, c->cap, c->pinv_list, c->pinv_count, c->prm.pinv_threshold);
  c->launches++;
  CU(cudaGetLastError());
  return DRC_OK;
}

#ifndef DRC_GJK_MINB
#define DRC_GJK_MINB 2
#endif
template <int NV, bool CHAIN>
static int launch_collision(drc_ctx* c, CollisionIO io, cudaStream_t s, bool epa_on_side_stream = false, const Scratch* scp = nullptr) {
  const Scratch sc = scp ? *scp : main_scratch(c);
  io.c_q = sc.c_q; io.c_qd = sc.c_qd; io.c_oMi = sc.c_oMi; io.Bc = sc.Bc;
  io.epa_flag = sc.epa_flag; io.cand_mask = sc.cand_mask; io.epa_list = sc.epa_list; io.epa_count = sc.epa_count;
  CU(cudaMemsetAsync(sc.epa_count, 0, sizeof(int), s));
  if (!io.dist) io.dist = sc.col_dist;
  if (!io.pair_out) io.pair_out = sc.col_pair;
  if (!io.witness) io.witness = sc.col_wit;
  constexpr int threads = 128;
  const int blocks = (io.B + threads - 1) / threads;
  // (register budgets of 168 / 128 registers were measured too: 3 blocks/SM is slower, 4 blocks/SM saves 0.1 ms here and loses it
  // again in the ADMM stage -- profiles/README.md)
  io.nar_k = sc.nar_k; io.nar_lb = sc.nar_lb;
  k_collision_closed<NV, CHAIN><<<blocks, threads, 0, s>>>(c->mdev, io);
  CU(cudaGetLastError());
  k_collision<NV, CHAIN, true, DRC_GJK_MINB><<<(io.B + kGjkThreads - 1) / kGjkThreads, kGjkThreads, 0, s>>>(c->mdev, c->prm, io);
  CU(cudaGetLastError());
  c->launches++;
  // the EPA pass touches ~0.1 % of the robots with one warp each: a long, nearly empty kernel.  The QP entry points run it
  // on the side stream, concurrently with the state / QP-build kernel (disjoint parts of the QP record), and join
  // before the ADMM launch (join_epa).
  cudaStream_t es = s;
  if (epa_on_side_stream) {
    CU(cudaEventRecord(c->ev_col, s));
    CU(cudaStreamWaitEvent(c->side, c->ev_col, 0));
    es = c->side;
  }
  const int epa_blocks = scp ? 16 : c->sm_count;
  k_collision_epa<NV, CHAIN><<<epa_blocks, kEpaWarps * 32, 0, es>>>(c->mdev, c->prm, io);
  CU(cudaGetLastError());
  if (epa_on_side_stream) CU(cudaEventRecord(c->ev_epa, c->side));
  c->launches += 2;
  return DRC_OK;
}

static int join_epa(drc_ctx* c, cudaStream_t s) {
  CU(cudaStreamWaitEvent(s, c->ev_epa, 0));
  return DRC_OK;
}
// The EPA pass (side stream) resolves the ~0.1 % of the robots with an overlapping cylinder / box pair; it is a long, nearly empty
// kernel.  Instead of holding the whole ADMM launch back until it ends, the main launch skips those robots (SolveIO::skip =
// epa_flag) and a second, small ADMM launch BEHIND the EPA pass on the side stream solves them.  `sio` is the main launch's
// argument block; the caller's stream waits for ev_epa (recorded here after the small launch) at the end of the call.
template <class Cfg, bool ID>
static int launch_admm(drc_ctx* c, SolveIO io, cudaStream_t s, unsigned unit_mask = (1u << Cfg::NC) - 1u, const double* gravity = nullptr,
                       const Scratch* scp = nullptr);
template <class Cfg, bool ID>
static int solve_epa_robots(drc_ctx* c, SolveIO sio, cudaStream_t main, unsigned unit_mask, const double* gravity) {
  CU(cudaEventRecord(c->ev_build, main));                  // the rest of their QP records (state / QP-build kernel)
  CU(cudaStreamWaitEvent(c->side, c->ev_build, 0));
  sio.order = c->epa_list; sio.count = c->epa_count; sio.order_off = nullptr; sio.skip = nullptr; sio.out_ids = nullptr;
  int rc = launch_admm<Cfg, ID>(c, sio, c->side, unit_mask, gravity, nullptr);
  if (rc) return rc;
  CU(cudaEventRecord(c->ev_epa, c->side));
  return DRC_OK;
}

// ADMM schedule from the previous tick's iteration counts (k_sched_* in drc_kernels.cuh): c->order = robots by descending
// count, *c->slow_count = how many of them reached kPrioIters (at most kPrioSlots).  Results do not depend on it.
static int launch_schedule(drc_ctx* c, int B, cudaStream_t s) {
  CU(cudaMemsetAsync(c->sched_hist, 0, kSchedBuckets * sizeof(int), s));
  CU(cudaMemsetAsync(c->slow_count, 0, sizeof(int), s));
  const int tb = 256, nb = (B + tb - 1) / tb;
  k_sched_hist<<<nb < 1024 ? nb : 1024, tb, 0, s>>>(c->prev_iters, B, c->sched_hist);
  k_sched_scan<<<1, kSchedBuckets, 0, s>>>(c->sched_hist, kPrioIters, kPrioSlots, c->slow_count);
  k_sched_scatter<<<nb, tb, 0, s>>>(c->prev_iters, B, c->sched_hist, c->order);
  c->launches += 3;
  CU(cudaGetLastError());
  return DRC_OK;
}

// io.B = slots to cover; io.order / order_off / out_ids / count select them (see SolveIO)
template <class Cfg, bool ID>
static int launch_admm(drc_ctx* c, SolveIO io, cudaStream_t s, unsigned unit_mask, const double* gravity, const Scratch* scp) {
  io.qp = scp ? scp->qp : c->qp;
  io.c_g = gravity ? gravity : (scp ? scp->c_g : c->c_g);
  io.Bc = scp ? scp->Bc : c->cap;
  const QpOptions o = qp_options(c->prm, unit_mask);
  const int per_block = kAdmmWarps * Cfg::NG, blocks = (io.B + per_block - 1) / per_block;
  io.iters_hint = c->prev_iters;
  io.qp_x = c->dbg_x; io.qp_y = c->dbg_y;
  if (c->warm_on && !ID) { io.qp_x = c->ws_x; io.qp_y = c->ws_y; io.warm_x = c->ws_x; io.warm_y = c->ws_y; }
  if (c->roll_q && !io.roll_q) { io.roll_q = c->roll_q; io.roll_qd = c->roll_qd; io.sroll = c->sroll; io.roll_dt = c->roll_dt; io.fail_ticks = c->roll_fail; io.iters_total = c->roll_iters; }
  // one instantiation per QP shape.  Register budget: 12 warps / SM = 168 registers for the whole-body QPIK shapes; 8 warps / SM = 255 registers for the QPID shapes -- with 4 unit bundles per core lane they spill 1.4 KB under the
  // 168-register cap and the spill traffic + a code size beyond the instruction cache made `long_scoreboard` and `no_instruction`
  // the top stalls (ncu, profiles/r02c_ncu_summary_siblings.md): same-box A/B 7.97e6 -> 1.37e7 FR3 QPID cycles/s
  // (profiles/r02c_lab_variants_qpid_registers.txt).  The dynamic shared-memory opt-in is a per-device
  // function attribute: set it for the context's device on every launch (cheap) so that contexts on several GPUs of one
  // process all get it.
  constexpr size_t smem = sizeof(GroupShared<Cfg>) * kAdmmWarps * Cfg::NG;
#ifndef DRC_ADMM_MINB_ID
#define DRC_ADMM_MINB_ID 8   // blocks (= warps) per SM of the QPID solver launches
#endif
#ifndef DRC_ADMM_MINB_IK
#define DRC_ADMM_MINB_IK 8   // manipulator QPIK: ptxas settles at 202 registers without a spill -> 9 warps / SM
#endif
  // manipulator QPIK (slack formulation): its step ends with the slowest robots' serial iterations, and the spill-free build
  // shortens them -- same-box A/B 2.29 -> 2.39e7 cycles/s with the schedule hint (1.60 -> 1.53e7 without it, where the bulk's
  // occupancy counts; profiles/r02c_lab_variants_qpik_registers.txt).  The whole-body QPIK shapes run at 262 144 ... 1 M robots,
  // where throughput counts: they keep 12 warps / SM.
  constexpr int MB = ID ? DRC_ADMM_MINB_ID : (Cfg::SLACK ? DRC_ADMM_MINB_IK : 3);
  CU(cudaFuncSetAttribute(k_admm<Cfg, ID, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_admm<Cfg, ID, MB><<<blocks, kAdmmWarps * 32, smem, s>>>(io, o);
  c->launches++;
  CU(cudaGetLastError());
  return DRC_OK;
}

// whole-body (tree-shaped) models behind the generic getters; defined in drc_moma.cu next to their kernel instantiations
int moma_get_frame_full(drc_ctx* c, int B, int frame, double* pose12, double* J, double* Jdot, double* vel, int layout, cudaStream_t s);
int moma_get_min_distance(drc_ctx* c, int B, int with_graddot, double* dist, double* grad, double* grad_dot, int* pair, int layout, cudaStream_t s);

static int check_batch(const drc_ctx* c, int B) {
  if (!c) return fail(DRC_E_INVALID, "null context");
  if (B <= 0 || B > c->cap) return fail(DRC_E_INVALID, "batch size out of range for this context");
  return DRC_OK;
}
static int check_frame(const drc_ctx* c, int frame) {
  if (frame < 0 || frame >= (int)c->model->hm.frames.size()) return fail(DRC_E_INVALID, "unknown frame id");
  return DRC_OK;
}
// Stream of this call.  The state cache and the scratch belong to the context, not to a stream: when consecutive calls use
// different streams (torch inputs run on the caller's stream, the drc_host_* getters on the context's non-blocking stream),
// the new stream first waits for everything the previous call enqueued -- an event recorded at the tail of the previous
// stream (calls are host-serialised per context, so that tail covers the whole previous call).
static cudaStream_t pick(drc_ctx* c, void* s) {
  cudaStream_t st = s ? (cudaStream_t)s : c->stream;
  if (c->last_stream_set && c->last_stream != st) {
    if (cudaEventRecord(c->ev_last, c->last_stream) == cudaSuccess) cudaStreamWaitEvent(st, c->ev_last, 0);
    else cudaGetLastError();  // the previous caller stream no longer exists: nothing of it can still be running
  }
  c->last_stream = st; c->last_stream_set = true;
  return st;
}


// ------------------------------------------------------------------------------------------------ host entry points
// Carve device staging buffers, copy inputs H2D on the context stream, run the device entry point, copy back.
struct Stage {
  drc_ctx* c;
  size_t used = 0, used_i = 0;
  struct Out { void* host; void* dev; size_t bytes; };
  std::vector<Out> outs;
  int err = DRC_OK;
  explicit Stage(drc_ctx* ctx) : c(ctx) {}
  double* in(const double* h, size_t cnt) {
    if (!h) return nullptr;
    double* d = take(cnt);
    if (d && cudaMemcpyAsync(d, h, cnt * sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) err = DRC_E_CUDA;
    return d;
  }
  // upload on the copy stream: overlaps the kernels that do not read it (run_qp waits for ev_late before stage 2)
  bool late_started = false;
  double* in_late(const double* h, size_t cnt) {
    if (!h) return nullptr;
    if (!late_started) {  // behind the inputs already queued on the main stream (they are needed first; one PCIe link)
      late_started = true;
      if (cudaEventRecord(c->ev_early, c->stream) != cudaSuccess || cudaStreamWaitEvent(c->copy, c->ev_early, 0) != cudaSuccess) err = DRC_E_CUDA;
    }
    double* d = take(cnt);
    if (d && cudaMemcpyAsync(d, h, cnt * sizeof(double), cudaMemcpyHostToDevice, c->copy) != cudaSuccess) err = DRC_E_CUDA;
    return d;
  }
  void late_done() {
    if (cudaEventRecord(c->ev_late, c->copy) != cudaSuccess) err = DRC_E_CUDA;
    c->late_pending = true;
  }
  double* out(double* h, size_t cnt) {
    if (!h) return nullptr;
    double* d = take(cnt);
    if (d) outs.push_back({h, d, cnt * sizeof(double)});
    return d;
  }
  int* out_i(int* h, size_t cnt) {
    if (!h) return nullptr;
    if (used_i + cnt > c->stage_ints) { err = DRC_E_NOMEM; return nullptr; }
    int* d = c->stage_i + used_i;
    used_i += cnt;
    outs.push_back({h, d, cnt * sizeof(int)});
    return d;
  }
  // The solver launches write every result once, from the lane that owns it, and nothing on the device reads it back.  When the
  // caller's buffer is page-locked host memory this device can address (cudaHostAlloc / cudaHostRegister under unified addressing),
  // the fused cycle entry points hand the kernels its device alias: the results cross PCIe while the solve is still running instead
  // of in a staged copy behind the last kernel.  Pageable buffers take the staged route.
  template <class T>
  static T* device_alias(T* h) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    return static_cast<T*>(at.devicePointer);
  }
  double* out_direct(double* h, size_t cnt) {
    if (!h) return nullptr;
    if (double* d = device_alias(h)) return d;
    return out(h, cnt);
  }
  int* out_i_direct(int* h, size_t cnt) {
    if (!h) return nullptr;
    if (int* d = device_alias(h)) return d;
    return out_i(h, cnt);
  }
  double* take(size_t cnt) {
    if (used + cnt > c->stage_doubles) { err = DRC_E_NOMEM; return nullptr; }
    double* d = c->stage + used;
    used += cnt;
    return d;
  }
  int finish(int rc) {
    if (err) return fail(err, "host staging failed (buffer too small or copy error)");
    if (rc) return rc;
    for (auto& o : outs)
      if (cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return fail(DRC_E_CUDA, "D2H copy failed");
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return fail(DRC_E_CUDA, std::string("kernel execution failed: ") + cudaGetErrorString(e));
    return DRC_OK;
  }
};
#define HOST_PRELUDE                                   \
  int rc0 = check_batch(c, B); if (rc0) return rc0;    \
  CU(cudaSetDevice(c->device));                        \
  const int n = c->model->hm.dev.nv; (void)n;          \
  const size_t Bz = (size_t)B; (void)Bz;               \
  Stage st(c);

