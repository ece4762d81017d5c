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
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/drc_b200.h"
#include "drc_cycle.h"
#include "model.h"

using namespace drc;

// ================================================================================================ kernels
template <int NV, bool CHAIN, unsigned FLAGS>
__global__ void __launch_bounds__(128) k_robot_job(const __grid_constant__ DrcModelDev m, const __grid_constant__ DrcParams prm,
                                                    const __grid_constant__ DrcFrame frame, const __grid_constant__ JobIO io) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < io.B) robot_job<NV, CHAIN, FLAGS>(m, prm, frame, io, b);
}

template <int NV, bool CHAIN>
__global__ void __launch_bounds__(128) k_collision(const __grid_constant__ DrcModelDev m, const __grid_constant__ DrcParams prm,
                                                    const __grid_constant__ CollisionIO io) {
  // stage the geometry table in shared memory: the GJK pass indexes it with per-thread pair ids
  __shared__ GeomTable G;
  {
    const int* src = reinterpret_cast<const int*>(&m.geom);
    int* dst = reinterpret_cast<int*>(&G);
    for (int i = threadIdx.x; i < (int)(sizeof(GeomTable) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < io.B) collision_job<NV, CHAIN>(m, G, prm, io, b);
}
// ---- EPA, one WARP per flagged robot (~0.1 % of a random batch).  Same algorithm and rules as the scalar
// epa_penetration (drc_geom.h: flood-fill horizon, uncommitted bad expansions, slot policy); the polytope lives in
// shared memory and the three face scans (closest face, neighbour across an edge, new faces) are spread over the
// lanes.  Control flow and the small flood-fill state are warp-uniform: every lane runs the same sequence.
struct EpaWarpSmem {
  SimplexVert P[kEpaMaxVert];
  double fd[kEpaMaxFace];
  double fn[kEpaMaxFace][3];
  short fv[kEpaMaxFace][3];
  unsigned char mark[kEpaMaxFace];
  short edge[kEpaMaxEdge][2];
  short killed[kEpaMaxEdge];
  short stack[kEpaMaxEdge];
};
__device__ __forceinline__ void epa_make_face(EpaWarpSmem& S, int slot, int a, int b, int c) {
  const Vec3 nrm = cross(S.P[b].w - S.P[a].w, S.P[c].w - S.P[a].w);
  const double l = norm(nrm);
  const Vec3 n = l > 0 ? (1.0 / l) * nrm : v3(0, 0, 1);
  S.fv[slot][0] = (short)a; S.fv[slot][1] = (short)b; S.fv[slot][2] = (short)c;
  S.fn[slot][0] = n.x; S.fn[slot][1] = n.y; S.fn[slot][2] = n.z;
  S.fd[slot] = dot(n, S.P[a].w);
}
__device__ void epa_warp(EpaWarpSmem& S, const Prim& A, const Prim& B, const GjkOut& g, double tol, int max_iter, PairResult& out,
                         int lane) {
  const unsigned full = 0xffffffffu;
  auto sup = [&](Vec3 d) { SimplexVert s; s.a = support(A, d); s.b = support(B, -d); s.w = s.a - s.b; return s; };
  SimplexVert T[4];
  int np = epa_seed(sup, g, T);  // warp-uniform, every lane on its private copy
  out.d = 0; out.pa = g.pa; out.pb = g.pb;
  if (np < 4) return;
  __syncwarp();
  if (lane < 4) S.P[lane] = T[lane];
  __syncwarp();
  if (lane == 0) epa_make_face(S, 0, 0, 1, 2);
  if (lane == 1) epa_make_face(S, 1, 0, 3, 1);
  if (lane == 2) epa_make_face(S, 2, 0, 2, 3);
  if (lane == 3) epa_make_face(S, 3, 1, 3, 2);
  int nf = 4;
  __syncwarp();
  int bestf = -1;
  for (int it = 0; it < max_iter; ++it) {
    // closest face to the origin (ties: smallest index); every slot below nf holds a live face
    double bd = 1e300;
    int bf = -1;
    for (int f = lane; f < nf; f += 32)
      if (S.fd[f] < bd) { bd = S.fd[f]; bf = f; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double od = __shfl_xor_sync(full, bd, off);
      const int of = __shfl_xor_sync(full, bf, off);
      if (of >= 0 && (bf < 0 || od < bd || (od == bd && of < bf))) { bd = od; bf = of; }
    }
    bestf = bf;
    if (bf < 0) break;
    const Vec3 n = v3(S.fn[bf][0], S.fn[bf][1], S.fn[bf][2]);
    const SimplexVert s = sup(n);
    if (dot(n, s.w) - S.fd[bf] <= tol) break;
    if (np >= kEpaMaxVert) break;
    // flood fill of the faces visible from s (uniform control flow; lanes share the neighbour search)
    for (int f = lane; f < nf; f += 32) S.mark[f] = 0;
    __syncwarp();
    int nk = 0, ne = 0, sp = 0;
    bool bad = false;
    if (lane == 0) { S.mark[bf] = 1; S.killed[0] = (short)bf; S.stack[0] = (short)bf; }
    nk = 1; sp = 1;
    __syncwarp();
    while (sp > 0 && !bad) {
      const int f = S.stack[--sp];
      for (int e = 0; e < 3 && !bad; ++e) {
        const short a = S.fv[f][e], b = S.fv[f][(e + 1) % 3];
        int gn = -1;
        for (int base = 0; base < nf && gn < 0; base += 32) {
          const int i = base + lane;
          bool hit = false;
          if (i < nf) {
            const short v0 = S.fv[i][0], v1 = S.fv[i][1], v2 = S.fv[i][2];
            hit = (v0 == b && v1 == a) || (v1 == b && v2 == a) || (v2 == b && v0 == a);
          }
          const unsigned mk = __ballot_sync(full, hit);
          if (mk) gn = base + __ffs(mk) - 1;
        }
        if (gn < 0) { bad = true; break; }
        const int mg = S.mark[gn];
        if (mg == 1) continue;
        if (mg == 0) {
          const Vec3 p0 = S.P[S.fv[gn][0]].w;
          const bool vis = S.fn[gn][0] * (s.w.x - p0.x) + S.fn[gn][1] * (s.w.y - p0.y) + S.fn[gn][2] * (s.w.z - p0.z) > kEpaVisEps;
          __syncwarp();
          if (lane == 0) S.mark[gn] = vis ? 1 : 2;
          if (vis) {
            if (nk >= kEpaMaxEdge - 2) { bad = true; break; }
            if (lane == 0) { S.killed[nk] = (short)gn; S.stack[sp] = (short)gn; }
            ++nk; ++sp;
            __syncwarp();
            continue;
          }
          __syncwarp();
        }
        if (ne >= kEpaMaxEdge) { bad = true; break; }
        if (lane == 0) { S.edge[ne][0] = a; S.edge[ne][1] = b; }
        ++ne;
      }
    }
    __syncwarp();
    if (bad || ne != nk + 2 || nf + 2 > kEpaMaxFace) break;
    for (int base = 0; base < ne; base += 32) {
      const int k = base + lane;
      bool deg = false;
      if (k < ne) {
        const Vec3 pa = S.P[S.edge[k][0]].w;
        deg = norm2(cross(S.P[S.edge[k][1]].w - pa, s.w - pa)) <= kEpaMinArea2;
      }
      if (__any_sync(full, deg)) bad = true;
    }
    if (bad) break;
    // commit: the new vertex, then one new face per horizon edge
    const int idx = np++;
    if (lane == 0) S.P[idx] = s;
    __syncwarp();
    for (int k = lane; k < ne; k += 32) epa_make_face(S, k < nk ? (int)S.killed[k] : nf + (k - nk), S.edge[k][0], S.edge[k][1], idx);
    nf += 2;
    __syncwarp();
  }
  if (bestf < 0) return;
  epa_witness(S.P[S.fv[bestf][0]], S.P[S.fv[bestf][1]], S.P[S.fv[bestf][2]], v3(S.fn[bestf][0], S.fn[bestf][1], S.fn[bestf][2]),
              S.fd[bestf], out);
  __syncwarp();
}

constexpr int kEpaWarps = 2;
template <int NV, bool CHAIN>
__global__ void __launch_bounds__(kEpaWarps * 32) k_collision_epa(const __grid_constant__ DrcModelDev m, const __grid_constant__ DrcParams prm,
                                                                   const __grid_constant__ CollisionIO io) {
  __shared__ EpaWarpSmem sm[kEpaWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  EpaWarpSmem& S = sm[warp];
  const int count = *io.epa_count;
  for (int i = blockIdx.x * kEpaWarps + warp; i < count; i += gridDim.x * kEpaWarps) {
    const int b = io.epa_list[i];
    unsigned long long deferred = io.cand_mask[b];
    BestPair best;
    best.d = io.dist[b]; best.id = io.pair_out[b];
    best.pa = v3(io.witness[6 * b + 0], io.witness[6 * b + 1], io.witness[6 * b + 2]);
    best.pb = v3(io.witness[6 * b + 3], io.witness[6 * b + 4], io.witness[6 * b + 5]);
    best.ja = -1; best.jb = -1;
    for (int k = 0; k < m.npair; ++k)
      if (m.geom.pair_id[k] == best.id) { best.ja = m.geom.parent[m.geom.pair_a[k]]; best.jb = m.geom.parent[m.geom.pair_b[k]]; }
    while (deferred) {
      const int bit = __ffsll((long long)deferred) - 1;
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
      if (g.intersect) epa_warp(S, A, Bp, g, prm.epa_tol, prm.epa_max_iter, r, lane);
      consider(best, r.d, m.geom.pair_id[k], ja, jb, r.pa, r.pb);
    }
    __syncwarp();
    if (lane == 0) collision_finish<NV, CHAIN>(m, prm, io, b, best);
    __syncwarp();
  }
}

constexpr int kAdmmWarps = 1;  // one warp per block: a finished warp frees its slot without waiting for block-mates
template <class Cfg, bool ID, int MINB>
__global__ void __launch_bounds__(kAdmmWarps * 32, 4 * MINB) k_admm(const __grid_constant__ SolveIO io, const __grid_constant__ QpOptions o) {
  extern __shared__ __align__(16) unsigned char admm_smem[];  // dynamic: the QPID record exceeds the 48 KB static limit
  GroupShared<Cfg>* sh = reinterpret_cast<GroupShared<Cfg>*>(admm_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int first = (blockIdx.x * kAdmmWarps + warp) * Cfg::NG;
  if (first >= io.B) return;  // no block-level barrier below: idle warps may leave
  int robots[Cfg::NG];
#pragma unroll
  for (int g = 0; g < Cfg::NG; ++g) robots[g] = first + g < io.B ? first + g : -1;
  WarpExec<Cfg> w;
  w.sh = sh + warp * Cfg::NG;
  w.lane = lane;
  lane_assign<Cfg>(w.L, lane);
  solve_and_emit<Cfg, ID>(w, robots, io, o);
}

// DyrosMath::getTaskSpaceCubic (math_type_define.h:62-144,235-281,647-685); rotation log/exp by Rodrigues
struct CubicIO {
  int B;
  const double *x_target, *xdot_target, *x_init, *xdot_init;
  Strided s12, s6;
  double t, t0, dur;
  double *x_des, *xdot_des;
};
__device__ __forceinline__ double cubic_pos(double t, double t0, double tf, double x0, double xf, double v0, double vf) {
  if (t < t0) return x0;
  if (t > tf) return xf;
  const double e = t - t0, T = tf - t0, T2 = T * T, T3 = T2 * T, dx = xf - x0;
  return x0 + v0 * e + (3 * dx / T2 - 2 * v0 / T - vf / T) * e * e + (-2 * dx / T3 + (v0 + vf) / T2) * e * e * e;
}
__device__ __forceinline__ double cubic_vel(double t, double t0, double tf, double x0, double xf, double v0, double vf) {
  if (t < t0) return v0;
  if (t > tf) return vf;
  const double e = t - t0, T = tf - t0, T2 = T * T, T3 = T2 * T, dx = xf - x0;
  return v0 + 2 * (3 * dx / T2 - 2 * v0 / T - vf / T) * e + 3 * (-2 * dx / T3 + (v0 + vf) / T2) * e * e;
}
__device__ Vec3 so3_log(const Mat3& R) {
  const double tr = R.m[0] + R.m[4] + R.m[8];
  const double c = dmin(dmax(0.5 * (tr - 1.0), -1.0), 1.0);
  const double th = acos(c);
  const Vec3 w = v3(R.m[7] - R.m[5], R.m[2] - R.m[6], R.m[3] - R.m[1]);
  if (th < 1e-8) return 0.5 * w;
  if (3.14159265358979323846 - th < 1e-6) {
    int k = 0;
    if (R.m[4] > R.m[0]) k = 1;
    if (R.m[8] > R.m[4 * k]) k = 2;
    Vec3 cl = v3(R.m[k], R.m[3 + k], R.m[6 + k]);
    if (k == 0) cl.x += 1.0; else if (k == 1) cl.y += 1.0; else cl.z += 1.0;
    Vec3 ax = (1.0 / norm(cl)) * cl;
    if (dot(ax, w) < 0) ax = -ax;
    return th * ax;
  }
  return (th / (2.0 * sin(th))) * w;
}
__device__ Mat3 so3_exp(Vec3 w) {
  const double th = norm(w);
  if (th < 1e-12) { Mat3 R = {{1, -w.z, w.y, w.z, 1, -w.x, -w.y, w.x, 1}}; return R; }
  double s, c;
  sincos(th, &s, &c);
  return rot_axis((1.0 / th) * w, s, c);
}
__global__ void k_task_cubic(const __grid_constant__ CubicIO io) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= io.B) return;
  auto ld12 = [&](const double* p, int k) { return p[b * io.s12.sb + k * io.s12.sk]; };
  auto ld6 = [&](const double* p, int k) { return p[b * io.s6.sb + k * io.s6.sk]; };
  const double tf = io.t0 + io.dur;
  Mat3 R0, Rf;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) { R0.m[3 * r + c] = ld12(io.x_init, 4 * r + c); Rf.m[3 * r + c] = ld12(io.x_target, 4 * r + c); }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double p0 = ld12(io.x_init, 4 * i + 3), pf = ld12(io.x_target, 4 * i + 3), v0 = ld6(io.xdot_init, i), vf = ld6(io.xdot_target, i);
    io.x_des[b * io.s12.sb + (4 * i + 3) * io.s12.sk] = cubic_pos(io.t, io.t0, tf, p0, pf, v0, vf);
    io.xdot_des[b * io.s6.sb + i * io.s6.sk] = cubic_vel(io.t, io.t0, tf, p0, pf, v0, vf);
  }
  const Vec3 r = so3_log(tmul(R0, Rf));
  Mat3 Rd;
  if (io.t >= tf) Rd = Rf;
  else if (io.t < io.t0) Rd = R0;
  else Rd = mul(R0, so3_exp(cubic_pos(io.t, io.t0, tf, 0, 1, 0, 0) * r));
#pragma unroll
  for (int rr = 0; rr < 3; ++rr)
#pragma unroll
    for (int c = 0; c < 3; ++c) io.x_des[b * io.s12.sb + (4 * rr + c) * io.s12.sk] = Rd.m[3 * rr + c];
  Vec3 rd = v3(cubic_vel(io.t, io.t0, tf, 0, r.x, 0, 0), cubic_vel(io.t, io.t0, tf, 0, r.y, 0, 0), cubic_vel(io.t, io.t0, tf, 0, r.z, 0, 0));
  rd = mul(R0, rd);
  const double tau = (io.t - io.t0) / (tf - io.t0);
  if (tau < 0 || tau > 1) rd = v3(0, 0, 0);
  io.xdot_des[b * io.s6.sb + 3 * io.s6.sk] = rd.x;
  io.xdot_des[b * io.s6.sb + 4 * io.s6.sk] = rd.y;
  io.xdot_des[b * io.s6.sb + 5 * io.s6.sk] = rd.z;
}

// cache (SoA [K][Bc]) -> user array; `sub` != null writes src - sub (coriolis = nle - g)
__global__ void k_copy_cache(const double* src, const double* sub, long long Bc, int K, int B, double* dst, Strided s) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)K * B) return;
  const int k = (int)(i / B), b = (int)(i % B);
  double v = src[(long long)k * Bc + b];
  if (sub) v -= sub[(long long)k * Bc + b];
  dst[b * s.sb + k * s.sk] = v;
}

// FP64 FMA peak: 8 independent chains per thread
__global__ void k_fp64_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double x = 1.0000001, y = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
    a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ================================================================================================ host side
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CU(x)                                                                                      \
  do {                                                                                             \
    cudaError_t e_ = (x);                                                                          \
    if (e_ != cudaSuccess) return fail(DRC_E_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); \
  } while (0)

struct drc_model {
  HostModel hm;
  std::string verbose;
};

struct drc_ctx {
  const drc_model* model;
  int device, cap;
  DrcParams prm;
  cudaStream_t stream;
  // state cache (SoA, stride cap)
  double *c_q, *c_qd, *c_oMi, *c_M, *c_Minv, *c_g, *c_nle;
  // QP + collision scratch
  double* qp;
  int qp_stride_max;
  int* epa_flag; unsigned long long* cand_mask; double* col_dist; int* col_pair; double* col_wit;
  int* epa_list; int* epa_count;
  int sm_count;
  // device staging for host entry points
  double* stage; size_t stage_doubles;
  int* stage_i; size_t stage_ints;
  long long launches;
  bool timing;
  cudaEvent_t ev[4];
  float last_ms[4];
};

static DrcFrame frame_of(const drc_model* m, int fid) {
  DrcFrame f;
  const HostFrame& hf = m->hm.frames[fid];
  f.parent = hf.parent;
  std::memcpy(f.R, hf.R, sizeof f.R);
  std::memcpy(f.p, hf.p, sizeof f.p);
  return f;
}
static Strided lay(int layout, int K, int B) { return layout == DRC_LAYOUT_SOA ? soa(B) : aos(K); }

static void bind_cache(const drc_ctx* c, JobIO& io) {
  io.c_q = c->c_q; io.c_qd = c->c_qd; io.c_oMi = c->c_oMi; io.c_M = c->c_M; io.c_Minv = c->c_Minv; io.c_g = c->c_g;
  io.c_nle = c->c_nle; io.Bc = c->cap;
}

// dispatch on the compile-time robot shape; extend the list to add robots
#define DRC_DISPATCH_NV(nv, chain, CALL)                                         \
  if ((nv) == 7 && (chain)) { constexpr int NV = 7; constexpr bool CHAIN = true; CALL; } \
  else return fail(DRC_E_UNSUPPORTED, "no kernel instantiation for this robot (dof / topology)");

template <int NV, bool CHAIN, unsigned FLAGS>
static int launch_job(drc_ctx* c, const DrcFrame& fr, const JobIO& io, cudaStream_t s) {
  static const int threads = [] { const char* e = getenv("DRC_JOB_THREADS"); const int t = e ? atoi(e) : 64; return t >= 32 && t <= 128 ? t : 64; }();
  const int blocks = (io.B + threads - 1) / threads;
  k_robot_job<NV, CHAIN, FLAGS><<<blocks, threads, 0, s>>>(c->model->hm.dev, c->prm, fr, io);
  c->launches++;
  CU(cudaGetLastError());
  return DRC_OK;
}

template <int NV, bool CHAIN>
static int launch_collision(drc_ctx* c, CollisionIO io, cudaStream_t s) {
  io.c_q = c->c_q; io.c_qd = c->c_qd; io.c_oMi = c->c_oMi; io.Bc = c->cap;
  io.epa_flag = c->epa_flag; io.cand_mask = c->cand_mask; io.epa_list = c->epa_list; io.epa_count = c->epa_count;
  CU(cudaMemsetAsync(c->epa_count, 0, sizeof(int), s));
  if (!io.dist) io.dist = c->col_dist;
  if (!io.pair_out) io.pair_out = c->col_pair;
  if (!io.witness) io.witness = c->col_wit;
  static const int threads = [] { const char* e = getenv("DRC_COL_THREADS"); const int t = e ? atoi(e) : 64; return t >= 32 && t <= 128 ? t : 64; }();
  const int blocks = (io.B + threads - 1) / threads;
  k_collision<NV, CHAIN><<<blocks, threads, 0, s>>>(c->model->hm.dev, c->prm, io);
  CU(cudaGetLastError());
  k_collision_epa<NV, CHAIN><<<c->sm_count, kEpaWarps * 32, 0, s>>>(c->model->hm.dev, c->prm, io);
  CU(cudaGetLastError());
  c->launches += 2;
  return DRC_OK;
}

template <class Cfg, bool ID>
static int launch_admm(drc_ctx* c, SolveIO io, cudaStream_t s) {
  io.qp = c->qp; io.c_g = c->c_g; io.Bc = c->cap;
  const QpOptions o = qp_options(c->prm, (1u << Cfg::NC) - 1u);
  const int per_block = kAdmmWarps * Cfg::NG, blocks = (io.B + per_block - 1) / per_block;
  // blocks/SM the kernel is compiled for (register cap 65536 / (128 * MINB)); tunable for experiments
  static const int minb = [] { const char* e = getenv("DRC_ADMM_MINB"); return e ? atoi(e) : 3; }();
  constexpr size_t smem = sizeof(GroupShared<Cfg>) * kAdmmWarps * Cfg::NG;
  static const cudaError_t attr = [] {
    cudaError_t e1 = cudaFuncSetAttribute(k_admm<Cfg, ID, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaError_t e2 = cudaFuncSetAttribute(k_admm<Cfg, ID, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaError_t e3 = cudaFuncSetAttribute(k_admm<Cfg, ID, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
  }();
  CU(attr);
  if (minb <= 2) k_admm<Cfg, ID, 2><<<blocks, kAdmmWarps * 32, smem, s>>>(io, o);
  else if (minb == 3) k_admm<Cfg, ID, 3><<<blocks, kAdmmWarps * 32, smem, s>>>(io, o);
  else k_admm<Cfg, ID, 4><<<blocks, kAdmmWarps * 32, smem, s>>>(io, o);
  c->launches++;
  CU(cudaGetLastError());
  return DRC_OK;
}

static int check_batch(const drc_ctx* c, int B) {
  if (!c) return fail(DRC_E_INVALID, "null context");
  if (B <= 0 || B > c->cap) return fail(DRC_E_INVALID, "batch size out of range for this context");
  return DRC_OK;
}
static int check_frame(const drc_ctx* c, int frame) {
  if (frame < 0 || frame >= (int)c->model->hm.frames.size()) return fail(DRC_E_INVALID, "unknown frame id");
  return DRC_OK;
}
static cudaStream_t pick(drc_ctx* c, void* s) { return s ? (cudaStream_t)s : c->stream; }

// one QP controller call on the cached state (or fused with the state update when q != null)
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
  if (c->timing) cudaEventRecord(c->ev[0], s);
  int rc;
  const bool fused = q != nullptr;
#define J(FL) launch_job<NV, CHAIN, FL>(c, fr, io, s)
  if (!id) {
    if (fused) rc = step ? J(F_DYN | F_STORE | F_QPIK | F_STEP) : J(F_DYN | F_STORE | F_QPIK);
    else rc = step ? J(F_FROM_CACHE | F_QPIK | F_STEP) : J(F_FROM_CACHE | F_QPIK);
  } else {
    if (fused) rc = step ? J(F_DYN | F_STORE | F_QPID | F_STEP) : J(F_DYN | F_STORE | F_QPID);
    else rc = step ? J(F_FROM_CACHE | F_QPID | F_STEP) : J(F_FROM_CACHE | F_QPID);
  }
#undef J
  if (rc) return rc;
  if (c->timing) cudaEventRecord(c->ev[1], s);
  CollisionIO cio;
  std::memset(&cio, 0, sizeof cio);
  cio.B = B; cio.mode = id ? 2 : 1; cio.qp = c->qp;
  cio.qp_stride = id ? QpidCfg<NV>::STRIDE : QpikCfg<NV>::STRIDE;
  cio.qp_row_off = (id ? QpidCfg<NV>::OFF_ROW : QpikCfg<NV>::OFF_ROW) + (NV + 1);
  rc = launch_collision<NV, CHAIN>(c, cio, s);
  if (rc) return rc;
  if (c->timing) cudaEventRecord(c->ev[2], s);
  SolveIO sio;
  std::memset(&sio, 0, sizeof sio);
  sio.B = B; sio.out = out; sio.sout = lay(layout, NV, B); sio.out2 = out2; sio.sout2 = sio.sout; sio.status = status; sio.iters = iters;
  rc = id ? launch_admm<QpidCfg<NV>, true>(c, sio, s) : launch_admm<QpikCfg<NV>, false>(c, sio, s);
  if (c->timing) cudaEventRecord(c->ev[3], s);
  return rc;
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
int drc_model_create_from_urdf(const char* urdf_path, const char* srdf_path, const char* packages_path, drc_model_t** out) {
  (void)packages_path;
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
  return drc_model_create_from_text(urdf.c_str(), srdf.c_str(), out);
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
const char* drc_model_verbose(const drc_model_t* m) { return m ? m->verbose.c_str() : ""; }

// ------------------------------------------------------------------------------------------------ context
int drc_ctx_create(const drc_model_t* m, int device, int max_batch, drc_ctx_t** out) {
  if (!m || !out || max_batch <= 0) return fail(DRC_E_INVALID, "bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(DRC_E_CUDA, "no CUDA device: drc_b200 has no CPU fallback");
  }
  if (device < 0 || device >= ndev) return fail(DRC_E_INVALID, "device index out of range");
  CU(cudaSetDevice(device));
  std::unique_ptr<drc_ctx> c(new drc_ctx);
  std::memset(c.get(), 0, sizeof(drc_ctx));
  c->model = m; c->device = device; c->cap = max_batch;
  c->prm = DrcParams();
  for (int i = 0; i < kMaxV; ++i) { c->prm.Kp_joint[i] = 400; c->prm.Kv_joint[i] = 40; }
  const int n = m->hm.dev.nv;
  const size_t B = (size_t)max_batch;
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  auto dalloc = [&](double** p, size_t cnt) { return cudaMalloc((void**)p, cnt * sizeof(double)); };
  CU(dalloc(&c->c_q, n * B)); CU(dalloc(&c->c_qd, n * B)); CU(dalloc(&c->c_oMi, 12 * n * B));
  CU(dalloc(&c->c_M, n * n * B)); CU(dalloc(&c->c_Minv, n * n * B)); CU(dalloc(&c->c_g, n * B)); CU(dalloc(&c->c_nle, n * B));
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
  CU(cudaMalloc((void**)&c->epa_list, B * sizeof(int)));
  CU(cudaMalloc((void**)&c->epa_count, sizeof(int)));
  {
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
  }
  // staging: enough for the largest host call (q, qd, pose, xdot in; J/Jdot/M outs)
  c->stage_doubles = B * (size_t)(2 * n + 12 + 6 + 2 * n + 12 * n + 2 * n * n + 64);
  CU(dalloc(&c->stage, c->stage_doubles));
  c->stage_ints = 2 * B;
  CU(cudaMalloc((void**)&c->stage_i, c->stage_ints * sizeof(int)));
  for (int i = 0; i < 4; ++i) CU(cudaEventCreate(&c->ev[i]));
  *out = c.release();
  return DRC_OK;
}
void drc_ctx_destroy(drc_ctx_t* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  double* ds[] = {c->c_q, c->c_qd, c->c_oMi, c->c_M, c->c_Minv, c->c_g, c->c_nle, c->qp, c->col_dist, c->col_wit, c->stage};
  for (double* p : ds) if (p) cudaFree(p);
  if (c->epa_flag) cudaFree(c->epa_flag);
  if (c->cand_mask) cudaFree(c->cand_mask);
  if (c->col_pair) cudaFree(c->col_pair);
  if (c->epa_list) cudaFree(c->epa_list);
  if (c->epa_count) cudaFree(c->epa_count);
  if (c->stage_i) cudaFree(c->stage_i);
  for (int i = 0; i < 4; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  cudaStreamDestroy(c->stream);
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
  p->pinv_threshold = s.pinv_threshold;
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
  s.pinv_threshold = p->pinv_threshold;
  return DRC_OK;
}
int drc_ctx_max_batch(const drc_ctx_t* c) { return c ? c->cap : DRC_E_INVALID; }
int drc_ctx_synchronize(drc_ctx_t* c) {
  if (!c) return fail(DRC_E_INVALID, "null context");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
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
  const double *dq = st.in(q, Bz * n), *dqd = st.in(qdot, Bz * n), *dxt = st.in(x_target, Bz * 12), *dx = st.in(xdot_target, Bz * 6);
  double* dout = st.out(qdot_out, Bz * n);
  int *ds = st.out_i(status, Bz), *di = st.out_i(iters, Bz);
  return st.finish(st.err ? st.err : drc_batch_cycle_qpik_step(c, B, dq, dqd, dxt, dx, frame, dout, ds, di, DRC_LAYOUT_AOS, nullptr));
}
int drc_host_cycle_qpid_step(drc_ctx_t* c, int B, const double* q, const double* qdot, const double* x_target, const double* xdot_target, int frame, double* tau_out, int* status, int* iters) {
  HOST_PRELUDE
  const double *dq = st.in(q, Bz * n), *dqd = st.in(qdot, Bz * n), *dxt = st.in(x_target, Bz * 12), *dx = st.in(xdot_target, Bz * 6);
  double* dout = st.out(tau_out, Bz * n);
  int *ds = st.out_i(status, Bz), *di = st.out_i(iters, Bz);
  return st.finish(st.err ? st.err : drc_batch_cycle_qpid_step(c, B, dq, dqd, dxt, dx, frame, dout, ds, di, DRC_LAYOUT_AOS, nullptr));
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
