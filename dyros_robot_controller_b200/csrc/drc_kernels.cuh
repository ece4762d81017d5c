// drc_b200 -- CUDA kernels (sm_100a).  Included by every translation unit of libdrc_b200.so that launches them
// (drc_lib.cu: manipulator path + C ABI core, drc_moma.cu: mobile-manipulator path); each unit instantiates the
// robot shapes it dispatches on.
//
// Kernels (one launch each, all fp64):
//   k_robot_job<NV,CHAIN,FLAGS,W> one robot per thread: state update, frame quantities, QP records
//   k_collision / k_collision_epa one robot per thread / one warp per flagged robot: min self-distance, gradients, QP row
//   k_admm<Cfg,ID,MINB>           GL lanes per robot (NG robots per warp): OSQP-algorithm ADMM
//   k_task_cubic                  one robot per thread: cubic task-space trajectory
//   k_copy_cache                  cache (SoA) -> user layout
// The robot model travels as a __grid_constant__ kernel parameter (constant bank, broadcast reads).
#pragma once
#include <cuda_runtime.h>

#include "drc_cycle.h"

namespace drc_kernels {
using namespace drc;

template <int NV, bool CHAIN, unsigned FLAGS, int W = 0>
__global__ void __launch_bounds__(128) k_robot_job(const __grid_constant__ DrcModelDev m, const __grid_constant__ DrcParams prm,
                                                    const __grid_constant__ DrcFrame frame, const __grid_constant__ JobIO io) {
#ifndef DRC_SYNTAX_CHECK   // -DDRC_SYNTAX_CHECK: empty kernel bodies, for a host-code syntax pass in seconds (build.py --check)
  if (io.redo) {   // follow-up launch over the robots the Cholesky route of the manipulability could not certify (JobIO::manip_list)
    const int n = *io.manip_count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) robot_job<NV, CHAIN, FLAGS, W>(m, prm, frame, io, io.manip_list[i]);
    return;
  }
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < io.B && (!io.count || b < *io.count)) robot_job<NV, CHAIN, FLAGS, W>(m, prm, frame, io, b);
#endif
}

// ---- narrow phase of one BLOCK of robots.  Closed-form stage: one robot per thread.  GJK stage: in rounds; every robot with a live
// candidate posts its most promising one (best first, as collision_job does) to a shared list, and the block's threads take the
// items in list order -- the GJK runs of a round sit in the first warps with all lanes busy, instead of each robot's runs being
// serialised inside its own diverged lane (ncu, round 2: 2.4 of 32 lanes active inside gjk_distance, 39 % of the kernel's samples).
// Per robot the sequence of GJK runs and every result are those of collision_job.  All threads of the block must call.
template <int T>
struct NarrowBlockSmem {
  int n_items;
  short owner[T];
  unsigned char bit[T];
  double best_d[T];
  GjkItemResult res[T];
};
template <int NV, bool CHAIN, int T, bool LOAD>
static __device__ __forceinline__ void collision_block(const DrcModelDev& m, const GeomTable& G, const DrcParams& prm, const CollisionIO& io,
                                                       int b0, bool valid, NarrowBlockSmem<T>& S) {
  const int tid = threadIdx.x;
  NarrowState st;
  st.cand = 0ull;
  if (valid) {
    if (LOAD) {  // the closed-form stage ran in its own launch (k_collision_closed): pick its result up
      const int b = b0 + tid;
      BestPair& best = st.best;
      best.d = io.dist[b]; best.id = io.pair_out[b];
      best.pa = v3(io.witness[6 * b + 0], io.witness[6 * b + 1], io.witness[6 * b + 2]);
      best.pb = v3(io.witness[6 * b + 3], io.witness[6 * b + 4], io.witness[6 * b + 5]);
      const int k = io.nar_k[b];
      best.ja = k >= 0 ? m.geom.parent[m.geom.pair_a[k]] : -1; best.jb = k >= 0 ? m.geom.parent[m.geom.pair_b[k]] : -1;
      st.cand = io.cand_mask[b]; st.deferred = 0ull;
      for (unsigned long long rem = st.cand; rem; rem &= rem - 1ull) {
        const int i = __ffsll((long long)rem) - 1;
        st.lbs[i] = io.nar_lb[(long long)i * io.Bc + b];
      }
    } else {
      narrow_closed_phase<NV, CHAIN>(m, io, b0 + tid, st);
    }
  }
  for (;;) {
    // every robot posts its most promising candidate ...
    const int bit0 = valid ? narrow_pick(st) : -1;
    constexpr int kSpec = 2;   // speculative candidates per robot and round (3 measured: no further gain)
    int slot0 = -1, bitx[kSpec], slotx[kSpec];
#pragma unroll
    for (int e = 0; e < kSpec; ++e) bitx[e] = -1;
    if (bit0 >= 0) {
      slot0 = atomicAdd(&S.n_items, 1);
      S.owner[slot0] = (short)tid; S.bit[slot0] = (unsigned char)bit0; S.best_d[tid] = st.best.d;
    }
    __syncthreads();
    // ... and the lanes that are still free take further candidates, speculatively: an item that an earlier candidate's outcome
    // would have culled cannot beat the running minimum (its distance is at least its bound), so the minimum over all pairs, its
    // tie-break and the set of overlapping pairs that matter are unchanged -- only the number of rounds drops
    if (bit0 >= 0) {
#pragma unroll
      for (int e = 0; e < kSpec; ++e) {
        if (*(volatile int*)&S.n_items >= T) break;
        const int bt = narrow_pick(st);
        if (bt < 0) break;
        const int sl = atomicAdd(&S.n_items, 1);
        if (sl >= T) { st.cand |= 1ull << bt; break; }   // no lane left: back into the candidate set
        S.owner[sl] = (short)tid; S.bit[sl] = (unsigned char)bt;
        bitx[e] = bt; slotx[e] = sl;
      }
    }
    __syncthreads();
    const int n = S.n_items < T ? S.n_items : T;
    if (n == 0) break;
    if (tid < n) {
      const int ow = S.owner[tid];
      narrow_gjk_item(m, G, prm, io, b0 + ow, S.bit[tid], S.best_d[ow], S.res[tid]);
    }
    __syncthreads();
    if (tid == 0) S.n_items = 0;
    if (bit0 >= 0) narrow_apply(m, G, st, bit0, S.res[slot0]);
#pragma unroll
    for (int e = 0; e < kSpec; ++e)
      if (bitx[e] >= 0) narrow_apply(m, G, st, bitx[e], S.res[slotx[e]]);
    __syncthreads();
  }
  if (valid) narrow_finish<NV, CHAIN>(m, prm, io, b0 + tid, st);
}

constexpr int kColThreads = 128;
// The narrow phase as TWO launches: the closed-form stage is straight-line FP64 code that lives on latency hiding -- on its own it
// fits a 128-register budget (16 warps / SM instead of the 8 the GJK code's 255 registers allow); the GJK stage + gradients follow.
template <int NV, bool CHAIN>
__global__ void __launch_bounds__(kColThreads, 4) k_collision_closed(const __grid_constant__ DrcModelDev m, const __grid_constant__ CollisionIO io) {
#ifndef DRC_SYNTAX_CHECK
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (!(b < io.B && (!io.count || b < *io.count))) return;
  NarrowState st;
  int best_k;
  narrow_closed_phase<NV, CHAIN>(m, io, b, st, &best_k);
  io.dist[b] = st.best.d; io.pair_out[b] = st.best.id; io.nar_k[b] = best_k;
  io.witness[6 * b + 0] = st.best.pa.x; io.witness[6 * b + 1] = st.best.pa.y; io.witness[6 * b + 2] = st.best.pa.z;
  io.witness[6 * b + 3] = st.best.pb.x; io.witness[6 * b + 4] = st.best.pb.y; io.witness[6 * b + 5] = st.best.pb.z;
  io.cand_mask[b] = st.cand;
  for (int i = 0; i < m.ngjk; ++i) io.nar_lb[(long long)i * io.Bc + b] = st.lbs[i];
#endif
}
#ifndef DRC_GJK_THREADS
#define DRC_GJK_THREADS 128
#endif
constexpr int kGjkThreads = DRC_GJK_THREADS;
template <int NV, bool CHAIN, bool LOAD, int MINB = 2>
__global__ void __launch_bounds__(kGjkThreads, MINB * (128 / kGjkThreads)) k_collision(const __grid_constant__ DrcModelDev m, const __grid_constant__ DrcParams prm,
                                                    const __grid_constant__ CollisionIO io) {
#ifndef DRC_SYNTAX_CHECK
  // stage the geometry table in shared memory: the GJK pass indexes it with per-thread pair ids
  __shared__ GeomTable G;
  __shared__ NarrowBlockSmem<kGjkThreads> S;
  {
    const int* src = reinterpret_cast<const int*>(&m.geom);
    int* dst = reinterpret_cast<int*>(&G);
    for (int i = threadIdx.x; i < (int)(sizeof(GeomTable) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
    if (threadIdx.x == 0) S.n_items = 0;
  }
  __syncthreads();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  collision_block<NV, CHAIN, kGjkThreads, LOAD>(m, G, prm, io, blockIdx.x * blockDim.x, b < io.B && (!io.count || b < *io.count), S);
#endif
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
static __device__ __forceinline__ void epa_make_face(EpaWarpSmem& S, int slot, int a, int b, int c) {
  const Vec3 nrm = cross(S.P[b].w - S.P[a].w, S.P[c].w - S.P[a].w);
  const double l = norm(nrm);
  const Vec3 n = l > 0 ? (1.0 / l) * nrm : v3(0, 0, 1);
  S.fv[slot][0] = (short)a; S.fv[slot][1] = (short)b; S.fv[slot][2] = (short)c;
  S.fn[slot][0] = n.x; S.fn[slot][1] = n.y; S.fn[slot][2] = n.z;
  S.fd[slot] = dot(n, S.P[a].w);
}
static __device__ void epa_warp(EpaWarpSmem& S, const Prim& A, const Prim& B, const GjkOut& g, double tol, int max_iter, PairResult& out,
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

// one flagged robot, handled by one warp: resolve its overlapping GJK-type pairs with EPA, then gradients / QP row
template <int NV, bool CHAIN>
static __device__ void epa_robot_warp(EpaWarpSmem& S, const DrcModelDev& m, const DrcParams& prm, const CollisionIO& io, int b, int lane) {
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

constexpr int kEpaWarps = 2;
template <int NV, bool CHAIN>
__global__ void __launch_bounds__(kEpaWarps * 32) k_collision_epa(const __grid_constant__ DrcModelDev m, const __grid_constant__ DrcParams prm,
                                                                   const __grid_constant__ CollisionIO io) {
#ifndef DRC_SYNTAX_CHECK
  __shared__ EpaWarpSmem sm[kEpaWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int count = *io.epa_count;
  for (int i = blockIdx.x * kEpaWarps + warp; i < count; i += gridDim.x * kEpaWarps) epa_robot_warp<NV, CHAIN>(sm[warp], m, prm, io, io.epa_list[i], lane);
#endif
}

// schedule bucket of an iteration count (descending: slow robots get the small buckets); used by k_admm and the k_sched_* kernels
constexpr int kSchedBuckets = 256;
static __device__ __forceinline__ int sched_bucket_of(int iters) {
  const int k = iters / 25;
  return kSchedBuckets - 1 - (k < kSchedBuckets ? k : kSchedBuckets - 1);
}
#ifndef DRC_ADMM_WARPS
#define DRC_ADMM_WARPS 1
#endif
constexpr int kAdmmWarps = DRC_ADMM_WARPS;  // one warp per block: a finished warp frees its slot without waiting for block-mates
#ifndef DRC_ADMM_MINBLOCKS   // lab builds measure other register budgets (blocks per SM; 12 = 168 registers)
#define DRC_ADMM_MINBLOCKS (MINB >= 8 ? MINB : (4 * MINB) / kAdmmWarps)   // MINB >= 8: blocks per SM given directly
#endif
template <class Cfg, bool ID, int MINB>
__global__ void __launch_bounds__(kAdmmWarps * 32, DRC_ADMM_MINBLOCKS) k_admm(const __grid_constant__ SolveIO io, const __grid_constant__ QpOptions o) {
#ifndef DRC_SYNTAX_CHECK
  extern __shared__ __align__(16) unsigned char admm_smem[];  // dynamic: the QPID record exceeds the 48 KB static limit
  GroupShared<Cfg>* sh = reinterpret_cast<GroupShared<Cfg>*>(admm_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // slots of this warp: entries [off + first, off + first + NG) of `order` (or the identity), below the slot count
  const int off = io.order_off ? *io.order_off : 0;
  const int nslot = io.count ? (*io.count < io.B ? *io.count : io.B) : io.B;
  const int first = off + (blockIdx.x * kAdmmWarps + warp) * Cfg::NG;
  int robots[Cfg::NG];
#pragma unroll
  for (int g = 0; g < Cfg::NG; ++g) {
    int r = first + g < nslot ? (io.order ? io.order[first + g] : first + g) : -1;
    if (r >= 0 && io.skip && io.skip[r]) r = -1;   // its self-collision row is still being computed (EPA pass): solved by the launch behind that pass
    robots[g] = r;
  }
  WarpExec<Cfg> w;
  w.sh = sh + warp * Cfg::NG;
  if (first < nslot) {   // no block-level barrier below: idle warps skip the solve
    w.lane = lane;
    lane_assign<Cfg>(w.L, lane);
    solve_and_emit<Cfg, ID>(w, robots, io, o);
  }
  if (io.hist_next) {
    // rollout ticks: histogram of this tick's iteration counts = the schedule buckets of the next tick; the LAST warp of the
    // launch turns it into the exclusive offsets the next tick's front kernel scatters with (no extra launch, no host sync)
    int mine = -1;
#pragma unroll
    for (int g = 0; g < Cfg::NG; ++g) if (lane == g) mine = robots[g];
    if (first < nslot && mine >= 0) atomicAdd(&io.hist_next[sched_bucket_of(w.sh[lane].iters)], 1);
    __threadfence();
    __syncwarp();
    int last = 0;
    if (lane == 0) last = atomicAdd(io.sched_ticket, 1) == (int)(gridDim.x * kAdmmWarps) - 1;
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
      __threadfence();
      constexpr int PER = kSchedBuckets / 32;
      int v[PER], sum = 0;
#pragma unroll
      for (int i = 0; i < PER; ++i) { v[i] = __ldcg(&io.hist_next[lane * PER + i]); sum += v[i]; }
      int incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
      int run = incl - sum;
#pragma unroll
      for (int i = 0; i < PER; ++i) { io.offs_next[lane * PER + i] = run; run += v[i]; io.hist_next[lane * PER + i] = 0; }
      if (lane == 0) *io.sched_ticket = 0;
    }
  }
#endif
}

// ---- ADMM schedule.  The iteration count of a QP is not known in advance and spreads from 50 to max_iter (4000):
// one late-starting slow robot keeps a single warp busy long after the rest of the grid has drained (measured: 41 %
// of the kernel), and the robots sharing a warp wait for its slowest member.  In a control loop consecutive ticks
// solve nearly the same QP, so the PREVIOUS tick's iteration count is an excellent predictor: robots are ordered by
// descending previous count (counting sort over count/25), i.e. longest first and warps of similar length.  Only the
// robot -> warp assignment changes; every robot's arithmetic, and therefore its result, is unaffected.
static __device__ __forceinline__ int sched_bucket(int iters) { return sched_bucket_of(iters); }
static __global__ void k_sched_hist(const int* prev, int B, int* hist) {
  __shared__ int h[kSchedBuckets];
  for (int i = threadIdx.x; i < kSchedBuckets; i += blockDim.x) h[i] = 0;
  __syncthreads();
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) atomicAdd(&h[sched_bucket(prev[b])], 1);
  __syncthreads();
  for (int i = threadIdx.x; i < kSchedBuckets; i += blockDim.x) if (h[i]) atomicAdd(&hist[i], h[i]);
}
static __global__ void k_sched_scan(int* hist, int slow_iters, int slow_max, int* slow_count) {  // one block of kSchedBuckets threads: exclusive prefix sum in place
  __shared__ int a[kSchedBuckets];
  const int t = threadIdx.x;
  a[t] = hist[t];
  __syncthreads();
  for (int off = 1; off < kSchedBuckets; off <<= 1) {
    const int v = t >= off ? a[t - off] : 0;
    __syncthreads();
    a[t] += v;
    __syncthreads();
  }
  hist[t] = a[t] - hist[t];
  // robots whose previous count reached slow_iters sort in front of bucket(slow_iters - 1)'s first entry: the priority launch
  if (slow_count && t == sched_bucket(slow_iters - 1)) *slow_count = hist[t] < slow_max ? hist[t] : slow_max;
}
static __global__ void k_sched_scatter(const int* prev, int B, int* offs, int* order) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) order[atomicAdd(&offs[sched_bucket(prev[b])], 1)] = b;
}

// DyrosMath::getTaskSpaceCubic (math_type_define.h:62-144,235-281,647-685); rotation log/exp by Rodrigues
struct CubicIO {
  int B;
  const double *x_target, *xdot_target, *x_init, *xdot_init;
  Strided s12, s6;
  double t, t0, dur;
  double *x_des, *xdot_des;
};
static __device__ __forceinline__ double cubic_pos(double t, double t0, double tf, double x0, double xf, double v0, double vf) {
  if (t < t0) return x0;
  if (t > tf) return xf;
  const double e = t - t0, T = tf - t0, T2 = T * T, T3 = T2 * T, dx = xf - x0;
  return x0 + v0 * e + (3 * dx / T2 - 2 * v0 / T - vf / T) * e * e + (-2 * dx / T3 + (v0 + vf) / T2) * e * e * e;
}
static __device__ __forceinline__ double cubic_vel(double t, double t0, double tf, double x0, double xf, double v0, double vf) {
  if (t < t0) return v0;
  if (t > tf) return vf;
  const double e = t - t0, T = tf - t0, T2 = T * T, T3 = T2 * T, dx = xf - x0;
  return v0 + 2 * (3 * dx / T2 - 2 * v0 / T - vf / T) * e + 3 * (-2 * dx / T3 + (v0 + vf) / T2) * e * e;
}
static __device__ Vec3 so3_log(const Mat3& R) {
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
static __device__ Mat3 so3_exp(Vec3 w) {
  const double th = norm(w);
  if (th < 1e-12) { Mat3 R = {{1, -w.z, w.y, w.z, 1, -w.x, -w.y, w.x, 1}}; return R; }
  double s, c;
  sincos(th, &s, &c);
  return rot_axis((1.0 / th) * w, s, c);
}
static __device__ __forceinline__ void task_cubic_job(const CubicIO& io, int b) {
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
static __global__ void k_task_cubic(const __grid_constant__ CubicIO io) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < io.B) task_cubic_job(io, b);
}

// ---- one control tick of a closed-loop rollout, front half: per robot, in ONE launch,
//   [schedule scatter for this tick's solver launch] -> [cubic profile] -> joint placements -> self-collision narrow phase
//   -> QPIK record;  the block's EPA-flagged robots (~0.1 %) are then resolved by warp 0 with the warp-parallel EPA.
// The solver launch (k_admm with SolveIO::roll_*) integrates the state and prepares the next tick's schedule: 2 launches per tick.
struct TickIO {
  JobIO job;            // q / qdot: the rollout state (read), x_target / xdot_target: the tick's desired pose / twist
  CollisionIO col;
  CubicIO cubic;        // B == 0: no profile (QPIKStep on the given target)
  const int* prev;      // previous tick's iteration counts; null: the schedule of this tick has been prepared by k_sched_*
  int* offs;            // exclusive bucket offsets (k_admm's last block, previous tick)
  int* order;
};
constexpr int kTickThreads = 128;
template <int NV, bool CHAIN>
__global__ void __launch_bounds__(kTickThreads) k_tick_front(const __grid_constant__ DrcModelDev m, const __grid_constant__ DrcParams prm,
                                                             const __grid_constant__ DrcFrame frame, const __grid_constant__ TickIO io) {
#ifndef DRC_SYNTAX_CHECK
  __shared__ GeomTable G;
  __shared__ EpaWarpSmem epa;
  __shared__ NarrowBlockSmem<kTickThreads> S;
  __shared__ int n_flag, flagged[kTickThreads];
  {
    const int* src = reinterpret_cast<const int*>(&m.geom);
    int* dst = reinterpret_cast<int*>(&G);
    for (int i = threadIdx.x; i < (int)(sizeof(GeomTable) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
    if (threadIdx.x == 0) { n_flag = 0; S.n_items = 0; }
  }
  __syncthreads();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = b < io.job.B;
  if (valid) {
    if (io.prev) io.order[atomicAdd(&io.offs[sched_bucket(io.prev[b])], 1)] = b;
    if (io.cubic.B > 0) task_cubic_job(io.cubic, b);
    robot_job<NV, CHAIN, F_STORE>(m, prm, frame, io.job, b);
  }
  collision_block<NV, CHAIN, kTickThreads, false>(m, G, prm, io.col, blockIdx.x * blockDim.x, valid, S);
  if (valid) {
    if (io.col.epa_flag[b]) flagged[atomicAdd(&n_flag, 1)] = b;
    robot_job<NV, CHAIN, F_FROM_CACHE | F_QPIK | F_STEP>(m, prm, frame, io.job, b);
  }
  __syncthreads();
  if (threadIdx.x < 32)
    for (int i = 0; i < n_flag; ++i) epa_robot_warp<NV, CHAIN>(epa, m, prm, io.col, flagged[i], threadIdx.x);
#endif
}

// ---- PinvCOD of full-model mass matrices that fail the Cholesky guard of spd_pinv (robot_data.cpp:118; whole-body models: every
// robot -- the arm's last joint has 1e-6 of the base's inertia, right at the reference's rank threshold), SIXTEEN LANES PER ROBOT.
// The thread-serial pinv_cpqr keeps R, Q and its work arrays in local memory (runtime pivots): 92 k instructions per robot at an IPC
// of 0.04, 77 % of the whole-body state kernel (ncu, round 2).  Here lane j of a group owns column j of R during the pivoted
// factorisation and row j of Q afterwards, in registers (the pivot loop is unrolled, a column's POSITION is a register); the
// pivot column, the reflectors and the final R travel through shared memory as broadcasts.  Same algorithm and rank rule as
// pinv_cpqr: A P = Q R, r = #{|R_kk| > threshold max|R_kk|}, pinv(A) = P pinv(R_1) Q_1' with R_1 = R_11 [I X] the first r rows:
//   pinv(R_1) = [I; X'] (I + X X')^-1 R_11^-1,  (I + X X')^-1 = I - X (I + X'X)^-1 X'   (X is r x (N - r): N - r is 0 or 1 here),
// i.e. two triangular solves per lane instead of the Gram matrix, its Cholesky factor and r solves of the serial routine.
constexpr int kPinvMaxDrop = 4;   // columns dropped by the rank rule that the lane-parallel path handles (more: serial routine)
template <int N>
struct PinvGroupSmem {
  double piv[16];
  double V[N][N];
  double vn2[N];
  double R[N][N];                 // final R, columns by POSITION
  double X[N][kPinvMaxDrop];
  int perm[16];                   // position -> original column
};
template <int N>
__global__ void __launch_bounds__(128) k_pinv_list(const double* c_M, double* c_Minv, long long Bc, const int* list, const int* count,
                                                    double threshold) {
#ifndef DRC_SYNTAX_CHECK
  static_assert(N <= 16, "one lane per column");
  __shared__ PinvGroupSmem<N> sm[8];
  const int lane = threadIdx.x & 31, g = lane & 15, grp = threadIdx.x >> 4;
  PinvGroupSmem<N>& S = sm[grp];
  const unsigned gmask = 0xffffu << (lane & 16);
  const bool act = g < N;
  const int n = *count, ngroups = gridDim.x * 8;
  for (int it = blockIdx.x * 8 + grp; it < n; it += ngroups) {
    const int b = list[it];
    double col[N];
#pragma unroll
    for (int i = 0; i < N; ++i) col[i] = act ? c_M[(long long)(i * N + g) * Bc + b] : 0.0;   // column g (M is symmetric)
    int pos = g;
    // ---- phase 1: Householder factorisation with column pivoting
#pragma unroll
    for (int k = 0; k < N; ++k) {
      double bn = -1.0;
      if (act && pos >= k) {
        bn = 0.0;
#pragma unroll
        for (int i = k; i < N; ++i) bn += col[i] * col[i];
      }
      int bp = pos;
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) {   // largest remaining column norm, ties to the smaller position (pinv_cpqr's scan order)
        const double on = __shfl_xor_sync(gmask, bn, off, 16);
        const int op = __shfl_xor_sync(gmask, bp, off, 16);
        if (on > bn || (on == bn && op < bp)) { bn = on; bp = op; }
      }
      if (act) { if (pos == bp) pos = k; else if (pos == k) pos = bp; }
      const double nrm = sqrt(bn);
      double vn2 = 0.0;
      if (nrm != 0.0) {
        if (act && pos == k) {
#pragma unroll
          for (int i = k; i < N; ++i) S.piv[i] = col[i];
        }
        __syncwarp(gmask);
        double v[N];
#pragma unroll
        for (int i = k; i < N; ++i) v[i] = S.piv[i];
        const double alpha = v[k] > 0 ? -nrm : nrm;
        v[k] -= alpha;
#pragma unroll
        for (int i = k; i < N; ++i) vn2 += v[i] * v[i];
        if (vn2 != 0.0) {
          if (act && pos >= k) {
            double sc = 0.0;
#pragma unroll
            for (int i = k; i < N; ++i) sc += v[i] * col[i];
            sc = 2 * sc / vn2;
#pragma unroll
            for (int i = k; i < N; ++i) col[i] -= sc * v[i];
          }
          if (g == 0) {
#pragma unroll
            for (int i = k; i < N; ++i) S.V[k][i] = v[i];
          }
        }
        __syncwarp(gmask);
      }
      if (g == 0) S.vn2[k] = vn2;
    }
    if (act) {
#pragma unroll
      for (int i = 0; i < N; ++i) S.R[i][pos] = col[i];
      S.perm[pos] = g;
    }
    __syncwarp(gmask);
    double maxpiv = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) maxpiv = dmax(maxpiv, fabs(S.R[k][k]));
    int r = 0;
#pragma unroll
    for (int k = 0; k < N; ++k) if (fabs(S.R[k][k]) > threshold * maxpiv) ++r;
    const int d = N - r;
    if (r == 0 || d > kPinvMaxDrop) {
      if (g == 0) {   // outside the lane-parallel path: the serial routine (never seen for a mass matrix)
        double A[N * N], P[N * N];
        for (int i = 0; i < N * N; ++i) A[i] = c_M[(long long)i * Bc + b];
        pinv_cpqr<N, N>(A, P, threshold);
        for (int i = 0; i < N * N; ++i) c_Minv[(long long)i * Bc + b] = P[i];
      }
      __syncwarp(gmask);
      continue;
    }
    // ---- phase 2: row g of Q = e_g' H_0 H_1 ...
    double y[N];
    {
      double qrow[N];
#pragma unroll
      for (int i = 0; i < N; ++i) qrow[i] = (i == g) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const double vn2 = S.vn2[k];
        if (vn2 != 0.0) {
          double sc = 0.0;
#pragma unroll
          for (int i = k; i < N; ++i) sc += qrow[i] * S.V[k][i];
          sc = 2 * sc / vn2;
#pragma unroll
          for (int i = k; i < N; ++i) qrow[i] -= sc * S.V[k][i];
        }
      }
      // y = R_11^-1 (Q_1' e_g)
#pragma unroll
      for (int i = N - 1; i >= 0; --i) {
        y[i] = 0.0;
        if (i < r) {
          double sc = qrow[i];
#pragma unroll
          for (int l = i + 1; l < N; ++l) if (l < r) sc -= S.R[i][l] * y[l];
          y[i] = sc / S.R[i][i];
        }
      }
    }
    if (d > 0) {
      if (g < d) {   // X = R_11^-1 R_12, one dropped column per lane
        double x[N];
#pragma unroll
        for (int i = N - 1; i >= 0; --i) {
          x[i] = 0.0;
          if (i < r) {
            double sc = S.R[i][r + g];
#pragma unroll
            for (int l = i + 1; l < N; ++l) if (l < r) sc -= S.R[i][l] * x[l];
            x[i] = sc / S.R[i][i];
          }
          S.X[i][g] = x[i];
        }
      }
      __syncwarp(gmask);
      // u = (I + X'X)^-1 X'y  (d x d, every lane on its own copy), then z = y - X u
      double w[kPinvMaxDrop], G[kPinvMaxDrop][kPinvMaxDrop];
#pragma unroll
      for (int c = 0; c < kPinvMaxDrop; ++c) {
        w[c] = 0.0;
#pragma unroll
        for (int c2 = 0; c2 < kPinvMaxDrop; ++c2) G[c][c2] = (c == c2) ? 1.0 : 0.0;
        if (c < d) {
#pragma unroll
          for (int i = 0; i < N; ++i) w[c] += S.X[i][c] * y[i];   // rows >= r of X and y are zero
#pragma unroll
          for (int c2 = 0; c2 < kPinvMaxDrop; ++c2) {
            if (c2 < d) {
              double sc = 0.0;
#pragma unroll
              for (int i = 0; i < N; ++i) sc += S.X[i][c] * S.X[i][c2];
              G[c][c2] += sc;
            }
          }
        }
      }
#pragma unroll
      for (int p = 0; p < kPinvMaxDrop; ++p) {   // SPD: elimination without pivoting (rows / columns >= d are the identity)
#pragma unroll
        for (int q2 = p + 1; q2 < kPinvMaxDrop; ++q2) {
          const double f = G[q2][p] / G[p][p];
#pragma unroll
          for (int t = p; t < kPinvMaxDrop; ++t) G[q2][t] -= f * G[p][t];
          w[q2] -= f * w[p];
        }
      }
#pragma unroll
      for (int p = kPinvMaxDrop - 1; p >= 0; --p) {
        double sc = w[p];
#pragma unroll
        for (int t = p + 1; t < kPinvMaxDrop; ++t) sc -= G[p][t] * w[t];
        w[p] = sc / G[p][p];
      }
#pragma unroll
      for (int i = 0; i < N; ++i) {
        if (i < r) {
#pragma unroll
          for (int c = 0; c < kPinvMaxDrop; ++c) if (c < d) y[i] -= S.X[i][c] * w[c];
        }
      }
    }
    // column g of the pseudo-inverse: positions < r take z, the dropped ones X'z; rows back in the original order
    if (act) {
#pragma unroll
      for (int p = 0; p < N; ++p) {
        double val = y[p];
        if (p >= r) {
          val = 0.0;
#pragma unroll
          for (int i = 0; i < N; ++i) val += S.X[i][p - r] * y[i];
        }
        c_Minv[(long long)(S.perm[p] * N + g) * Bc + b] = val;
      }
    }
    __syncwarp(gmask);
  }
#endif
}

// cache (SoA [K][Bc]) -> user array; `sub` != null writes src - sub (coriolis = nle - g)
static __global__ void k_copy_cache(const double* src, const double* sub, long long Bc, int K, int B, double* dst, Strided s) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)K * B) return;
  const int k = (int)(i / B), b = (int)(i % B);
  double v = src[(long long)k * Bc + b];
  if (sub) v -= sub[(long long)k * Bc + b];
  dst[b * s.sb + k * s.sk] = v;
}

// FP64 FMA peak: 8 independent chains per thread
static __global__ void k_fp64_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double x = 1.0000001, y = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
    a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}


}  // namespace drc_kernels
