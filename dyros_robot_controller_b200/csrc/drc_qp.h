// drc_b200 -- batched structured ADMM QP (OSQP's algorithm) for the reference's four controller QPs.
//
// Replaces QPBase::solveQP (reference include/dyros_robot_controller/QP_base.h:100-180: a fresh
// OsqpEigen::Solver per control cycle, cold start, OSQP defaults, success iff Solved).
//
// Structure exploited (all four QPs of the reference share it):
//   x = [ x_c (NC "core" variables: qdot / qddot / eta) ; singletons (slacks, torques) ]
//   P = blkdiag(P_cc, 0)
//   A = [ I (bound rows, optional) ; unit rows ; dense rows ; equality rows ]
//     unit row  (KU per core var j):  s x_c[j] (+ slack)        >= l      s = +1,-1,+1,-1
//     dense row (ND):                 a' x_c   (+ slack)        >= l      manipulability / self-collision
//     eq row    (NE):                 a' x_c   -  tau           == b      M qddot - tau = -g
//   every singleton appears in exactly one non-bound row ("bundle" = row + singleton + its bound row).
// Hence K = P + sigma I + A' diag(rho) A is [K_cc K_cd; K_dc diag] and OSQP's linear system reduces
// EXACTLY to the NC x NC Schur complement S; with U = [I; A_dense] one iteration's solve is
//   [x~_c ; A_dense x~_c] = (U S^-1 U') u          u = [local right-hand sides ; dense-row terms]
// i.e. ONE dot product of length GL = NC+ND+NE per lane against a vector exchanged through shared
// memory.  The iterates are OSQP's (same Ruiz scaling, rho vector, relaxation, termination and
// infeasibility tests, adaptive rho); only the linear algebra of the KKT solve differs (QDLDL vs
// Schur complement) -- rounding-level differences.
//
// Row state is carried as v = alpha z~ + (1-alpha) z + y/rho, from which OSQP's iterates follow:
//   z = Proj(v),  y = rho (v - Proj(v)),  rho z - y = rho (2 Proj(v) - v),  v+ = v + alpha (z~+ - Proj(v))
// (one register per row instead of two, no divisions in the loop).
//
// Work mapping: a GROUP of GL lanes owns one robot; a warp carries NG = 32/GL robots.
//   lane j < NC          : core variable j, its bound row, its KU unit bundles
//   lane NC+r (r<ND+NE)  : dense / equality row r as bundle 0
// Hot per-lane state lives in registers, cold data (scalings, deltas) in shared memory; the only
// cross-lane traffic is shared memory + __syncwarp().  The hot loop is branch-free: absent rows /
// singletons are encoded as zero coefficients.  The body is written against a `Warp` executor so
// tests/kernel_emu can run the identical code on the CPU (lanes emulated by a loop).
#pragma once
#include "drc_math.h"

namespace drc {

constexpr double kOsqpInfty = 1e30;
constexpr double kMinScaling = 1e-4, kMaxScaling = 1e4;
constexpr double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoTol = 1e-4, kRhoEqOverIneq = 1e3;

template <int NC_, int KU_, int ND_, int NE_, bool SLACK_, bool BOUNDS_>
struct QpCfg {
  static constexpr int NC = NC_, KU = KU_, ND = ND_, NE = NE_;
  static constexpr bool SLACK = SLACK_, BOUNDS = BOUNDS_;
  static constexpr int NR = ND + NE;                       // rows with a dense core part
  static constexpr int GL = NC + NR;                       // lanes per robot
  static constexpr int NG = 32 / GL;                       // robots per warp
  static constexpr int NB = KU > 0 ? KU : 1;               // bundle slots per lane
  static constexpr int NPK = NC * (NC + 1) / 2;
  // per-robot record in global memory (doubles)
  static constexpr int OFF_P = 0;
  static constexpr int OFF_Q = OFF_P + NPK;
  static constexpr int OFF_LO = OFF_Q + NC;
  static constexpr int OFF_HI = OFF_LO + NC;
  static constexpr int OFF_UNIT = OFF_HI + NC;             // [KU][NC] lower bounds of the unit rows
  static constexpr int OFF_ROW = OFF_UNIT + KU * NC;       // [NR][NC+1]: coefficients then l (== u for eq rows)
  static constexpr int STRIDE = OFF_ROW + NR * (NC + 1);
};

// 1/sqrt(x): the device uses the rsqrt sequence (cheaper than sqrt + divide), the host emulation 1/sqrt
DRC_HD double inv_sqrt(double x) {
#if defined(__CUDA_ARCH__)
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}
DRC_HD double limit_scaling(double d) {
  d = d < kMinScaling ? 1.0 : d;
  return d > kMaxScaling ? kMaxScaling : d;
}
DRC_HD bool is_inf_hi(double u) { return u > kOsqpInfty * kMinScaling; }
DRC_HD bool is_inf_lo(double l) { return l < -kOsqpInfty * kMinScaling; }
// OSQP set_rho_vec constraint classes: -1 loose (both infinite), 0 inequality, 1 equality
DRC_HD int row_class(double l, double u) {
  if (is_inf_lo(l) && is_inf_hi(u)) return -1;
  if (u - l < kRhoTol) return 1;
  return 0;
}
DRC_HD double class_rho(int cls, double rho) { return cls < 0 ? kRhoMin : (cls > 0 ? kRhoEqOverIneq * rho : rho); }

// Hot part of a bundle (registers).
struct BundleHot {
  double c;      // core multiplier: scaled coefficient of the own variable (unit rows) | 1 (dense rows) | 0 (absent)
  double v, vb;  // row state, singleton-bound-row state
  double xd;     // singleton value (scaled)
  double e, qd, beta, l;  // singleton coefficient in the row, its cost, its bound-row coefficient, row lower bound
  double kinv, gam;       // 1/kappa, rho e / kappa
  double pz, pzb, bd;     // scratch carried from phase A to phase B
};

template <class Cfg>
struct Lane {
  static constexpr bool kHasEq = Cfg::NE > 0;  // only equality bundles have row_eq / sb_free set
  int gl, grp;            // lane within the group (-1 = idle), group within the warp
  bool is_core;
  bool done;               // register copy of the group's done flag (refreshed after every termination check)
  bool row_eq;            // this lane's bundle rows are equalities  (projection onto {l})
  bool sb_free;           // this lane's singleton bound rows are (-inf, inf)
  double x, vc, q, betac, lc, uc, rhoc;  // core variable and its bound row
  double sig, al;         // sigma / alpha on core lanes, 0 on row lanes
  double rho_r, rho_b;    // rho of this lane's bundle rows / singleton bound rows
  double pzc;             // scratch
  double et;              // set-up scratch: D_temp of the own column (core lanes) / E_temp of the own row (row lanes)
  double pr[Cfg::NC];     // set-up scratch: row j of P (core lanes) / row r of A (row lanes) during the equilibration
  BundleHot b[Cfg::NB];
  double w[Cfg::GL];      // operator row of U S^-1 U'
};

// Cold per-lane data (shared memory; each slot is touched by its own lane only).
template <class Cfg>
struct ColdLane {
  double D, Ecb, dx, dyc, Dinv, Ecbinv;
  double E[Cfg::NB], Eb[Cfg::NB], Dd[Cfg::NB], dxd[Cfg::NB], dy[Cfg::NB], dyb[Cfg::NB];
  double Einv[Cfg::NB], Ebinv[Cfg::NB], Ddinv[Cfg::NB];
  int clsc, cls[Cfg::NB], clsb[Cfg::NB];
  unsigned char active[Cfg::NB], has_sing[Cfg::NB], has_sb[Cfg::NB];
};

constexpr int kNumRed = 16;  // group-wide reductions of one termination check
template <class Cfg>
struct GroupShared {
  double u[Cfg::GL];                                 // exchange vector of the hot loop
  double v[Cfg::GL];                                 // checks: dx_c | projected dy_r;  setup: D_temp / omega
  double xy[Cfg::GL];                                // checks: x_c | y_r
  double P[Cfg::NC * Cfg::NC];                       // scaled P_cc (symmetric, dense)
  double A[(Cfg::NR > 0 ? Cfg::NR : 1) * Cfg::NC];   // scaled dense-row coefficients
  // S (Schur complement -> inverse) and T (S^-1 a_r) live only inside factor(); red (per-lane partial
  // reductions) only inside the scaling / check phases: they share storage.
  static constexpr int kST = Cfg::NC * Cfg::NC + (Cfg::NR > 0 ? Cfg::NR : 1) * Cfg::NC;
  static constexpr int kScr = kST > Cfg::GL * kNumRed ? kST : Cfg::GL * kNumRed;
  union {
    struct { double S[Cfg::NC * Cfg::NC]; double T[(Cfg::NR > 0 ? Cfg::NR : 1) * Cfg::NC]; };
    double red[kScr];
  };
  ColdLane<Cfg> cold[Cfg::GL];
  double tot[kNumRed];                               // group totals of one termination check
  double c, cinv, rho, rho_prev;
  double pri_res, dua_res;
  int status, iters, done, robot, nx, rho_updates, need_factor;
};

struct QpOptions {
  double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, adaptive_rho_tolerance, slack_weight;
  int max_iter, check_termination, scaling, adaptive_rho, adaptive_rho_interval;
  unsigned unit_mask;  // bit j: core variable j owns unit rows (all ones for manipulator QPs)
};
DRC_HD QpOptions qp_options(const DrcParams& p, unsigned unit_mask) {
  QpOptions o;
  o.rho = p.rho; o.sigma = p.sigma; o.alpha = p.osqp_alpha; o.eps_abs = p.eps_abs; o.eps_rel = p.eps_rel;
  o.eps_prim_inf = p.eps_prim_inf; o.eps_dual_inf = p.eps_dual_inf; o.adaptive_rho_tolerance = p.adaptive_rho_tolerance;
  o.slack_weight = p.slack_weight; o.max_iter = p.max_iter; o.check_termination = p.check_termination;
  o.scaling = p.scaling; o.adaptive_rho = p.adaptive_rho; o.adaptive_rho_interval = p.adaptive_rho_interval;
  o.unit_mask = unit_mask;
  return o;
}

// ------------------------------------------------------------------------------------------------
// Warp executors.  Device: one register-resident Lane per thread, phases end with __syncwarp().
// ------------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
template <class Cfg>
struct WarpExec {
  Lane<Cfg> L;
  GroupShared<Cfg>* sh;  // this warp's NG groups
  int lane;
  template <class F>
  __device__ __forceinline__ void each(F f) {
    if (L.gl >= 0) f(L, sh[L.grp]);
    __syncwarp();
  }
  __device__ __forceinline__ GroupShared<Cfg>& group(int g) { return sh[g]; }
};
#endif
// Host emulation (tests/kernel_emu only).
template <class Cfg>
struct WarpEmu {
  Lane<Cfg> Ls[32];
  GroupShared<Cfg>* sh;
  template <class F>
  void each(F f) {
    for (int t = 0; t < 32; ++t)
      if (Ls[t].gl >= 0) f(Ls[t], sh[Ls[t].grp]);
  }
  GroupShared<Cfg>& group(int g) { return sh[g]; }
};

// projections of the row kinds
template <class LaneT>
DRC_HD double proj_row(const LaneT& L, double v, double l) { return (LaneT::kHasEq && L.row_eq) ? l : dmax(v, l); }
// max(v, 0) = (v + |v|) / 2 exactly: two FP64 operations instead of a compare + select chain
template <class LaneT>
DRC_HD double proj_sb(const LaneT& L, double v) { return (LaneT::kHasEq && L.sb_free) ? v : 0.5 * (v + fabs(v)); }

// ------------------------------------------------------------------------------------------------
// The solver.  `qp` points at the per-robot records (Cfg::STRIDE doubles each); robots[g] is the
// robot handled by group g of this warp (-1 = none).  On return the scaled iterates are still in
// the lanes; the caller unscales what it needs (x_c[j] = D x, singleton = Dd xd with D, Dd in the
// group's cold data) and reads status / iters from the group's shared record.
// ------------------------------------------------------------------------------------------------
// Optional warm start (an extension: the reference creates a fresh solver every cycle and never warm starts, QP_base.h:146).
// OSQP's osqp_warm_start semantics: x = D^-1 x0, z = A x, y = c E^-1 y0 in the scaled problem, then ordinary iterations.
// x0 / y0 are a previous solve's primal / dual vectors in the structured order of SolveIO::qp_x / qp_y; all zeros is the cold start.
struct WarmStart {
  const double* x = nullptr;   // [robot][NC (1 + KU) + NR]
  const double* y = nullptr;   // [robot][NC (1 + 2 KU) + 2 NR]
  const int* ids = nullptr;    // optional: QP slot -> robot index of the two arrays
};

template <class Cfg, class W>
DRC_HD void admm_solve(W& w, const double* qp, const int* robots, const QpOptions& o, const WarmStart& warm = WarmStart()) {
  constexpr int NC = Cfg::NC, KU = Cfg::KU, ND = Cfg::ND, NR = Cfg::NR, GL = Cfg::GL, NG = Cfg::NG, NB = Cfg::NB;
  typedef Lane<Cfg> LaneT;
  typedef GroupShared<Cfg> GS;
  typedef ColdLane<Cfg> Cold;
  const double sigma = o.sigma, alpha = o.alpha, oma = 1.0 - o.alpha;

  // ---------------------------------------------------------------- load
  DRC_PHASE(PH_QP_LOAD);
  w.each([&](LaneT& L, GS& S) {
    const int rb = robots[L.grp];
    Cold& C = S.cold[L.gl];
    if (L.gl == 0) {
      S.robot = rb; S.status = kQpUnsolved; S.iters = 0; S.done = rb < 0 ? 1 : 0; S.c = 1.0; S.rho_updates = 0;
      S.need_factor = 0; S.pri_res = 0; S.dua_res = 0;
      int na = 0;
      for (int j = 0; j < NC; ++j) na += (int)((o.unit_mask >> j) & 1u);
      S.nx = NC + (Cfg::SLACK ? na * KU + ND : 0) + Cfg::NE;
    }
    L.x = 0; L.vc = 0; L.q = 0; L.betac = 0; L.lc = -kOsqpInfty; L.uc = kOsqpInfty; L.rhoc = 0;
    L.sig = L.is_core ? sigma : 0.0; L.al = L.is_core ? alpha : 0.0;
    L.rho_r = 0; L.rho_b = 0; L.pzc = 0; L.et = 1.0;
    L.row_eq = false; L.sb_free = false; L.done = rb < 0;
    C.D = 1; C.Ecb = 1; C.dx = 0; C.dyc = 0; C.clsc = -1;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      BundleHot& b = L.b[k];
      b.c = 0; b.v = 0; b.vb = 0; b.xd = 0; b.e = 0; b.qd = 0; b.beta = 0; b.l = -kOsqpInfty; b.kinv = 0; b.gam = 0;
      b.pz = 0; b.pzb = 0; b.bd = 0;
      C.E[k] = 1; C.Eb[k] = 1; C.Dd[k] = 1; C.dxd[k] = 0; C.dy[k] = 0; C.dyb[k] = 0; C.cls[k] = 0; C.clsb[k] = 0;
      C.active[k] = 0; C.has_sing[k] = 0; C.has_sb[k] = 0;
    }
    if (rb < 0) return;
    const double* rec = qp + (long long)rb * Cfg::STRIDE;
    if (L.is_core) {
      const int j = L.gl;
#pragma unroll
      for (int i = 0; i < NC; ++i) L.pr[i] = rec[Cfg::OFF_P + symidx<NC>(i, j)];
      L.q = rec[Cfg::OFF_Q + j];
      if (Cfg::BOUNDS) { L.betac = 1.0; L.lc = rec[Cfg::OFF_LO + j]; L.uc = rec[Cfg::OFF_HI + j]; }
      const bool act = ((o.unit_mask >> j) & 1u) != 0u;
#pragma unroll
      for (int k = 0; k < KU; ++k) {
        if (!act) continue;
        BundleHot& b = L.b[k];
        C.active[k] = 1;
        b.c = (k & 1) ? -1.0 : 1.0;
        b.l = rec[Cfg::OFF_UNIT + k * NC + j];
        if (Cfg::SLACK) {
          C.has_sing[k] = 1; b.e = 1.0; b.qd = o.slack_weight;
          if (Cfg::BOUNDS) { C.has_sb[k] = 1; b.beta = 1.0; }
        }
      }
    } else {
      const int r = L.gl - NC;
      const double* row = rec + Cfg::OFF_ROW + r * (NC + 1);
#pragma unroll
      for (int i = 0; i < NC; ++i) { L.pr[i] = row[i]; S.A[r * NC + i] = row[i]; }
      BundleHot& b = L.b[0];
      C.active[0] = 1;
      b.c = 1.0;
      b.l = row[NC];
      if (r >= ND) {  // equality row with its torque singleton
        L.row_eq = true; L.sb_free = true;
        C.has_sing[0] = 1; b.e = -1.0; b.qd = 0.0;
        if (Cfg::BOUNDS) { C.has_sb[0] = 1; b.beta = 1.0; }
      } else if (Cfg::SLACK) {
        C.has_sing[0] = 1; b.e = 1.0; b.qd = o.slack_weight;
        if (Cfg::BOUNDS) { C.has_sb[0] = 1; b.beta = 1.0; }
      }
    }
  });

  // ---------------------------------------------------------------- Ruiz equilibration (OSQP scale_data)
  // Two phases per pass.  During the passes every lane keeps ITS row of P (core lanes) / of A (row lanes) in registers
  // (L.pr): P is symmetric, so lane j only ever needs row j; the row lanes publish their scaled row to shared memory for
  // the column norms of the core lanes.  The KU unit bundles of a core lane start from the same magnitudes
  // (|c| = e = beta = 1, cost = slack weight) and every scaling factor depends on magnitudes only, so they stay
  // identical through all passes: bundle 0 is scaled as their representative and copied to the others afterwards.
  // The cost normalisation of a pass (c_temp) is applied at the start of the next one (apply_cost), which saves a phase.
  DRC_PHASE(PH_QP_SCALE);
  auto apply_cost = [&](LaneT& L, GS& S) {
    // c_temp = 1 / limit(max(mean_j ||P_j||_inf, limit(||q||_inf))) from the per-lane values published by phase Y
    double sum = 0, qm = 0;
#pragma unroll
    for (int i = 0; i < NC; ++i) sum += S.red[i];
#pragma unroll
    for (int i = 0; i < GL; ++i) qm = dmax(qm, S.red[GL + i]);
    const double ct = 1.0 / limit_scaling(dmax(sum / (double)S.nx, limit_scaling(qm)));
    if (L.is_core) {
      L.q *= ct;
#pragma unroll
      for (int i = 0; i < NC; ++i) L.pr[i] *= ct;
    }
    L.b[0].qd *= ct;
    if (L.gl == 0) S.c *= ct;
  };
  for (int it = 0; it < o.scaling; ++it) {
    // (X) D_temp of the core columns, E_temp of every row, D_temp of the singleton columns; apply to everything lane-local
    w.each([&](LaneT& L, GS& S) {
      if (S.done) return;
      Cold& C = S.cold[L.gl];
      if (it > 0) apply_cost(L, S);
      BundleHot& b = L.b[0];
      double core_norm, dj = 1.0;
      if (L.is_core) {
        const int j = L.gl;
        double cn = dmax(fabs(L.betac), fabs(b.c));
#pragma unroll
        for (int i = 0; i < NC; ++i) cn = dmax(cn, fabs(L.pr[i]));
#pragma unroll
        for (int r = 0; r < NR; ++r) cn = dmax(cn, fabs(S.A[r * NC + j]));
        dj = inv_sqrt(limit_scaling(cn));
        S.u[j] = dj;
        core_norm = fabs(b.c);
      } else {
        double rn = 0;
#pragma unroll
        for (int i = 0; i < NC; ++i) rn = dmax(rn, fabs(L.pr[i]));
        core_norm = rn;
      }
      double et = 1.0;
      if (C.active[0]) {
        double dt = 1.0, eb = 1.0;
        if (C.has_sing[0]) {
          double cn = fabs(b.e);
          if (C.has_sb[0]) { cn = dmax(cn, fabs(b.beta)); eb = inv_sqrt(limit_scaling(fabs(b.beta))); }
          dt = inv_sqrt(limit_scaling(cn));
        }
        et = inv_sqrt(limit_scaling(dmax(core_norm, fabs(b.e))));
        C.E[0] *= et;
        if (C.has_sing[0]) {
          b.e *= et * dt; C.Dd[0] *= dt; b.qd *= dt;
          if (C.has_sb[0]) { b.beta *= eb * dt; C.Eb[0] *= eb; }
        }
      }
      if (L.is_core) {
        if (Cfg::BOUNDS) { const double eb = inv_sqrt(limit_scaling(fabs(L.betac))); L.betac *= eb * dj; C.Ecb *= eb; }
        b.c *= et * dj;
        L.q *= dj; C.D *= dj;
        L.et = dj;
      } else {
        L.et = et;
      }
    });
    // (Y) scale P and the dense rows; publish the inputs of the cost normalisation
    w.each([&](LaneT& L, GS& S) {
      if (S.done) return;
      Cold& C = S.cold[L.gl];
      double cn = 0, qm = 0;
      if (L.is_core) {
#pragma unroll
        for (int i = 0; i < NC; ++i) { const double pv = L.pr[i] * L.et * S.u[i]; L.pr[i] = pv; cn = dmax(cn, fabs(pv)); }
        qm = fabs(L.q);
      } else {
        const int r = L.gl - NC;
#pragma unroll
        for (int i = 0; i < NC; ++i) { const double av = L.pr[i] * (L.et * S.u[i]); L.pr[i] = av; S.A[r * NC + i] = av; }
      }
      if (C.has_sing[0]) qm = dmax(qm, fabs(L.b[0].qd));
      S.red[L.gl] = cn; S.red[GL + L.gl] = qm;
    });
  }
  // last cost normalisation, P to shared memory, unit bundles 1.. = copies of bundle 0, scaled bounds, constraint classes
  w.each([&](LaneT& L, GS& S) {
    if (S.done) return;
    Cold& C = S.cold[L.gl];
    if (o.scaling > 0) apply_cost(L, S);
    if (L.is_core) {
      const int j = L.gl;
#pragma unroll
      for (int i = 0; i < NC; ++i) S.P[j * NC + i] = L.pr[i];
#pragma unroll
      for (int k = 1; k < KU; ++k) {
        if (!C.active[k]) continue;
        L.b[k].c = (k & 1) ? -L.b[0].c : L.b[0].c;
        L.b[k].e = L.b[0].e; L.b[k].beta = L.b[0].beta; L.b[k].qd = L.b[0].qd;
        C.E[k] = C.E[0]; C.Eb[k] = C.Eb[0]; C.Dd[k] = C.Dd[0];
      }
    }
    if (L.is_core && Cfg::BOUNDS) { L.lc *= C.Ecb; L.uc *= C.Ecb; C.clsc = row_class(L.lc, L.uc); }
    C.Dinv = 1.0 / C.D; C.Ecbinv = 1.0 / C.Ecb;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      C.Einv[k] = 1.0 / C.E[k]; C.Ebinv[k] = 1.0 / C.Eb[k]; C.Ddinv[k] = 1.0 / C.Dd[k];
      if (!C.active[k]) continue;
      L.b[k].l *= C.E[k];
      const double us = L.row_eq ? L.b[k].l : kOsqpInfty * C.E[k];
      C.cls[k] = row_class(L.b[k].l, us);
      if (C.has_sb[k]) C.clsb[k] = L.sb_free ? row_class(-kOsqpInfty * C.Eb[k], kOsqpInfty * C.Eb[k]) : row_class(0.0, kOsqpInfty * C.Eb[k]);
    }
    if (L.gl == 0) { S.cinv = 1.0 / S.c; S.rho = dmin(dmax(o.rho, kRhoMin), kRhoMax); S.rho_prev = S.rho; S.need_factor = 1; }
  });

  // ---------------------------------------------------------------- factorisation (setup and rho updates)
  auto factor = [&]() {
    DRC_PHASE(PH_QP_FACTOR);
    // (a) rho vector, per-bundle constants, Schur complement S = P + sigma I + diag(...) + sum_r omega_r a_r a_r'
    w.each([&](LaneT& L, GS& S) {
      if (!S.need_factor) return;
      Cold& C = S.cold[L.gl];
      const double rho = S.rho, ratio = S.rho_prev / S.rho;
      double diag = sigma, omega0 = 0;
      if (L.is_core && Cfg::BOUNDS) {
        const double rn = class_rho(C.clsc, rho);
        if (C.clsc >= 0) { const double pz = clampd(L.vc, L.lc, L.uc); L.vc = pz + ratio * (L.vc - pz); }
        L.rhoc = rn;
        diag += rn * L.betac * L.betac;
      }
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        BundleHot& b = L.b[k];
        if (!C.active[k]) continue;
        const double rr = class_rho(C.cls[k], rho);
        if (C.cls[k] >= 0) { const double pz = proj_row(L, b.v, b.l); b.v = pz + ratio * (b.v - pz); }
        L.rho_r = rr;
        double omega = rr;
        if (C.has_sing[k]) {
          double kap = sigma + rr * b.e * b.e;
          if (C.has_sb[k]) {
            const double rbv = class_rho(C.clsb[k], rho);
            if (C.clsb[k] >= 0) { const double pz = proj_sb(L, b.vb); b.vb = pz + ratio * (b.vb - pz); }
            L.rho_b = rbv;
            kap += rbv * b.beta * b.beta;
          }
          b.kinv = 1.0 / kap;
          b.gam = rr * b.e * b.kinv;
          omega = rr * (1.0 - b.gam * b.e);
        }
        if (L.is_core) diag += omega * b.c * b.c;
        else omega0 = omega;
      }
      if (L.is_core) {
        const int j = L.gl;
#pragma unroll
        for (int i = 0; i < NC; ++i) S.S[j * NC + i] = S.P[j * NC + i] + (i == j ? diag : 0.0);
      } else {
        S.v[L.gl] = omega0;
      }
    });
    w.each([&](LaneT& L, GS& S) {
      if (!S.need_factor || !L.is_core) return;
      const int j = L.gl;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const double wa = S.v[NC + r] * S.A[r * NC + j];
#pragma unroll
        for (int i = 0; i < NC; ++i) S.S[j * NC + i] += wa * S.A[r * NC + i];
      }
    });
    // (b) in-place Gauss-Jordan inverse (S is SPD: no pivoting); the pivot loop stays rolled (code size)
#pragma unroll 1
    for (int kk = 0; kk < NC; ++kk) {
      w.each([&](LaneT& L, GS& S) {
        if (!S.need_factor || !L.is_core) return;
        const int j = L.gl;
        const double piv = 1.0 / S.S[kk * NC + kk];
        S.u[j] = (j == kk) ? piv : S.S[j * NC + kk] * piv;
      });
      w.each([&](LaneT& L, GS& S) {
        if (!S.need_factor || !L.is_core) return;
        const int j = L.gl;
        if (j == kk) return;
        const double f = S.u[j];
#pragma unroll
        for (int i = 0; i < NC; ++i)
          if (i != kk) S.S[j * NC + i] -= f * S.S[kk * NC + i];
        S.S[j * NC + kk] = -f;
      });
      w.each([&](LaneT& L, GS& S) {
        if (!S.need_factor || !L.is_core || L.gl != kk) return;
        const double piv = S.u[kk];
#pragma unroll
        for (int i = 0; i < NC; ++i) S.S[kk * NC + i] = (i == kk) ? piv : S.S[kk * NC + i] * piv;
      });
    }
    // (c) T_r = S^-1 a_r and the operator rows of U S^-1 U'
    w.each([&](LaneT& L, GS& S) {
      if (!S.need_factor || L.is_core) return;
      const int r = L.gl - NC;
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < NC; ++k) s += S.S[i * NC + k] * S.A[r * NC + k];
        S.T[r * NC + i] = s;
      }
    });
    w.each([&](LaneT& L, GS& S) {
      if (!S.need_factor) return;
      if (L.is_core) {
        const int j = L.gl;
#pragma unroll
        for (int i = 0; i < NC; ++i) L.w[i] = S.S[j * NC + i];
#pragma unroll
        for (int r = 0; r < NR; ++r) L.w[NC + r] = S.T[r * NC + j];
      } else {
        const int r = L.gl - NC;
#pragma unroll
        for (int i = 0; i < NC; ++i) L.w[i] = S.T[r * NC + i];
#pragma unroll
        for (int r2 = 0; r2 < NR; ++r2) {
          double s = 0;
#pragma unroll
          for (int i = 0; i < NC; ++i) s += S.A[r * NC + i] * S.T[r2 * NC + i];
          L.w[NC + r2] = s;
        }
      }
    });
    w.each([&](LaneT& L, GS& S) { if (L.gl == 0) { S.need_factor = 0; S.rho_prev = S.rho; } });
  };
  auto warp_done = [&]() {
    bool d = true;
    for (int g = 0; g < NG; ++g) d = d && (w.group(g).done != 0);
    return d;
  };

  // ---------------------------------------------------------------- hot loop phases (branch-free)
  // phase A: u = [sigma x - q + A_local'(rho z - y) ; t_r]
  // (OSQP's cold start sets z = y = 0 without projecting: on the first iteration Proj(v) is replaced by 0)
  // (warm start: z of the first iteration is the given A x0, preloaded into pzc / pz / pzb, instead of a projection)
  auto phase_a_impl = [&](LaneT& L, GS& S, const bool first, const bool warm_first = false) {
    if (L.done) return;
    const double pzc = warm_first ? L.pzc : (first ? 0.0 : clampd(L.vc, L.lc, L.uc));
    L.pzc = pzc;
    double u = L.sig * L.x - L.q + L.betac * (L.rhoc * (2.0 * pzc - L.vc));
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      BundleHot& b = L.b[k];
      const double pz = warm_first ? b.pz : (first ? 0.0 : proj_row(L, b.v, b.l)), pzb = warm_first ? b.pzb : (first ? 0.0 : proj_sb(L, b.vb));
      // singleton bound row: 2 Proj(vb) - vb = |vb| on [0, inf), vb on a free row (exact)
      const double wr = L.rho_r * (2.0 * pz - b.v);
      const double wb = L.rho_b * (warm_first ? (2.0 * pzb - b.vb) : (first ? -b.vb : ((LaneT::kHasEq && L.sb_free) ? b.vb : fabs(b.vb))));
      const double bd = sigma * b.xd - b.qd + b.e * wr + b.beta * wb;
      b.pz = pz; b.pzb = pzb; b.bd = bd;
      u += b.c * (wr - b.gam * bd);
    }
    S.u[L.gl] = u;
  };
  auto phase_a = [&](LaneT& L, GS& S) { phase_a_impl(L, S, false); };
  auto phase_a_first = [&](LaneT& L, GS& S) { phase_a_impl(L, S, true); };
  auto phase_a_warm = [&](LaneT& L, GS& S) { phase_a_impl(L, S, false, true); };
  const bool warm_on = warm.x != nullptr;
  // phase B: s = (U S^-1 U' u)_lane, then the x / v updates
  auto phase_b = [&](LaneT& L, GS& S) {
    if (L.done) return;
    double s0 = 0, s1 = 0, s2 = 0;  // three independent chains (fp64 FMA latency)
#pragma unroll
    for (int i = 0; i < GL; i += 3) {
      s0 += L.w[i] * S.u[i];
      if (i + 1 < GL) s1 += L.w[i + 1] * S.u[i + 1];
      if (i + 2 < GL) s2 += L.w[i + 2] * S.u[i + 2];
    }
    const double s = (s0 + s1) + s2;
    L.x = L.al * s + oma * L.x;
    L.vc += alpha * (L.betac * s - L.pzc);
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      BundleHot& b = L.b[k];
      const double sk = b.c * s;
      const double xtd = b.kinv * b.bd - b.gam * sk;
      b.v += alpha * (sk + b.e * xtd - b.pz);
      b.vb += alpha * (b.beta * xtd - b.pzb);
      b.xd = alpha * xtd + oma * b.xd;
    }
  };
  // project delta_y onto the polar of the recession cone of [l,u] (is_primal_infeasible)
  auto proj_dy = [](double dy, bool lo_inf, bool hi_inf) {
    if (hi_inf) return lo_inf ? 0.0 : dmin(dy, 0.0);
    if (lo_inf) return dmax(dy, 0.0);
    return dy;
  };
  // phase B on a checked iteration: also records delta_x / delta_y of this iteration (cold data) and publishes the
  // exchange vectors of the termination check: x_c | y_r  and  dx_c | projected dy_r
  auto phase_b_keep = [&](LaneT& L, GS& S) {
    if (L.done) return;
    Cold& C = S.cold[L.gl];
    double s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
    for (int i = 0; i < GL; i += 3) {
      s0 += L.w[i] * S.u[i];
      if (i + 1 < GL) s1 += L.w[i + 1] * S.u[i + 1];
      if (i + 2 < GL) s2 += L.w[i + 2] * S.u[i + 2];
    }
    const double s = (s0 + s1) + s2;
    const double xn = L.al * s + oma * L.x;
    C.dx = xn - L.x;
    L.x = xn;
    {
      const double y0 = L.rhoc * (L.vc - L.pzc);
      L.vc += alpha * (L.betac * s - L.pzc);
      C.dyc = L.rhoc * (L.vc - clampd(L.vc, L.lc, L.uc)) - y0;
    }
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      BundleHot& b = L.b[k];
      const double sk = b.c * s;
      const double xtd = b.kinv * b.bd - b.gam * sk;
      const double y0 = L.rho_r * (b.v - b.pz), yb0 = L.rho_b * (b.vb - b.pzb);
      b.v += alpha * (sk + b.e * xtd - b.pz);
      b.vb += alpha * (b.beta * xtd - b.pzb);
      const double xdn = alpha * xtd + oma * b.xd;
      C.dxd[k] = xdn - b.xd;
      b.xd = xdn;
      C.dy[k] = L.rho_r * (b.v - proj_row(L, b.v, b.l)) - y0;
      C.dyb[k] = L.rho_b * (b.vb - proj_sb(L, b.vb)) - yb0;
    }
    if (L.is_core) { S.xy[L.gl] = L.x; S.v[L.gl] = C.dx; }
    else {
      const BundleHot& b = L.b[0];
      S.xy[L.gl] = L.rho_r * (b.v - proj_row(L, b.v, b.l));
      S.v[L.gl] = proj_dy(C.dy[0], is_inf_lo(b.l), !L.row_eq && is_inf_hi(kOsqpInfty * C.E[0]));
    }
  };

  // ---------------------------------------------------------------- warm start (osqp_warm_start), optional
  if (warm_on) {
    factor();   // the rho vector of the first factorisation is needed to place y0 in the row states
    constexpr int NX = NC * (1 + KU) + NR, NY = NC * (1 + 2 * KU) + 2 * NR;
    // (W1) x = D^-1 x0; the scaled core part goes to shared memory for the dense rows
    w.each([&](LaneT& L, GS& S) {
      if (S.done) return;
      const Cold& C = S.cold[L.gl];
      const double* x0 = warm.x + (long long)(warm.ids ? warm.ids[S.robot] : S.robot) * NX;
      if (L.is_core) {
        const int j = L.gl;
        L.x = x0[j] * C.Dinv;
        S.xy[j] = L.x;
#pragma unroll
        for (int k = 0; k < KU; ++k) L.b[k].xd = C.has_sing[k] ? x0[NC * (1 + k) + j] * C.Ddinv[k] : 0.0;
      } else {
        L.b[0].xd = C.has_sing[0] ? x0[NC * (1 + KU) + (L.gl - NC)] * C.Ddinv[0] : 0.0;
      }
    });
    // (W2) z = A x, y = c E^-1 y0, row states v = z + y / rho; z stays in pzc / pz / pzb for the first iteration
    w.each([&](LaneT& L, GS& S) {
      if (S.done) return;
      const Cold& C = S.cold[L.gl];
      const double* y0 = warm.y + (long long)(warm.ids ? warm.ids[S.robot] : S.robot) * NY;
      auto place = [&](double z, double y_unscaled, double Einv, double rho, double& pz, double& v) {
        pz = z;
        v = rho > 0 ? z + (S.c * Einv * y_unscaled) / rho : z;
      };
      if (L.is_core) {
        const int j = L.gl;
        if (Cfg::BOUNDS) place(L.betac * L.x, y0[j], C.Ecbinv, L.rhoc, L.pzc, L.vc);
#pragma unroll
        for (int k = 0; k < KU; ++k) {
          BundleHot& b = L.b[k];
          if (!C.active[k]) continue;
          place(b.c * L.x + b.e * b.xd, y0[NC * (1 + k) + j], C.Einv[k], L.rho_r, b.pz, b.v);
          if (C.has_sb[k]) place(b.beta * b.xd, y0[NC * (1 + KU + k) + j], C.Ebinv[k], L.rho_b, b.pzb, b.vb);
        }
      } else {
        const int r = L.gl - NC;
        BundleHot& b = L.b[0];
        double ax = b.e * b.xd;
#pragma unroll
        for (int i = 0; i < NC; ++i) ax += S.A[r * NC + i] * S.xy[i];
        place(ax, y0[NC * (1 + 2 * KU) + r], C.Einv[0], L.rho_r, b.pz, b.v);
        if (C.has_sb[0]) place(b.beta * b.xd, y0[NC * (1 + 2 * KU) + NR + r], C.Ebinv[0], L.rho_b, b.pzb, b.vb);
      }
    });
  }

  // ---------------------------------------------------------------- ADMM iterations (osqp_solve)
  int iter = 0;
  bool all_done = warp_done();
  const int chk = o.check_termination > 0 ? o.check_termination : 0x7fffffff;
  const int adp = (o.adaptive_rho && o.adaptive_rho_interval > 0) ? o.adaptive_rho_interval : 0x7fffffff;
  int to_check = chk, to_adapt = adp;  // iterations left until the next termination check / rho adaptation
  while (!all_done && iter < o.max_iter) {
    {  // (re)factorise when a group asks for it: setup, and after an accepted rho update
      bool nf = false;
      for (int g = 0; g < NG; ++g) nf = nf || (w.group(g).need_factor != 0 && w.group(g).done == 0);
      if (nf) factor();
    }
    DRC_PHASE(PH_QP_ITER);
    // plain iterations up to (excluding) the next event: nothing but the two hot phases
    int nplain = to_check < to_adapt ? to_check : to_adapt;
    if (nplain > o.max_iter - iter) nplain = o.max_iter - iter;
    nplain -= 1;
    for (int k = 0; k < nplain; ++k) {
      ++iter;
      if (iter == 1) { if (warm_on) w.each(phase_a_warm); else w.each(phase_a_first); } else w.each(phase_a);
      w.each(phase_b);
    }
    to_check -= nplain; to_adapt -= nplain;
    // the event iteration
    ++iter; --to_check; --to_adapt;
    const bool last = iter == o.max_iter;
    const bool can_check = to_check == 0 || last;
    const bool can_adapt = to_adapt == 0;
    if (to_check == 0) to_check = chk;
    if (to_adapt == 0) to_adapt = adp;
    if (iter == 1) { if (warm_on) w.each(phase_a_warm); else w.each(phase_a_first); } else w.each(phase_a);
    w.each(phase_b_keep);

    // ---------------- OSQP update_info + check_termination + adapt_rho
    DRC_PHASE(PH_QP_CHECK);
    // C1: per-lane partial reductions (registers; published once at the end).  Group totals:
    //   rows:      0 pri_u | 1 max(ax_u, z_u) | 2 pri_s | 3 max(ax_s, z_s) | 4 ||E dy|| | 5 sum(u dy+ + l dy-)
    //              6 max Adx over rows with a finite u | 7 max -Adx over rows with a finite l
    //   variables: 8 dua_u | 9 max(px_u, aty_u, q_u) | 10 dua_s | 11 max(px_s, aty_s, q_s) | 12 ||Dinv A'dy||
    //              13 ||D dx|| | 14 q'dx | 15 ||Dinv P dx||
    // (_u = unscaled, what OSQP tests with scaled_termination off; _s = scaled, what adapt_rho uses.)
    // One code path for core and row lanes: bundle 0 is the dense / equality row of a row lane and the first unit bundle
    // of a core lane; only its core part (A_row x_c) differs.
    w.each([&](LaneT& L, GS& S) {
      if (S.done) return;
      Cold& C = S.cold[L.gl];
      double m[kNumRed];
#pragma unroll
      for (int i = 0; i < kNumRed; ++i) m[i] = 0.0;
      m[6] = -1e300; m[7] = -1e300;
      auto row_acc = [&](double ax, double z, double E, double Ei, double l, double u, bool lo_inf, bool hi_inf, double dy, double adx) {
        const double r = ax - z;
        m[0] = dmax(m[0], fabs(Ei * r)); m[1] = dmax(m[1], dmax(fabs(Ei * ax), fabs(Ei * z)));
        m[2] = dmax(m[2], fabs(r)); m[3] = dmax(m[3], dmax(fabs(ax), fabs(z)));
        const double pdy = proj_dy(dy, lo_inf, hi_inf);
        m[4] = dmax(m[4], fabs(E * pdy));
        m[5] += u * dmax(pdy, 0.0) + l * dmin(pdy, 0.0);
        if (!hi_inf) m[6] = dmax(m[6], Ei * adx);
        if (!lo_inf) m[7] = dmax(m[7], -(Ei * adx));
      };
      auto var_acc = [&](double px, double aty, double q, double D, double Di, double atdy, double dx, double pdx) {
        const double r = px + q + aty;
        m[8] = dmax(m[8], fabs(Di * r)); m[9] = dmax(m[9], dmax(fabs(Di * px), dmax(fabs(Di * aty), fabs(Di * q))));
        m[10] = dmax(m[10], fabs(r)); m[11] = dmax(m[11], dmax(fabs(px), dmax(fabs(aty), fabs(q))));
        m[12] = dmax(m[12], fabs(Di * atdy));
        m[13] = dmax(m[13], fabs(D * dx)); m[14] += q * dx; m[15] = dmax(m[15], fabs(Di * pdx));
      };
      // a bundle: its row, its singleton variable and that variable's bound row; y_out / pdy_out return the
      // row's multiplier and projected delta for the core column sum (core lanes)
      auto bundle_acc = [&](int k, double core_ax, double core_adx, double& y_out, double& pdy_out) {
        y_out = 0; pdy_out = 0;
        if (!C.active[k]) return;
        const BundleHot& b = L.b[k];
        const double z = proj_row(L, b.v, b.l), y = L.rho_r * (b.v - z);
        const double us = L.row_eq ? b.l : kOsqpInfty * C.E[k];
        const bool lo_inf = is_inf_lo(b.l), hi_inf = is_inf_hi(us);
        row_acc(core_ax + b.e * b.xd, z, C.E[k], C.Einv[k], b.l, us, lo_inf, hi_inf, C.dy[k], core_adx + b.e * C.dxd[k]);
        y_out = y; pdy_out = proj_dy(C.dy[k], lo_inf, hi_inf);
        if (C.has_sing[k]) {
          double aty = b.e * y, atdy = b.e * pdy_out;
          if (C.has_sb[k]) {
            const double zb = proj_sb(L, b.vb), yb = L.rho_b * (b.vb - zb);
            const double lb = L.sb_free ? -kOsqpInfty * C.Eb[k] : 0.0, ub = kOsqpInfty * C.Eb[k];
            const bool blo = is_inf_lo(lb), bhi = is_inf_hi(ub);
            aty += b.beta * yb; atdy += b.beta * proj_dy(C.dyb[k], blo, bhi);
            row_acc(b.beta * b.xd, zb, C.Eb[k], C.Ebinv[k], lb, ub, blo, bhi, C.dyb[k], b.beta * C.dxd[k]);
          }
          var_acc(0.0, aty, b.qd, C.Dd[k], C.Ddinv[k], atdy, C.dxd[k], 0.0);
        }
      };
      double cax, cadx;                          // core part of bundle 0's row:  A_row x_c,  A_row dx_c
      double px = 0, pdx = 0, aty = 0, atdy = 0;  // core variable: (P x)_j, (P dx)_j, (A'y)_j, (A'dy)_j
      if (L.is_core) {
        const int j = L.gl;
#pragma unroll
        for (int i = 0; i < NC; ++i) { px += S.P[j * NC + i] * S.xy[i]; pdx += S.P[j * NC + i] * S.v[i]; }
#pragma unroll
        for (int r = 0; r < NR; ++r) { aty += S.A[r * NC + j] * S.xy[NC + r]; atdy += S.A[r * NC + j] * S.v[NC + r]; }
        if (Cfg::BOUNDS) {
          const double zc = clampd(L.vc, L.lc, L.uc), yc = L.rhoc * (L.vc - zc);
          const bool lo_inf = is_inf_lo(L.lc), hi_inf = is_inf_hi(L.uc);
          aty += L.betac * yc; atdy += L.betac * proj_dy(C.dyc, lo_inf, hi_inf);
          row_acc(L.betac * L.x, zc, C.Ecb, C.Ecbinv, L.lc, L.uc, lo_inf, hi_inf, C.dyc, L.betac * C.dx);
        }
        cax = L.b[0].c * L.x; cadx = L.b[0].c * C.dx;
      } else {
        const int r = L.gl - NC;
        cax = 0; cadx = 0;
#pragma unroll
        for (int i = 0; i < NC; ++i) { cax += S.A[r * NC + i] * S.xy[i]; cadx += S.A[r * NC + i] * S.v[i]; }
      }
      {
        double y, pdy;
        bundle_acc(0, cax, cadx, y, pdy);
        aty += L.b[0].c * y; atdy += L.b[0].c * pdy;
      }
      if (L.is_core) {
#pragma unroll
        for (int k = 1; k < KU; ++k) {
          double y, pdy;
          bundle_acc(k, L.b[k].c * L.x, L.b[k].c * C.dx, y, pdy);
          aty += L.b[k].c * y; atdy += L.b[k].c * pdy;
        }
        var_acc(px, aty, L.q, C.D, C.Dinv, atdy, C.dx, pdx);
      }
#pragma unroll
      for (int i = 0; i < kNumRed; ++i) S.red[i * GL + L.gl] = m[i];
    });
    // C2a: the lanes share the kNumRed group reductions
    w.each([&](LaneT& L, GS& S) {
      if (S.done) return;
      for (int i = L.gl; i < kNumRed; i += GL) {
        const bool is_sum = (i == 5 || i == 14);
        double acc = (i == 6 || i == 7) ? -1e300 : 0.0;
#pragma unroll
        for (int l = 0; l < GL; ++l) {
          const double val = S.red[i * GL + l];
          acc = is_sum ? acc + val : dmax(acc, val);
        }
        S.tot[i] = acc;
      }
    });
    // C2b: lane 0 decides
    w.each([&](LaneT& L, GS& S) {
      if (S.done || L.gl != 0) return;
      double m[kNumRed];
#pragma unroll
      for (int i = 0; i < kNumRed; ++i) m[i] = S.tot[i];
      const double pri_res = m[0], dua_res = S.cinv * m[8];
      S.pri_res = pri_res; S.dua_res = dua_res;
      if (can_check) {
        auto evaluate = [&](double mult) -> int {
          const double ea = o.eps_abs * mult, er = o.eps_rel * mult, epi = o.eps_prim_inf * mult, edi = o.eps_dual_inf * mult;
          if (pri_res > kOsqpInfty || dua_res > kOsqpInfty) return kQpNonConvex;
          bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
          const double eps_prim = ea + er * m[1];
          if (pri_res < eps_prim) prim_ok = true;
          else if (m[4] > epi && m[5] < -epi * m[4]) prim_inf = m[12] < epi * m[4];
          const double eps_dual = ea + er * S.cinv * m[9];
          if (dua_res < eps_dual) dual_ok = true;
          else if (m[13] > edi && m[14] < -S.c * edi * m[13] && m[15] < S.c * edi * m[13])
            dual_inf = !(m[6] > edi * m[13]) && !(m[7] > edi * m[13]);
          if (prim_ok && dual_ok) return mult > 1.0 ? kQpSolvedInaccurate : kQpSolved;
          if (prim_inf) return kQpPrimalInfeasible;
          if (dual_inf) return kQpDualInfeasible;
          return kQpUnsolved;
        };
        int status = evaluate(1.0);
        if (status == kQpUnsolved && last) {
          status = evaluate(10.0);
          if (status == kQpUnsolved) status = kQpMaxIter;
        }
        if (status != kQpUnsolved) { S.status = status; S.iters = iter; S.done = 1; return; }
      }
      if (can_adapt) {
        // compute_rho_estimate on the SCALED residuals
        const double pr = m[2] / (m[3] + 1e-10);
        const double dr = m[10] / (m[11] + 1e-10);
        double rho_new = S.rho * sqrt(pr / (dr + 1e-10));
        rho_new = dmin(dmax(rho_new, kRhoMin), kRhoMax);
        if (rho_new > S.rho * o.adaptive_rho_tolerance || rho_new < S.rho / o.adaptive_rho_tolerance) {
          S.rho_prev = S.rho; S.rho = rho_new; S.rho_updates += 1; S.need_factor = 1;
        }
      }
    });
    w.each([&](LaneT& L, GS& S) { L.done = S.done != 0; });
    all_done = warp_done();
  }
}

}  // namespace drc
