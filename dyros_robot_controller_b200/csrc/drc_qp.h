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
//   every singleton appears in exactly one non-bound row.
// Hence K = P + sigma I + A' diag(rho) A is [K_cc K_cd; K_dc diag] and the OSQP linear system
// reduces EXACTLY to the NC x NC Schur complement S; with U = [I; A_dense] the per-iteration solve is
//   [x~_c ; A_dense x~_c] = (U S^-1 U') u          u = [local right-hand sides ; dense-row terms]
// i.e. one dot product of length GL = NC+ND+NE per lane against a vector exchanged through shared
// memory.  The iterates are those of OSQP (same scaling, rho vector, relaxation, termination,
// infeasibility tests, adaptive rho); only the linear algebra of the KKT solve differs (QDLDL vs
// Schur complement), which changes results at rounding level.
//
// Work mapping: a GROUP of GL lanes owns one robot; a warp carries NG = 32/GL robots.
//   lane j < NC          : core variable j, its bound row, its KU unit rows (+slacks, +slack bounds)
//   lane NC+r (r<ND+NE)  : dense / equality row r (+ singleton, + singleton bound)
// Per-lane state lives in registers; the only cross-lane traffic is shared memory + __syncwarp().
// The kernel body is written against a `Warp` executor so that tests/kernel_emu can run the
// identical code on the CPU (lanes emulated by a loop).
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
  static constexpr int NX = NC + (SLACK ? NC * KU + ND : 0) + NE;  // OSQP n (when every unit row is active)
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

// One constraint row together with the singleton variable it owns (if any) and that variable's
// bound row (if any).  "cp" = core part of the row: a scalar on the lane's own core variable for unit
// rows, a dense NC-vector (held by the lane) for dense/equality rows.
struct Bundle {
  // row
  double E, l, u, z, y, dy;
  int cls;
  // singleton variable
  double Dd, e, qd, xd, dxd;       // e = scaled coefficient of the singleton in the row
  // singleton bound row
  double Eb, beta, lb, ub, zb, yb, dyb;
  int clsb;
  // per-factorisation constants
  double rho, rinv, rhob, rinvb, kinv, gam;
  // scratch carried between phases
  double bd, zt_keep;
  bool active, has_sing, has_sb;
};

template <class Cfg>
struct Lane {
  // group bookkeeping
  int gl;        // lane within the group (0..GL-1), -1 for idle lanes
  int grp;       // group within the warp
  bool is_core;  // gl < NC
  // core variable (core lanes)
  double D, q, x, dx;
  double Ecb, betac, lc, uc, zc, yc, dyc, rhoc, rinvc;  // core bound row
  int clsc;
  double a_unit[Cfg::KU > 0 ? Cfg::KU : 1];             // scaled coefficient of own variable in unit row k
  Bundle ub[Cfg::KU > 0 ? Cfg::KU : 1];
  // dense / equality row (row lanes)
  Bundle rb;
  // operator row  W' = U S^-1 U'
  double w[Cfg::GL];
  double ucore;  // scratch: local right-hand side
};

// Shared scratch of one robot (group).
constexpr int kNumRed = 22;  // group-wide reductions of one termination check
template <class Cfg>
struct GroupShared {
  double P[Cfg::NC * Cfg::NC];                       // scaled P_cc (symmetric, dense)
  double A[(Cfg::NR > 0 ? Cfg::NR : 1) * Cfg::NC];   // scaled dense-row coefficients
  double S[Cfg::NC * Cfg::NC];                       // Schur complement -> its inverse
  double T[(Cfg::NR > 0 ? Cfg::NR : 1) * Cfg::NC];   // S^-1 a_r
  double u[Cfg::GL];                                 // exchange vector (x_c | row terms)
  double v[Cfg::GL];                                 // second exchange vector
  double red[Cfg::GL * kNumRed];                     // per-lane partial reductions
  double c, cinv, rho;
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
// bundle helpers
// ------------------------------------------------------------------------------------------------
DRC_HD void bundle_init(Bundle& b, bool active, double l, double u, bool has_sing, double e, double qd, bool has_sb,
                        double lb, double ub) {
  b.active = active; b.has_sing = active && has_sing; b.has_sb = b.has_sing && has_sb;
  b.E = 1.0; b.l = l; b.u = u; b.z = 0; b.y = 0; b.dy = 0; b.cls = 0;
  b.Dd = 1.0; b.e = b.has_sing ? e : 0.0; b.qd = b.has_sing ? qd : 0.0; b.xd = 0; b.dxd = 0;
  b.Eb = 1.0; b.beta = b.has_sb ? 1.0 : 0.0; b.lb = lb; b.ub = ub; b.zb = 0; b.yb = 0; b.dyb = 0; b.clsb = 0;
  b.rho = b.rinv = b.rhob = b.rinvb = b.kinv = b.gam = 0; b.bd = 0; b.zt_keep = 0;
}
// finish scaling: scaled bounds + OSQP constraint classes
DRC_HD void bundle_finalize(Bundle& b) {
  if (!b.active) return;
  b.l *= b.E; b.u *= b.E;
  b.cls = row_class(b.l, b.u);
  if (b.has_sb) { b.lb *= b.Eb; b.ub *= b.Eb; b.clsb = row_class(b.lb, b.ub); }
}
// per-factorisation constants; returns omega = effective rho of the row after eliminating the singleton
DRC_HD double bundle_factor(Bundle& b, double rho, double sigma) {
  if (!b.active) return 0.0;
  b.rho = class_rho(b.cls, rho); b.rinv = 1.0 / b.rho;
  if (!b.has_sing) { b.kinv = 0; b.gam = 0; return b.rho; }
  double kap = sigma + b.rho * b.e * b.e;
  if (b.has_sb) { b.rhob = class_rho(b.clsb, rho); b.rinvb = 1.0 / b.rhob; kap += b.rhob * b.beta * b.beta; }
  b.kinv = 1.0 / kap;
  b.gam = b.rho * b.e * b.kinv;
  return b.rho * (1.0 - b.gam * b.e);
}
// first half of an iteration: returns t (the row's contribution weight to the core right-hand side)
DRC_HD double bundle_pre(Bundle& b, double sigma) {
  if (!b.active) return 0.0;
  const double wr = b.rho * b.z - b.y;
  if (!b.has_sing) return wr;
  double bd = sigma * b.xd - b.qd + b.e * wr;
  if (b.has_sb) bd += b.beta * (b.rhob * b.zb - b.yb);
  b.bd = bd;
  return wr - b.gam * bd;
}
DRC_HD double proj(double v, double l, double u) { return dmin(dmax(v, l), u); }
// second half: s = core part of the row applied to x~_c
DRC_HD void bundle_post(Bundle& b, double s, double alpha, bool keep_delta) {
  if (!b.active) return;
  double zt = s;
  if (b.has_sing) {
    const double xtd = b.kinv * b.bd - b.gam * s;
    zt += b.e * xtd;
    const double xn = alpha * xtd + (1.0 - alpha) * b.xd;
    if (keep_delta) b.dxd = xn - b.xd;
    b.xd = xn;
    if (b.has_sb) {
      const double zr = alpha * (b.beta * xtd) + (1.0 - alpha) * b.zb;
      const double zn = proj(zr + b.yb * b.rinvb, b.lb, b.ub);
      const double d = b.rhob * (zr - zn);
      if (keep_delta) b.dyb = d;
      b.yb += d; b.zb = zn;
    }
  }
  const double zr = alpha * zt + (1.0 - alpha) * b.z;
  const double zn = proj(zr + b.y * b.rinv, b.l, b.u);
  const double d = b.rho * (zr - zn);
  if (keep_delta) b.dy = d;
  b.y += d; b.z = zn;
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


// ------------------------------------------------------------------------------------------------
// The solver.  `qp` points at the per-robot records (Cfg::STRIDE doubles each); robots[g] is the
// robot handled by group g of this warp (-1 = none).  On return the (scaled) iterates are still in
// the lanes; the caller unscales what it needs:  x_c[j] = D * x   (core lane j),
// singleton = Dd * xd (bundle), and reads status / iters from the group's shared record.
// ------------------------------------------------------------------------------------------------
template <class Cfg, class W>
DRC_HD void admm_solve(W& w, const double* qp, const int* robots, const QpOptions& o) {
  constexpr int NC = Cfg::NC, KU = Cfg::KU, ND = Cfg::ND, NR = Cfg::NR, GL = Cfg::GL, NG = Cfg::NG;
  typedef Lane<Cfg> LaneT;
  typedef GroupShared<Cfg> GS;
  const double sigma = o.sigma, alpha = o.alpha;

  // ---------------------------------------------------------------- load
  w.each([&](LaneT& L, GS& S) {
    const int rb = robots[L.grp];
    if (L.gl == 0) {
      S.robot = rb; S.status = kQpUnsolved; S.iters = 0; S.done = rb < 0 ? 1 : 0; S.c = 1.0; S.rho_updates = 0;
      S.need_factor = 0; S.pri_res = 0; S.dua_res = 0;
      int na = 0;
      for (int j = 0; j < NC; ++j) na += (int)((o.unit_mask >> j) & 1u);
      S.nx = NC + (Cfg::SLACK ? na * KU + ND : 0) + Cfg::NE;
    }
    L.dx = 0; L.dyc = 0; L.x = 0; L.zc = 0; L.yc = 0; L.ucore = 0; L.q = 0; L.D = 1; L.Ecb = 1; L.betac = 0;
    L.lc = -kOsqpInfty; L.uc = kOsqpInfty; L.rhoc = 0; L.rinvc = 0; L.clsc = -1;
#pragma unroll
    for (int k = 0; k < KU; ++k) { L.a_unit[k] = 0; bundle_init(L.ub[k], false, 0, 0, false, 0, 0, false, 0, 0); }
    bundle_init(L.rb, false, 0, 0, false, 0, 0, false, 0, 0);
    if (rb < 0) return;
    const double* rec = qp + (long long)rb * Cfg::STRIDE;
    if (L.is_core) {
      const int j = L.gl;
#pragma unroll
      for (int i = 0; i < NC; ++i) S.P[j * NC + i] = rec[Cfg::OFF_P + symidx<NC>(i, j)];
      L.q = rec[Cfg::OFF_Q + j];
      if (Cfg::BOUNDS) { L.betac = 1.0; L.lc = rec[Cfg::OFF_LO + j]; L.uc = rec[Cfg::OFF_HI + j]; }
      const bool act = ((o.unit_mask >> j) & 1u) != 0u;
#pragma unroll
      for (int k = 0; k < KU; ++k) {
        L.a_unit[k] = act ? ((k & 1) ? -1.0 : 1.0) : 0.0;
        bundle_init(L.ub[k], act, rec[Cfg::OFF_UNIT + k * NC + j], kOsqpInfty, Cfg::SLACK, 1.0, o.slack_weight,
                    Cfg::BOUNDS, 0.0, kOsqpInfty);
      }
    } else {
      const int r = L.gl - NC;
      const double* row = rec + Cfg::OFF_ROW + r * (NC + 1);
#pragma unroll
      for (int i = 0; i < NC; ++i) S.A[r * NC + i] = row[i];
      const double l = row[NC];
      if (r >= ND) bundle_init(L.rb, true, l, l, true, -1.0, 0.0, Cfg::BOUNDS, -kOsqpInfty, kOsqpInfty);
      else bundle_init(L.rb, true, l, kOsqpInfty, Cfg::SLACK, 1.0, o.slack_weight, Cfg::BOUNDS, 0.0, kOsqpInfty);
    }
  });

  // ---------------------------------------------------------------- Ruiz equilibration (OSQP scale_data)
  for (int it = 0; it < o.scaling; ++it) {
    // (1) D_temp of the core columns: inf-norm over P, the bound row, unit rows and dense rows
    w.each([&](LaneT& L, GS& S) {
      if (S.done || !L.is_core) return;
      const int j = L.gl;
      double cn = fabs(L.betac);
#pragma unroll
      for (int i = 0; i < NC; ++i) cn = dmax(cn, fabs(S.P[j * NC + i]));
#pragma unroll
      for (int r = 0; r < NR; ++r) cn = dmax(cn, fabs(S.A[r * NC + j]));
#pragma unroll
      for (int k = 0; k < KU; ++k) cn = dmax(cn, fabs(L.a_unit[k]));
      S.u[j] = 1.0 / sqrt(limit_scaling(cn));
    });
    // (2) E_temp of every row, D_temp of the singleton columns; apply to everything lane-local
    w.each([&](LaneT& L, GS& S) {
      if (S.done) return;
      auto scale_bundle = [&](Bundle& b, double core_norm) -> double {
        if (!b.active) return 1.0;
        double dt = 1.0, eb = 1.0;
        if (b.has_sing) {
          double cn = fabs(b.e);
          if (b.has_sb) { cn = dmax(cn, fabs(b.beta)); eb = 1.0 / sqrt(limit_scaling(fabs(b.beta))); }
          dt = 1.0 / sqrt(limit_scaling(cn));
        }
        const double et = 1.0 / sqrt(limit_scaling(dmax(core_norm, fabs(b.e))));
        b.E *= et;
        if (b.has_sing) {
          b.e *= et * dt; b.Dd *= dt; b.qd *= dt;
          if (b.has_sb) { b.beta *= eb * dt; b.Eb *= eb; }
        }
        return et;
      };
      if (L.is_core) {
        const double dj = S.u[L.gl];
        if (Cfg::BOUNDS) { const double eb = 1.0 / sqrt(limit_scaling(fabs(L.betac))); L.betac *= eb * dj; L.Ecb *= eb; }
#pragma unroll
        for (int k = 0; k < KU; ++k) {
          const double et = scale_bundle(L.ub[k], fabs(L.a_unit[k]));
          L.a_unit[k] *= et * dj;
        }
        L.q *= dj; L.D *= dj;
      } else {
        const int r = L.gl - NC;
        double rn = 0;
#pragma unroll
        for (int i = 0; i < NC; ++i) rn = dmax(rn, fabs(S.A[r * NC + i]));
        L.ucore = scale_bundle(L.rb, rn);  // E_temp of the dense row, applied to S.A in (3)
      }
    });
    // (3) scale P and the dense rows; publish the inputs of the cost normalisation
    w.each([&](LaneT& L, GS& S) {
      if (S.done) return;
      double cn = 0, qm = 0;
      if (L.is_core) {
        const int j = L.gl;
        const double dj = S.u[j];
#pragma unroll
        for (int i = 0; i < NC; ++i) { const double pv = S.P[j * NC + i] * dj * S.u[i]; S.P[j * NC + i] = pv; cn = dmax(cn, fabs(pv)); }
        qm = fabs(L.q);
#pragma unroll
        for (int k = 0; k < KU; ++k) if (L.ub[k].has_sing) qm = dmax(qm, fabs(L.ub[k].qd));
      } else {
        const int r = L.gl - NC;
#pragma unroll
        for (int i = 0; i < NC; ++i) S.A[r * NC + i] *= L.ucore * S.u[i];
        if (L.rb.has_sing) qm = fabs(L.rb.qd);
      }
      S.red[L.gl] = cn; S.red[GL + L.gl] = qm;
    });
    // (4) c_temp = 1 / limit(max(mean_j ||P_j||_inf, limit(||q||_inf)))
    w.each([&](LaneT& L, GS& S) {
      if (S.done) return;
      double sum = 0, qm = 0;
#pragma unroll
      for (int i = 0; i < GL; ++i) { sum += S.red[i]; qm = dmax(qm, S.red[GL + i]); }
      const double ct = 1.0 / limit_scaling(dmax(sum / (double)S.nx, limit_scaling(qm)));
      if (L.is_core) {
        const int j = L.gl;
        L.q *= ct;
#pragma unroll
        for (int k = 0; k < KU; ++k) L.ub[k].qd *= ct;
#pragma unroll
        for (int i = 0; i < NC; ++i) S.P[j * NC + i] *= ct;
      } else {
        L.rb.qd *= ct;
      }
      if (L.gl == 0) S.c *= ct;
    });
  }
  // scaled bounds, constraint classes
  w.each([&](LaneT& L, GS& S) {
    if (S.done) return;
    if (L.is_core) {
      if (Cfg::BOUNDS) { L.lc *= L.Ecb; L.uc *= L.Ecb; L.clsc = row_class(L.lc, L.uc); }
#pragma unroll
      for (int k = 0; k < KU; ++k) bundle_finalize(L.ub[k]);
    } else {
      bundle_finalize(L.rb);
    }
    if (L.gl == 0) { S.cinv = 1.0 / S.c; S.rho = dmin(dmax(o.rho, kRhoMin), kRhoMax); S.need_factor = 1; }
  });

  // ---------------------------------------------------------------- factorisation (setup and rho updates)
  auto factor = [&]() {
    // (a) Schur complement S = P + sigma I + diag(bound, unit rows) + sum_r omega_r a_r a_r'
    w.each([&](LaneT& L, GS& S) {
      if (!S.need_factor) return;
      const double rho = S.rho;
      if (L.is_core) {
        const int j = L.gl;
        double diag = sigma;
        if (Cfg::BOUNDS) { L.rhoc = class_rho(L.clsc, rho); L.rinvc = 1.0 / L.rhoc; diag += L.rhoc * L.betac * L.betac; }
#pragma unroll
        for (int k = 0; k < KU; ++k) diag += bundle_factor(L.ub[k], rho, sigma) * L.a_unit[k] * L.a_unit[k];
#pragma unroll
        for (int i = 0; i < NC; ++i) S.S[j * NC + i] = S.P[j * NC + i] + (i == j ? diag : 0.0);
      } else {
        S.v[L.gl] = bundle_factor(L.rb, rho, sigma);
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
    // (b) in-place Gauss-Jordan inverse (S is SPD: no pivoting)
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
    w.each([&](LaneT& L, GS& S) { if (L.gl == 0) S.need_factor = 0; });
  };
  factor();

  auto warp_done = [&]() {
    bool d = true;
    for (int g = 0; g < NG; ++g) d = d && (w.group(g).done != 0);
    return d;
  };

  // ---------------------------------------------------------------- ADMM iterations (osqp_solve)
  int iter = 0;
  bool all_done = warp_done();
  while (!all_done && iter < o.max_iter) {
    ++iter;
    const bool last = iter == o.max_iter;
    const bool can_check = (o.check_termination > 0 && (iter % o.check_termination == 0)) || last;
    const bool can_adapt = o.adaptive_rho && o.adaptive_rho_interval > 0 && (iter % o.adaptive_rho_interval == 0);
    // phase A: local right-hand sides  u = [sigma x - q + A_local'(rho z - y) ; t_r]
    w.each([&](LaneT& L, GS& S) {
      if (S.done) return;
      if (L.is_core) {
        double uc = sigma * L.x - L.q;
        if (Cfg::BOUNDS) uc += L.betac * (L.rhoc * L.zc - L.yc);
#pragma unroll
        for (int k = 0; k < KU; ++k) uc += L.a_unit[k] * bundle_pre(L.ub[k], sigma);
        S.u[L.gl] = uc;
      } else {
        S.u[L.gl] = bundle_pre(L.rb, sigma);
      }
    });
    // phase B: [x~_c ; A_dense x~_c] = (U S^-1 U') u, then the x / z / y updates
    w.each([&](LaneT& L, GS& S) {
      if (S.done) return;
      double s = 0;
#pragma unroll
      for (int i = 0; i < GL; ++i) s += L.w[i] * S.u[i];
      if (L.is_core) {
#pragma unroll
        for (int k = 0; k < KU; ++k) bundle_post(L.ub[k], L.a_unit[k] * s, alpha, can_check);
        const double xn = alpha * s + (1.0 - alpha) * L.x;
        if (can_check) L.dx = xn - L.x;
        L.x = xn;
        if (Cfg::BOUNDS) {
          const double zr = alpha * (L.betac * s) + (1.0 - alpha) * L.zc;
          const double zn = proj(zr + L.yc * L.rinvc, L.lc, L.uc);
          const double d = L.rhoc * (zr - zn);
          if (can_check) L.dyc = d;
          L.yc += d; L.zc = zn;
        }
      } else {
        bundle_post(L.rb, s, alpha, can_check);
      }
    });
    if (!(can_check || can_adapt)) continue;

    // ---------------- OSQP update_info + check_termination + adapt_rho
    // project delta_y onto the polar of the recession cone of [l,u] (is_primal_infeasible)
    auto proj_dy = [](double dy, double l, double u) {
      if (is_inf_hi(u)) return is_inf_lo(l) ? 0.0 : dmin(dy, 0.0);
      if (is_inf_lo(l)) return dmax(dy, 0.0);
      return dy;
    };
    // C0: exchange x_c | y_r and dx_c | projected dy_r
    w.each([&](LaneT& L, GS& S) {
      if (S.done) return;
      if (L.is_core) { S.u[L.gl] = L.x; S.v[L.gl] = L.dx; }
      else { S.u[L.gl] = L.rb.y; S.v[L.gl] = proj_dy(L.rb.dy, L.rb.l, L.rb.u); }
    });
    // C1: per-lane partial reductions
    w.each([&](LaneT& L, GS& S) {
      if (S.done) return;
      double m[kNumRed];
#pragma unroll
      for (int i = 0; i < kNumRed; ++i) m[i] = 0.0;
      m[20] = -1e300; m[21] = 1e300;
      // rows:      0 pri_u 1 ax_u 2 z_u | 7 pri_s 8 ax_s 9 z_s | 14 ||E dy|| 15 sum(u dy+ + l dy-) | 20 max Adx (finite u) 21 min Adx (finite l)
      // variables: 3 dua_u 4 px_u 5 aty_u 6 q_u | 10 dua_s 11 px_s 12 aty_s 13 q_s | 16 ||Dinv A'dy|| | 17 ||D dx|| 18 q'dx 19 ||Dinv P dx||
      auto row_acc = [&](double ax, double z, double E, double l, double u, double dy, double adx) {
        const double Ei = 1.0 / E, r = ax - z;
        m[0] = dmax(m[0], fabs(Ei * r)); m[1] = dmax(m[1], fabs(Ei * ax)); m[2] = dmax(m[2], fabs(Ei * z));
        m[7] = dmax(m[7], fabs(r)); m[8] = dmax(m[8], fabs(ax)); m[9] = dmax(m[9], fabs(z));
        const double pdy = proj_dy(dy, l, u);
        m[14] = dmax(m[14], fabs(E * pdy));
        m[15] += u * dmax(pdy, 0.0) + l * dmin(pdy, 0.0);
        if (!is_inf_hi(u)) m[20] = dmax(m[20], Ei * adx);
        if (!is_inf_lo(l)) m[21] = dmin(m[21], Ei * adx);
      };
      auto var_acc = [&](double px, double aty, double q, double D, double atdy, double dx, double pdx) {
        const double Di = 1.0 / D, r = px + q + aty;
        m[3] = dmax(m[3], fabs(Di * r)); m[4] = dmax(m[4], fabs(Di * px)); m[5] = dmax(m[5], fabs(Di * aty)); m[6] = dmax(m[6], fabs(Di * q));
        m[10] = dmax(m[10], fabs(r)); m[11] = dmax(m[11], fabs(px)); m[12] = dmax(m[12], fabs(aty)); m[13] = dmax(m[13], fabs(q));
        m[16] = dmax(m[16], fabs(Di * atdy));
        m[17] = dmax(m[17], fabs(D * dx)); m[18] += q * dx; m[19] = dmax(m[19], fabs(Di * pdx));
      };
      // a bundle: its row, its singleton variable and that variable's bound row
      auto bundle_acc = [&](const Bundle& b, double core_ax, double core_adx) {
        if (!b.active) return;
        row_acc(core_ax + (b.has_sing ? b.e * b.xd : 0.0), b.z, b.E, b.l, b.u, b.dy, core_adx + (b.has_sing ? b.e * b.dxd : 0.0));
        if (b.has_sing) {
          double aty = b.e * b.y, atdy = b.e * proj_dy(b.dy, b.l, b.u);
          if (b.has_sb) {
            aty += b.beta * b.yb; atdy += b.beta * proj_dy(b.dyb, b.lb, b.ub);
            row_acc(b.beta * b.xd, b.zb, b.Eb, b.lb, b.ub, b.dyb, b.beta * b.dxd);
          }
          var_acc(0.0, aty, b.qd, b.Dd, atdy, b.dxd, 0.0);
        }
      };
      if (L.is_core) {
        const int j = L.gl;
        double px = 0, pdx = 0, aty = 0, atdy = 0;
#pragma unroll
        for (int i = 0; i < NC; ++i) { px += S.P[j * NC + i] * S.u[i]; pdx += S.P[j * NC + i] * S.v[i]; }
#pragma unroll
        for (int r = 0; r < NR; ++r) { aty += S.A[r * NC + j] * S.u[NC + r]; atdy += S.A[r * NC + j] * S.v[NC + r]; }
        if (Cfg::BOUNDS) {
          aty += L.betac * L.yc; atdy += L.betac * proj_dy(L.dyc, L.lc, L.uc);
          row_acc(L.betac * L.x, L.zc, L.Ecb, L.lc, L.uc, L.dyc, L.betac * L.dx);
        }
#pragma unroll
        for (int k = 0; k < KU; ++k) {
          if (L.ub[k].active) { aty += L.a_unit[k] * L.ub[k].y; atdy += L.a_unit[k] * proj_dy(L.ub[k].dy, L.ub[k].l, L.ub[k].u); }
          bundle_acc(L.ub[k], L.a_unit[k] * L.x, L.a_unit[k] * L.dx);
        }
        var_acc(px, aty, L.q, L.D, atdy, L.dx, pdx);
      } else {
        const int r = L.gl - NC;
        double ax = 0, adx = 0;
#pragma unroll
        for (int i = 0; i < NC; ++i) { ax += S.A[r * NC + i] * S.u[i]; adx += S.A[r * NC + i] * S.v[i]; }
        bundle_acc(L.rb, ax, adx);
      }
#pragma unroll
      for (int i = 0; i < kNumRed; ++i) S.red[i * GL + L.gl] = m[i];
    });
    // C2: lane 0 reduces and decides
    w.each([&](LaneT& L, GS& S) {
      if (S.done || L.gl != 0) return;
      double m[kNumRed];
#pragma unroll
      for (int i = 0; i < kNumRed; ++i) {
        const bool is_sum = (i == 15 || i == 18), is_min = (i == 21);
        double acc = is_sum ? 0.0 : (is_min ? 1e300 : (i == 20 ? -1e300 : 0.0));
        for (int l = 0; l < GL; ++l) {
          const double val = S.red[i * GL + l];
          acc = is_sum ? acc + val : (is_min ? dmin(acc, val) : dmax(acc, val));
        }
        m[i] = acc;
      }
      const double pri_res = m[0], dua_res = S.cinv * m[3];
      S.pri_res = pri_res; S.dua_res = dua_res;
      if (can_check) {
        int status = kQpUnsolved;
        auto evaluate = [&](double mult) -> int {
          const double ea = o.eps_abs * mult, er = o.eps_rel * mult, epi = o.eps_prim_inf * mult, edi = o.eps_dual_inf * mult;
          if (pri_res > kOsqpInfty || dua_res > kOsqpInfty) return kQpNonConvex;
          bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
          const double eps_prim = ea + er * dmax(m[2], m[1]);
          if (pri_res < eps_prim) prim_ok = true;
          else if (m[14] > epi && m[15] < -epi * m[14]) prim_inf = m[16] < epi * m[14];
          const double eps_dual = ea + er * S.cinv * dmax(m[6], dmax(m[5], m[4]));
          if (dua_res < eps_dual) dual_ok = true;
          else if (m[17] > edi && m[18] < -S.c * edi * m[17] && m[19] < S.c * edi * m[17])
            dual_inf = !(m[20] > edi * m[17]) && !(m[21] < -edi * m[17]);
          if (prim_ok && dual_ok) return mult > 1.0 ? kQpSolvedInaccurate : kQpSolved;
          if (prim_inf) return kQpPrimalInfeasible;
          if (dual_inf) return kQpDualInfeasible;
          return kQpUnsolved;
        };
        status = evaluate(1.0);
        if (status == kQpUnsolved && last) {
          status = evaluate(10.0);
          if (status == kQpUnsolved) status = kQpMaxIter;
        }
        if (status != kQpUnsolved) { S.status = status; S.iters = iter; S.done = 1; return; }
      }
      if (can_adapt) {
        // compute_rho_estimate on the SCALED residuals
        const double pr = m[7] / (dmax(m[9], m[8]) + 1e-10);
        const double dr = m[10] / (dmax(m[13], dmax(m[12], m[11])) + 1e-10);
        double rho_new = S.rho * sqrt(pr / (dr + 1e-10));
        rho_new = dmin(dmax(rho_new, kRhoMin), kRhoMax);
        if (rho_new > S.rho * o.adaptive_rho_tolerance || rho_new < S.rho / o.adaptive_rho_tolerance) {
          S.rho = rho_new; S.rho_updates += 1; S.need_factor = 1;
        }
      }
    });
    {
      bool nf = false;
      for (int g = 0; g < NG; ++g) nf = nf || (w.group(g).need_factor != 0 && w.group(g).done == 0);
      if (nf) factor();
    }
    all_done = warp_done();
  }
}

}  // namespace drc
