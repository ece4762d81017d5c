// drc_b200 -- per-thread narrow phase for primitive self-collision pairs (fp64).
//
// B200-native replacement of pinocchio::computeDistances -> hpp-fcl distance() as used by the
// reference's getMinDistance (src/manipulator/robot_data.cpp:424-443).  All work is done in the
// frame of the first shape's parent joint: shapes of link A are model constants, shapes of link B
// are moved by the relative transform T_AB, so the pair loop touches no per-robot arrays.
//   * sphere-X pairs: closed form (exact).
//   * cylinder-cylinder / cylinder-box: certified LOWER BOUNDS first (capsule bound and a
//     support-function bound); only pairs whose bound beats the best exact distance found so far
//     run GJK (separated) and, if they overlap, EPA -- the minimum over all pairs is unchanged.
#pragma once
#include "drc_math.h"

namespace drc {

// A primitive placed in the working frame.
struct Prim {
  int type;
  double r, h;   // sphere: r | cylinder/capsule: r, half length | box: unused (see hb)
  Vec3 hb;       // box half extents
  Vec3 c;        // centre
  Vec3 a;        // cylinder / capsule axis (unit); box: unused
  Mat3 R;        // box / convex orientation (only valid for those)
  const double* verts;  // convex: hull vertices in the shape's frame (about c), nvert of them; r = bounding radius
  int nvert;
};

struct PairResult {
  double d;
  Vec3 pa, pb;
};

DRC_HD Vec3 any_perp(Vec3 a) {  // some unit vector orthogonal to the unit vector a
  Vec3 t = fabs(a.x) < 0.9 ? v3(1, 0, 0) : v3(0, 1, 0);
  Vec3 p = cross(a, t);
  return (1.0 / norm(p)) * p;
}

DRC_HD PairResult sphere_sphere(Vec3 c1, double r1, Vec3 c2, double r2) {
  const Vec3 diff = c2 - c1;
  const double len = norm(diff);
  const Vec3 n = len > 0 ? (1.0 / len) * diff : v3(1, 0, 0);
  return PairResult{len - r1 - r2, c1 + r1 * n, c2 - r2 * n};
}

// sphere (A) against a solid whose closest surface point q and signed centre distance sd are known
DRC_HD PairResult sphere_vs_surface(Vec3 cs, double rs, Vec3 q, double sd) {
  const Vec3 dir = q - cs;
  const double len = norm(dir);
  const Vec3 n = len > 0 ? (1.0 / len) * dir : v3(1, 0, 0);
  PairResult r;
  r.d = sd - rs;
  r.pb = q;
  r.pa = sd >= 0 ? cs + rs * n : cs - rs * n;
  return r;
}

DRC_HD PairResult sphere_cylinder(Vec3 cs, double rs, const Prim& cy) {
  const Vec3 x = cs - cy.c;
  const double z = dot(x, cy.a);
  const Vec3 rad = x - z * cy.a;
  const double rho = norm(rad);
  Vec3 q;
  double sd;
  if (fabs(z) <= cy.h && rho <= cy.r) {
    const double dr = cy.r - rho, dz = cy.h - fabs(z);
    if (dr < dz) {
      const Vec3 u = rho > 0 ? (1.0 / rho) * rad : any_perp(cy.a);
      q = cy.c + z * cy.a + cy.r * u;
      sd = -dr;
    } else {
      q = cy.c + (z >= 0 ? cy.h : -cy.h) * cy.a + rad;
      sd = -dz;
    }
  } else {
    const double s = rho > cy.r ? cy.r / rho : 1.0;
    const double zc = clampd(z, -cy.h, cy.h);
    q = cy.c + zc * cy.a + s * rad;
    sd = norm(cs - q);
  }
  return sphere_vs_surface(cs, rs, q, sd);
}

DRC_HD PairResult sphere_box(Vec3 cs, double rs, const Prim& bx) {
  const Vec3 x = tmul(bx.R, cs - bx.c);
  const double ax = fabs(x.x), ay = fabs(x.y), az = fabs(x.z);
  Vec3 ql;
  double sd;
  if (ax <= bx.hb.x && ay <= bx.hb.y && az <= bx.hb.z) {
    const double dx = bx.hb.x - ax, dy = bx.hb.y - ay, dz = bx.hb.z - az;
    ql = x;
    if (dx <= dy && dx <= dz) { ql.x = x.x >= 0 ? bx.hb.x : -bx.hb.x; sd = -dx; }
    else if (dy <= dz) { ql.y = x.y >= 0 ? bx.hb.y : -bx.hb.y; sd = -dy; }
    else { ql.z = x.z >= 0 ? bx.hb.z : -bx.hb.z; sd = -dz; }
  } else {
    ql = v3(clampd(x.x, -bx.hb.x, bx.hb.x), clampd(x.y, -bx.hb.y, bx.hb.y), clampd(x.z, -bx.hb.z, bx.hb.z));
    sd = norm(x - ql);
  }
  return sphere_vs_surface(cs, rs, mul(bx.R, ql) + bx.c, sd);
}

// closest points of two segments p1 + s d1, p2 + t d2, s,t in [0,1]
DRC_HD void segment_segment(Vec3 p1, Vec3 d1, Vec3 p2, Vec3 d2, double& s, double& t) {
  const Vec3 r = p1 - p2;
  const double a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r);
  const double eps = 1e-300;
  if (a <= eps && e <= eps) { s = t = 0; return; }
  if (a <= eps) { s = 0; t = clampd(f / e, 0.0, 1.0); return; }
  const double c = dot(d1, r);
  if (e <= eps) { t = 0; s = clampd(-c / a, 0.0, 1.0); return; }
  const double b = dot(d1, d2), den = a * e - b * b;
  s = den > 1e-14 * a * e ? clampd((b * f - c * e) / den, 0.0, 1.0) : 0.0;
  t = (b * s + f) / e;
  if (t < 0) { t = 0; s = clampd(-c / a, 0.0, 1.0); }
  else if (t > 1) { t = 1; s = clampd((b - c) / a, 0.0, 1.0); }
}

DRC_HD PairResult capsule_capsule(const Prim& A, const Prim& B) {
  double s, t;
  segment_segment(A.c - A.h * A.a, (2 * A.h) * A.a, B.c - B.h * B.a, (2 * B.h) * B.a, s, t);
  return sphere_sphere(A.c + ((2 * s - 1) * A.h) * A.a, A.r, B.c + ((2 * t - 1) * B.h) * B.a, B.r);
}
DRC_HD PairResult sphere_capsule(Vec3 cs, double rs, const Prim& cp) {
  const double t = clampd(dot(cs - cp.c, cp.a), -cp.h, cp.h);
  return sphere_sphere(cs, rs, cp.c + t * cp.a, cp.r);
}

// ---- support mapping and support width (about the centre) of a primitive
DRC_HD Vec3 support(const Prim& S, Vec3 d) {
  switch (S.type) {
    case kSphere: {
      const double n = norm(d);
      return n > 0 ? S.c + (S.r / n) * d : S.c;
    }
    case kBox: {
      const Vec3 dl = tmul(S.R, d);
      return mul(S.R, v3(dl.x >= 0 ? S.hb.x : -S.hb.x, dl.y >= 0 ? S.hb.y : -S.hb.y, dl.z >= 0 ? S.hb.z : -S.hb.z)) + S.c;
    }
    case kConvex: {  // convex hull of a mesh: the vertex furthest along d (brute force over the hull's vertex set)
      const Vec3 dl = tmul(S.R, d);
      int bi = 0;
      double bv = -1e300;
      for (int i = 0; i < S.nvert; ++i) {
        const double t = dl.x * S.verts[3 * i] + dl.y * S.verts[3 * i + 1] + dl.z * S.verts[3 * i + 2];
        if (t > bv) { bv = t; bi = i; }
      }
      return mul(S.R, v3(S.verts[3 * bi], S.verts[3 * bi + 1], S.verts[3 * bi + 2])) + S.c;
    }
    case kCylinder: {
      // radial part of d, orthogonalised twice: for d (nearly) along the axis the first difference is rounding noise
      // with an axial component, which would push the rim point out of the cylinder by up to r
      const double da = dot(d, S.a);
      Vec3 perp = d - da * S.a;
      perp = perp - dot(perp, S.a) * S.a;
      const double sg2 = dot(perp, perp);
      Vec3 s = S.c + (da >= 0 ? S.h : -S.h) * S.a;
      if (sg2 > 1e-28 * dot(d, d)) s = s + (S.r / sqrt(sg2)) * perp;
      return s;
    }
    default: {  // capsule
      const double n = norm(d);
      Vec3 s = S.c + (dot(d, S.a) >= 0 ? S.h : -S.h) * S.a;
      if (n > 0) s = s + (S.r / n) * d;
      return s;
    }
  }
}
DRC_HD double support_width(const Prim& S, Vec3 n) {  // max over the shape of n.(x - c), |n| = 1
  switch (S.type) {
    case kSphere: return S.r;
    case kConvex: return S.r;  // bounding radius: an upper bound of the width keeps the pair bound a certified lower bound
    case kBox: {
      const Vec3 dl = tmul(S.R, n);
      return S.hb.x * fabs(dl.x) + S.hb.y * fabs(dl.y) + S.hb.z * fabs(dl.z);
    }
    case kCylinder: {
      const double da = dot(n, S.a);
      return S.h * fabs(da) + S.r * sqrt(dmax(0.0, 1.0 - da * da));
    }
    default: return S.h * fabs(dot(n, S.a)) + S.r;
  }
}

// Certified lower bound of the signed distance between two convex primitives, at least one of
// which is a cylinder (the other a cylinder, capsule or box).
DRC_HD double pair_lower_bound(const Prim& A, const Prim& B) {
  // inner segments (a box / a convex hull / a sphere is reduced to its centre with its circumscribed radius)
  const bool blobA = A.type == kBox || A.type == kConvex || A.type == kSphere, blobB = B.type == kBox || B.type == kConvex || B.type == kSphere;
  const double hA = blobA ? 0.0 : A.h, hB = blobB ? 0.0 : B.h;
  const double rA = A.type == kBox ? norm(A.hb) : A.r, rB = B.type == kBox ? norm(B.hb) : B.r;
  const Vec3 aA = blobA ? v3(0, 0, 1) : A.a, aB = blobB ? v3(0, 0, 1) : B.a;
  double s, t;
  segment_segment(A.c - hA * aA, (2 * hA) * aA, B.c - hB * aB, (2 * hB) * aB, s, t);
  const Vec3 pA = A.c + ((2 * s - 1) * hA) * aA, pB = B.c + ((2 * t - 1) * hB) * aB;
  const Vec3 dv = pB - pA;
  const double dist = norm(dv);
  double lb = dist - rA - rB;  // swept-sphere (capsule) bound
  if (dist > 1e-12) {
    const Vec3 n = (1.0 / dist) * dv;
    lb = dmax(lb, dot(n, B.c - A.c) - support_width(A, n) - support_width(B, -n));
  }
  const Vec3 cc = B.c - A.c;
  const double cl = norm(cc);
  if (cl > 1e-12) {
    const Vec3 n = (1.0 / cl) * cc;
    lb = dmax(lb, cl - support_width(A, n) - support_width(B, -n));
  }
  return lb;
}

// ---------------------------------------------------------------- GJK
struct SimplexVert {
  Vec3 w, a, b;
};

DRC_HD void closest_segment(SimplexVert* v, int& n, double* lam) {
  const Vec3 a = v[0].w, ab = v[1].w - v[0].w;
  const double t = -dot(a, ab), den = dot(ab, ab);
  if (t <= 0 || den <= 0) { n = 1; lam[0] = 1; return; }
  if (t >= den) { v[0] = v[1]; n = 1; lam[0] = 1; return; }
  lam[1] = t / den; lam[0] = 1 - lam[1];
}
// closest point to the origin on the triangle (a, b, c): which vertices support it (idx, in input order) and their weights
DRC_HD int closest_triangle_w(Vec3 a, Vec3 b, Vec3 c, int* idx, double* lam) {
  const Vec3 ab = b - a, ac = c - a;
  const double d1 = -dot(ab, a), d2 = -dot(ac, a);
  if (d1 <= 0 && d2 <= 0) { idx[0] = 0; lam[0] = 1; return 1; }
  const double d3 = -dot(ab, b), d4 = -dot(ac, b);
  if (d3 >= 0 && d4 <= d3) { idx[0] = 1; lam[0] = 1; return 1; }
  const double vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) { const double t = d1 / (d1 - d3); idx[0] = 0; idx[1] = 1; lam[0] = 1 - t; lam[1] = t; return 2; }
  const double d5 = -dot(ab, c), d6 = -dot(ac, c);
  if (d6 >= 0 && d5 <= d6) { idx[0] = 2; lam[0] = 1; return 1; }
  const double vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { const double t = d2 / (d2 - d6); idx[0] = 0; idx[1] = 2; lam[0] = 1 - t; lam[1] = t; return 2; }
  const double va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    const double t = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    idx[0] = 1; idx[1] = 2; lam[0] = 1 - t; lam[1] = t; return 2;
  }
  const double den = 1.0 / (va + vb + vc);
  idx[0] = 0; idx[1] = 1; idx[2] = 2;
  lam[0] = va * den; lam[1] = vb * den; lam[2] = vc * den;
  return 3;
}
DRC_HD void closest_triangle(SimplexVert* v, int& n, double* lam) {
  int idx[3];
  n = closest_triangle_w(v[0].w, v[1].w, v[2].w, idx, lam);
  // idx is increasing: compacting in place never reads an overwritten slot
  for (int i = 0; i < n; ++i)
    if (idx[i] != i) v[i] = v[idx[i]];
}
// true when the origin lies inside the tetrahedron.  The faces are examined on the Minkowski points only (three Vec3 each);
// the simplex vertices move once, at the end (this routine used to copy three 9-double vertices per face into local memory).
DRC_HD bool closest_tetra(SimplexVert* v, int& n, double* lam) {
  const int F[4][3] = {{0, 1, 2}, {0, 2, 3}, {0, 3, 1}, {1, 3, 2}};
  const int OPP[4] = {3, 1, 2, 0};
  double best = 1e300;
  int bi[3] = {0, 0, 0};
  double bl[3] = {0, 0, 0};
  int bn = 0;
  bool outside = false;
  for (int f = 0; f < 4; ++f) {
    const Vec3 a = v[F[f][0]].w, b = v[F[f][1]].w, c = v[F[f][2]].w, d = v[OPP[f]].w;
    const Vec3 nrm = cross(b - a, c - a);
    const double so = -dot(a, nrm), sd = dot(d - a, nrm), n2 = dot(nrm, nrm);
    // origin and the opposite vertex on different sides, or a (nearly) flat tetrahedron whose side test is noise
    if (so * sd < 0 || sd * sd <= 1e-20 * n2 * sqrt(n2)) {
      outside = true;
      int ti[3];
      double l[3] = {0, 0, 0};
      const int tn = closest_triangle_w(a, b, c, ti, l);
      Vec3 p = v3(0, 0, 0);
      for (int i = 0; i < tn; ++i) p = p + l[i] * v[F[f][ti[i]]].w;
      const double dd = dot(p, p);
      if (dd < best) {
        best = dd; bn = tn;
        for (int i = 0; i < tn; ++i) { bi[i] = F[f][ti[i]]; bl[i] = l[i]; }
      }
    }
  }
  if (!outside) return true;
  n = bn;
  SimplexVert keep[3];
  for (int i = 0; i < n; ++i) keep[i] = v[bi[i]];
  for (int i = 0; i < n; ++i) { v[i] = keep[i]; lam[i] = bl[i]; }
  return false;
}

struct GjkOut {
  bool intersect;
  double dist;
  Vec3 pa, pb;
  SimplexVert sv[4];
  int n, iters;
};

DRC_HD_NOINLINE void gjk_distance(const Prim& A, const Prim& B, double tol, int max_iter, GjkOut& out) {
  Vec3 d0 = B.c - A.c;
  if (dot(d0, d0) == 0) d0 = v3(1, 0, 0);
  SimplexVert* sv = out.sv;
  double lam[4] = {1, 0, 0, 0};
  int n = 1;
  sv[0].a = support(A, d0); sv[0].b = support(B, -d0); sv[0].w = sv[0].a - sv[0].b;
  Vec3 v = sv[0].w;
  // witness points of the current closest point, carried along with v: they are what a precision-loss restart and the
  // numerical-floor exit need from the PREVIOUS simplex, so that simplex itself is never copied (the per-iteration backup of
  // 4 x 9 doubles was most of the loop's local-memory traffic)
  Vec3 pa = sv[0].a, pb = sv[0].b;
  bool inter = false;
  int it = 0;
  for (; it < max_iter; ++it) {
    const double vv = dot(v, v);
    if (vv <= 1e-30) { inter = true; break; }
    SimplexVert nw;
    nw.a = support(A, -v); nw.b = support(B, v); nw.w = nw.a - nw.b;
    const double gap = vv - dot(v, nw.w);
    if (gap <= tol * sqrt(vv)) break;
    bool dup = false;
    for (int i = 0; i < n; ++i) dup = dup || (norm2(sv[i].w - nw.w) <= 1e-30);
    if (dup) break;
    sv[n++] = nw;
    bool inside = false;
    if (n == 2) closest_segment(sv, n, lam);
    else if (n == 3) closest_triangle(sv, n, lam);
    else inside = closest_tetra(sv, n, lam);
    if (inside) { lam[0] = lam[1] = lam[2] = lam[3] = 0.25; n = 4; inter = true; break; }
    Vec3 nv = v3(0, 0, 0);
    for (int i = 0; i < n; ++i) nv = nv + lam[i] * sv[i].w;
    if (dot(nv, nv) >= vv) {
      // the simplex sub-algorithm lost precision (thin simplex): restart from the segment [current closest point,
      // new vertex] -- a Frank-Wolfe step with exact line search, which improves whenever the gap is positive.
      // The closest point is a valid vertex: a convex combination of support points of A and of B.
      SimplexVert cp;
      cp.w = v; cp.a = pa; cp.b = pb;
      sv[0] = cp; sv[1] = nw; n = 2;
      closest_segment(sv, n, lam);
      nv = v3(0, 0, 0);
      for (int i = 0; i < n; ++i) nv = nv + lam[i] * sv[i].w;
      if (dot(nv, nv) >= vv) {  // numerical floor: the previous closest point stands
        sv[0] = cp; lam[0] = 1; n = 1;
        break;
      }
    }
    v = nv;
    pa = v3(0, 0, 0); pb = v3(0, 0, 0);
    for (int i = 0; i < n; ++i) { pa = pa + lam[i] * sv[i].a; pb = pb + lam[i] * sv[i].b; }
  }
  out.intersect = inter;
  out.n = n;
  out.iters = it + 1;
  if (inter) {  // the overlapping simplex (EPA's seed): witnesses from its weights
    pa = v3(0, 0, 0); pb = v3(0, 0, 0);
    for (int i = 0; i < n; ++i) { pa = pa + lam[i] * sv[i].a; pb = pb + lam[i] * sv[i].b; }
  }
  out.pa = pa; out.pb = pb;
  out.dist = inter ? 0.0 : norm(v);
}

// ---------------------------------------------------------------- EPA (fixed-capacity polytope)
// Expanding polytope algorithm (van den Bergen 2001) for the penetration depth of overlapping shapes.
// Robustness rules (the same in oracle/src/ogeom.h and in the warp-parallel device variant in drc_lib.cu):
//   * the faces removed by a new vertex are found by a FLOOD FILL from the closest face across shared edges, so the
//     removed region is connected and its border (the horizon) is a closed loop even when large coplanar regions of
//     the Minkowski difference (box face x cylinder cap) make the visibility sign of a face noise;
//   * an expansion that would leave a hole (no neighbour across an edge), a degenerate new face or exceed the
//     capacity is not committed: the current closest face is returned (a lower bound of the depth);
//   * new face i (horizon order) reuses the slot of the i-th removed face, the last two are appended.
constexpr int kEpaMaxVert = 104, kEpaMaxFace = 208, kEpaMaxEdge = 96;
constexpr double kEpaVisEps = 1e-12, kEpaMinArea2 = 1e-28;
struct EpaFace {
  short v[3];
  short alive;
  Vec3 n;
  double d;
};
// grow the final GJK simplex to a tetrahedron (T[0..3]); returns the number of vertices reached
template <class Sup>
DRC_HD int epa_seed(Sup sup, const GjkOut& g, SimplexVert* T) {
  int np = g.n;
  for (int i = 0; i < np; ++i) T[i] = g.sv[i];
  const Vec3 axes[6] = {v3(1, 0, 0), v3(-1, 0, 0), v3(0, 1, 0), v3(0, -1, 0), v3(0, 0, 1), v3(0, 0, -1)};
  if (np == 1) {
    for (int k = 0; k < 6; ++k) {
      const SimplexVert s = sup(axes[k]);
      if (norm2(T[0].w - s.w) >= 1e-20) { T[np++] = s; break; }
    }
  }
  if (np == 2) {
    const Vec3 e = T[1].w - T[0].w;
    Vec3 best = v3(0, 0, 0);
    double bl = -1;
    for (int k = 0; k < 6; ++k) {
      const Vec3 c = cross(e, axes[k]);
      if (dot(c, c) <= 1e-20) continue;
      for (int sgn = -1; sgn <= 1; sgn += 2) {
        const SimplexVert s = sup((double)sgn * c);
        const double area = norm(cross(e, s.w - T[0].w));
        if (area > bl) { bl = area; best = (double)sgn * c; }
      }
    }
    T[np++] = sup(best);
  }
  if (np == 3) {
    const Vec3 nrm = cross(T[1].w - T[0].w, T[2].w - T[0].w);
    const SimplexVert s1 = sup(nrm), s2 = sup(-nrm);
    const double h1 = fabs(dot(s1.w - T[0].w, nrm)), h2 = fabs(dot(s2.w - T[0].w, nrm));
    T[np++] = h1 >= h2 ? s1 : s2;
  }
  if (np == 4 && dot(cross(T[1].w - T[0].w, T[2].w - T[0].w), T[3].w - T[0].w) > 0) { const SimplexVert t = T[1]; T[1] = T[2]; T[2] = t; }
  return np;
}
// witness points: barycentric coordinates of the origin's projection on the closest face
DRC_HD void epa_witness(const SimplexVert& t0, const SimplexVert& t1, const SimplexVert& t2, Vec3 fn, double fd, PairResult& out) {
  const Vec3 pr = fd * fn;
  const Vec3 v0 = t1.w - t0.w, v1 = t2.w - t0.w, v2 = pr - t0.w;
  const double d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
  const double den = d00 * d11 - d01 * d01;
  const double l1 = den != 0 ? (d11 * d20 - d01 * d21) / den : 0.0, l2 = den != 0 ? (d00 * d21 - d01 * d20) / den : 0.0;
  const double l0 = 1 - l1 - l2;
  out.pa = l0 * t0.a + l1 * t1.a + l2 * t2.a;
  out.pb = l0 * t0.b + l1 * t1.b + l2 * t2.b;
  out.d = -fd;
}

DRC_HD_NOINLINE void epa_penetration(const Prim& A, const Prim& B, const GjkOut& g, double tol, int max_iter, PairResult& out) {
  SimplexVert P[kEpaMaxVert];
  EpaFace F[kEpaMaxFace];
  short E[kEpaMaxEdge][2];
  short killed[kEpaMaxEdge], stack[kEpaMaxEdge];
  unsigned char mark[kEpaMaxFace];  // 0 untested, 1 visible, 2 tested: not visible
  auto sup = [&](Vec3 d) { SimplexVert s; s.a = support(A, d); s.b = support(B, -d); s.w = s.a - s.b; return s; };
  int np = epa_seed(sup, g, P), nf = 0;
  out.d = 0; out.pa = g.pa; out.pb = g.pb;
  if (np < 4) return;
  auto make_face = [&](int slot, int a, int b, int c) {
    EpaFace& f = F[slot];
    f.v[0] = (short)a; f.v[1] = (short)b; f.v[2] = (short)c;
    const Vec3 nrm = cross(P[b].w - P[a].w, P[c].w - P[a].w);
    const double l = norm(nrm);
    f.n = l > 0 ? (1.0 / l) * nrm : v3(0, 0, 1);
    f.d = dot(f.n, P[a].w);
    f.alive = 1;
  };
  make_face(0, 0, 1, 2); make_face(1, 0, 3, 1); make_face(2, 0, 2, 3); make_face(3, 1, 3, 2);
  nf = 4;
  int bestf = -1;
  for (int it = 0; it < max_iter; ++it) {
    bestf = -1;
    double bd = 1e300;
    for (int i = 0; i < nf; ++i) if (F[i].alive && F[i].d < bd) { bd = F[i].d; bestf = i; }
    if (bestf < 0) break;
    const SimplexVert s = sup(F[bestf].n);
    if (dot(F[bestf].n, s.w) - F[bestf].d <= tol) break;
    if (np >= kEpaMaxVert) break;
    // flood fill of the faces visible from s
    for (int i = 0; i < nf; ++i) mark[i] = 0;
    int nk = 0, ne = 0, sp = 0;
    bool bad = false;
    mark[bestf] = 1; killed[nk++] = (short)bestf; stack[sp++] = (short)bestf;
    while (sp > 0 && !bad) {
      const int f = stack[--sp];
      for (int e = 0; e < 3 && !bad; ++e) {
        const short a = F[f].v[e], b = F[f].v[(e + 1) % 3];
        int gn = -1;
        for (int i = 0; i < nf; ++i) {
          if (!F[i].alive) continue;
          const short* v = F[i].v;
          if ((v[0] == b && v[1] == a) || (v[1] == b && v[2] == a) || (v[2] == b && v[0] == a)) { gn = i; break; }
        }
        if (gn < 0) { bad = true; break; }
        if (mark[gn] == 1) continue;
        if (mark[gn] == 0) {
          const bool vis = dot(F[gn].n, s.w - P[F[gn].v[0]].w) > kEpaVisEps;
          mark[gn] = vis ? 1 : 2;
          if (vis) {
            if (nk >= kEpaMaxEdge - 2) { bad = true; break; }
            killed[nk++] = (short)gn; stack[sp++] = (short)gn;
            continue;
          }
        }
        if (ne >= kEpaMaxEdge) { bad = true; break; }
        E[ne][0] = a; E[ne][1] = b; ++ne;
      }
    }
    if (bad || ne != nk + 2 || nf + 2 > kEpaMaxFace) break;
    for (int k = 0; k < ne && !bad; ++k)
      bad = norm2(cross(P[E[k][1]].w - P[E[k][0]].w, s.w - P[E[k][0]].w)) <= kEpaMinArea2;
    if (bad) break;
    // commit
    const int idx = np;
    P[np++] = s;
    for (int k = 0; k < ne; ++k) make_face(k < nk ? (int)killed[k] : nf + (k - nk), E[k][0], E[k][1], idx);
    nf += 2;
  }
  if (bestf < 0) return;
  const EpaFace& f = F[bestf];
  epa_witness(P[f.v[0]], P[f.v[1]], P[f.v[2]], f.n, f.d, out);
}

// ---------------------------------------------------------------- exact closed-form dispatcher
DRC_HD bool has_closed_form(int ta, int tb) {
  if (ta == kConvex || tb == kConvex) return false;  // mesh hulls always go through GJK / EPA
  return ta == kSphere || tb == kSphere || (ta == kCapsule && tb == kCapsule);
}
DRC_HD PairResult closed_form_distance(const Prim& A, const Prim& B) {
  PairResult r;
  if (A.type == kSphere) {
    if (B.type == kSphere) r = sphere_sphere(A.c, A.r, B.c, B.r);
    else if (B.type == kCylinder) r = sphere_cylinder(A.c, A.r, B);
    else if (B.type == kBox) r = sphere_box(A.c, A.r, B);
    else r = sphere_capsule(A.c, A.r, B);
  } else if (B.type == kSphere) {
    if (A.type == kCylinder) r = sphere_cylinder(B.c, B.r, A);
    else if (A.type == kBox) r = sphere_box(B.c, B.r, A);
    else r = sphere_capsule(B.c, B.r, A);
    const Vec3 t = r.pa; r.pa = r.pb; r.pb = t;
  } else {
    r = capsule_capsule(A, B);
  }
  return r;
}

}  // namespace drc
