// ORACLE -- test infrastructure only. PARITY UNPINNED (no reference goldens exist).
// Shape-shape signed distance with witness points, restating what the reference obtains from
// pinocchio::computeDistances -> hpp-fcl/coal `distance()` (robot_data.cpp:429; un-vendored,
// un-pinned).  Published algorithms restated:
//   * sphere-sphere, sphere-cylinder, sphere-box, sphere-capsule, capsule-capsule: closed form
//     (hpp-fcl ships specialisations for these pairs);
//   * every other pair of convex primitives: GJK (Gilbert-Johnson-Keerthi 1988) on the Minkowski
//     difference for the separated case, EPA (van den Bergen 2001) for the penetration depth.
// Results are in the WORLD frame: distance (negative = penetration depth), pa on shape A, pb on shape B.
#pragma once
#include "orbd.h"

namespace orc {

struct Shape {
  int type;
  V3 prm;  // sphere: r | cylinder/capsule: r, half length (axis = local z) | box: half extents
  SE3 T;   // world placement
  const double* verts = nullptr;  // GEOM_CONVEX: hull vertices (geometry frame), nvert of them
  int nvert = 0;
};
inline Shape make_shape(const Model& m, int g, const SE3& T) {
  Shape s;
  s.type = m.geom_type[g]; s.prm = m.geom_param[g]; s.T = T;
  if (s.type == GEOM_CONVEX) { s.verts = m.hull.data() + 3 * m.hull_off[g]; s.nvert = m.hull_n[g]; }
  return s;
}
struct DistResult {
  double d;
  V3 pa, pb;
  int gjk_iters = 0, epa_iters = 0;
};

struct GeomParams {
  double gjk_tol = 1e-10;   // absolute tolerance on the distance duality gap
  int gjk_max_iter = 128;
  double epa_tol = 1e-6;   // hpp-fcl's default EPA tolerance; curved pairs converge like 1/k^2, 1e-10 is never reached
  int epa_max_iter = 96;
};

// ---------------------------------------------------------------- closed forms
inline DistResult sphere_sphere(const V3& c1, double r1, const V3& c2, double r2) {
  V3 diff = c2 - c1;
  double len = norm(diff);
  V3 n = len > 0 ? (1.0 / len) * diff : V3(1, 0, 0);
  return {len - r1 - r2, c1 + r1 * n, c2 - r2 * n};
}

// Closest point on the surface of a solid (given in its local frame by `closest_local`) to a point,
// then dress it as sphere(A)-vs-solid(B).  sd = signed distance of the centre to the solid.
inline DistResult sphere_vs_solid(const V3& cs, double rs, const SE3& T, const V3& q_local, double sd) {
  V3 q = T.R * q_local + T.p;
  V3 dir = q - cs;  // towards the closest surface point
  double len = norm(dir);
  V3 n = len > 0 ? (1.0 / len) * dir : V3(1, 0, 0);
  DistResult r;
  r.d = sd - rs;
  r.pb = q;
  // centre outside: the sphere's witness faces the solid; centre inside: it is the deepest point.
  r.pa = sd >= 0 ? cs + rs * n : cs - rs * n;
  return r;
}

inline DistResult sphere_cylinder(const V3& cs, double rs, const SE3& T, double r, double h) {
  V3 x = tmul(T.R, cs - T.p);
  double rho = std::sqrt(x.x * x.x + x.y * x.y);
  V3 q;
  double sd;
  if (std::fabs(x.z) <= h && rho <= r) {  // centre inside the solid cylinder
    double dr = r - rho, dz = h - std::fabs(x.z);
    if (dr < dz) {
      q = rho > 0 ? V3(x.x / rho * r, x.y / rho * r, x.z) : V3(r, 0, x.z);
      sd = -dr;
    } else {
      q = V3(x.x, x.y, x.z >= 0 ? h : -h);
      sd = -dz;
    }
  } else {
    double s = rho > r ? r / rho : 1.0;
    q = V3(x.x * s, x.y * s, std::min(std::max(x.z, -h), h));
    sd = norm(x - q);
  }
  return sphere_vs_solid(cs, rs, T, q, sd);
}

inline DistResult sphere_box(const V3& cs, double rs, const SE3& T, const V3& hb) {
  V3 x = tmul(T.R, cs - T.p);
  V3 q;
  double sd;
  if (std::fabs(x.x) <= hb.x && std::fabs(x.y) <= hb.y && std::fabs(x.z) <= hb.z) {
    double best = std::numeric_limits<double>::max();
    int ax = 0;
    for (int i = 0; i < 3; ++i) {
      double di = hb[i] - std::fabs(x[i]);
      if (di < best) { best = di; ax = i; }
    }
    q = x;
    q[ax] = x[ax] >= 0 ? hb[ax] : -hb[ax];
    sd = -best;
  } else {
    q = V3(std::min(std::max(x.x, -hb.x), hb.x), std::min(std::max(x.y, -hb.y), hb.y),
           std::min(std::max(x.z, -hb.z), hb.z));
    sd = norm(x - q);
  }
  return sphere_vs_solid(cs, rs, T, q, sd);
}

// closest points between segments p1+s*d1 (s in [0,1]) and p2+t*d2 (Ericson, RTCD 5.1.9)
inline void segment_segment(const V3& p1, const V3& d1, const V3& p2, const V3& d2, double& s, double& t) {
  V3 r = p1 - p2;
  double a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r);
  const double eps = 1e-300;
  if (a <= eps && e <= eps) { s = t = 0; return; }
  if (a <= eps) { s = 0; t = std::min(std::max(f / e, 0.0), 1.0); return; }
  double c = dot(d1, r);
  if (e <= eps) { t = 0; s = std::min(std::max(-c / a, 0.0), 1.0); return; }
  double b = dot(d1, d2), den = a * e - b * b;
  s = den > 1e-14 * a * e ? std::min(std::max((b * f - c * e) / den, 0.0), 1.0) : 0.0;
  t = (b * s + f) / e;
  if (t < 0) { t = 0; s = std::min(std::max(-c / a, 0.0), 1.0); }
  else if (t > 1) { t = 1; s = std::min(std::max((b - c) / a, 0.0), 1.0); }
}
inline DistResult capsule_capsule(const SE3& T1, double r1, double h1, const SE3& T2, double r2, double h2) {
  V3 a1 = T1.R.col(2), a2 = T2.R.col(2);
  V3 p1 = T1.p - h1 * a1, p2 = T2.p - h2 * a2;
  double s, t;
  segment_segment(p1, (2 * h1) * a1, p2, (2 * h2) * a2, s, t);
  return sphere_sphere(p1 + (2 * h1 * s) * a1, r1, p2 + (2 * h2 * t) * a2, r2);
}
inline DistResult sphere_capsule(const V3& cs, double rs, const SE3& T, double r, double h) {
  V3 a = T.R.col(2);
  double t = std::min(std::max(dot(cs - T.p, a), -h), h);
  return sphere_sphere(cs, rs, T.p + t * a, r);
}

// ---------------------------------------------------------------- support mappings (world frame)
inline V3 support(const Shape& S, const V3& d) {
  V3 dl = tmul(S.T.R, d), s;
  switch (S.type) {
    case GEOM_SPHERE: {
      double n = norm(dl);
      s = n > 0 ? (S.prm.x / n) * dl : V3();
      break;
    }
    case GEOM_BOX:
      s = V3(dl.x >= 0 ? S.prm.x : -S.prm.x, dl.y >= 0 ? S.prm.y : -S.prm.y, dl.z >= 0 ? S.prm.z : -S.prm.z);
      break;
    case GEOM_CONVEX: {  // mesh hull: the vertex furthest along the direction
      double bv = -1e300;
      for (int i = 0; i < S.nvert; ++i) {
        const V3 v(S.verts[3 * i], S.verts[3 * i + 1], S.verts[3 * i + 2]);
        const double t = dot(dl, v);
        if (t > bv) { bv = t; s = v; }
      }
      break;
    }
    case GEOM_CYLINDER: {
      double sg = std::sqrt(dl.x * dl.x + dl.y * dl.y);
      double k = sg > 0 ? S.prm.x / sg : 0.0;
      s = V3(k * dl.x, k * dl.y, dl.z >= 0 ? S.prm.y : -S.prm.y);
      break;
    }
    default: {  // capsule
      double n = norm(dl);
      s = n > 0 ? (S.prm.x / n) * dl : V3();
      s.z += dl.z >= 0 ? S.prm.y : -S.prm.y;
    }
  }
  return S.T.R * s + S.T.p;
}

struct SVert { V3 w, a, b; };

// ---- closest point to the origin on a simplex; returns barycentric weights and keeps only the
//      vertices that support the closest point (Ericson RTCD 5.1.2/5.1.5/5.1.6 Voronoi-region tests).
inline void closest_on_segment(SVert* v, int& n, double* lam) {
  V3 a = v[0].w, b = v[1].w, ab = b - a;
  double t = -dot(a, ab), den = dot(ab, ab);
  if (t <= 0 || den <= 0) { n = 1; lam[0] = 1; return; }
  if (t >= den) { v[0] = v[1]; n = 1; lam[0] = 1; return; }
  t /= den;
  lam[0] = 1 - t; lam[1] = t;
}
inline void closest_on_triangle(SVert* v, int& n, double* lam) {
  V3 a = v[0].w, b = v[1].w, c = v[2].w;
  V3 ab = b - a, ac = c - a, ap = -a;
  double d1 = dot(ab, ap), d2 = dot(ac, ap);
  if (d1 <= 0 && d2 <= 0) { n = 1; lam[0] = 1; return; }
  V3 bp = -b;
  double d3 = dot(ab, bp), d4 = dot(ac, bp);
  if (d3 >= 0 && d4 <= d3) { v[0] = v[1]; n = 1; lam[0] = 1; return; }
  double vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) { double t = d1 / (d1 - d3); n = 2; lam[0] = 1 - t; lam[1] = t; return; }
  V3 cp = -c;
  double d5 = dot(ab, cp), d6 = dot(ac, cp);
  if (d6 >= 0 && d5 <= d6) { v[0] = v[2]; n = 1; lam[0] = 1; return; }
  double vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { double t = d2 / (d2 - d6); v[1] = v[2]; n = 2; lam[0] = 1 - t; lam[1] = t; return; }
  double va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    double t = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    v[0] = v[1]; v[1] = v[2]; n = 2; lam[0] = 1 - t; lam[1] = t; return;
  }
  double den = 1.0 / (va + vb + vc);
  lam[0] = va * den; lam[1] = vb * den; lam[2] = vc * den;
}
// returns true when the origin is inside the tetrahedron
inline bool closest_on_tetra(SVert* v, int& n, double* lam) {
  // test each face whose outside half-space contains the origin; keep the best
  static const int F[4][3] = {{0, 1, 2}, {0, 2, 3}, {0, 3, 1}, {1, 3, 2}};
  static const int OPP[4] = {3, 1, 2, 0};
  double best = std::numeric_limits<double>::max();
  SVert bestv[3];
  double bestl[3];
  int bestn = 0;
  bool outside_any = false;
  for (int f = 0; f < 4; ++f) {
    const V3 &a = v[F[f][0]].w, &b = v[F[f][1]].w, &c = v[F[f][2]].w, &d = v[OPP[f]].w;
    V3 nrm = cross(b - a, c - a);
    double so = dot(-a, nrm), sd = dot(d - a, nrm), n2 = dot(nrm, nrm);
    // origin and the opposite vertex on different sides, or a (nearly) flat tetrahedron whose side test is noise
    if (so * sd < 0 || sd * sd <= 1e-20 * n2 * std::sqrt(n2)) {
      outside_any = true;
      SVert t[3] = {v[F[f][0]], v[F[f][1]], v[F[f][2]]};
      double l[3];
      int tn = 3;
      closest_on_triangle(t, tn, l);
      V3 p;
      for (int i = 0; i < tn; ++i) p += l[i] * t[i].w;
      double dd = dot(p, p);
      if (dd < best) {
        best = dd; bestn = tn;
        for (int i = 0; i < tn; ++i) { bestv[i] = t[i]; bestl[i] = l[i]; }
      }
    }
  }
  if (!outside_any) return true;
  n = bestn;
  for (int i = 0; i < n; ++i) { v[i] = bestv[i]; lam[i] = bestl[i]; }
  return false;
}

struct GjkResult {
  bool intersect = false;
  double dist = 0;
  V3 pa, pb;
  SVert simplex[4];
  int nsimplex = 0;
  int iters = 0;
};

inline GjkResult gjk(const Shape& A, const Shape& B, const GeomParams& gp) {
  GjkResult r;
  V3 d0 = B.T.p - A.T.p;
  if (dot(d0, d0) == 0) d0 = V3(1, 0, 0);
  SVert sv[4];
  double lam[4] = {1, 0, 0, 0};
  int n = 1;
  sv[0].a = support(A, d0); sv[0].b = support(B, -d0); sv[0].w = sv[0].a - sv[0].b;
  V3 v = sv[0].w;
  auto finish = [&](bool inter) {
    r.intersect = inter;
    r.nsimplex = n;
    V3 pa, pb;
    for (int i = 0; i < n; ++i) { r.simplex[i] = sv[i]; pa += lam[i] * sv[i].a; pb += lam[i] * sv[i].b; }
    r.pa = pa; r.pb = pb; r.dist = inter ? 0.0 : norm(v);
    return r;
  };
  for (int it = 0; it < gp.gjk_max_iter; ++it) {
    r.iters = it + 1;
    double vv = dot(v, v);
    if (vv <= 1e-30) return finish(true);
    SVert nw;
    nw.a = support(A, -v); nw.b = support(B, v); nw.w = nw.a - nw.b;
    double gap = vv - dot(v, nw.w);  // >= 0 up to rounding;  gap/|v| bounds the distance error
    if (gap <= gp.gjk_tol * std::sqrt(vv)) return finish(false);
    bool dup = false;
    for (int i = 0; i < n; ++i) {
      V3 e = sv[i].w - nw.w;
      if (dot(e, e) <= 1e-30) dup = true;
    }
    if (dup) return finish(false);
    SVert prev_sv[4];
    double prev_lam[4];
    const int prev_n = n;
    for (int i = 0; i < n; ++i) { prev_sv[i] = sv[i]; prev_lam[i] = lam[i]; }
    sv[n++] = nw;
    bool inside = false;
    if (n == 2) closest_on_segment(sv, n, lam);
    else if (n == 3) closest_on_triangle(sv, n, lam);
    else inside = closest_on_tetra(sv, n, lam);
    if (inside) { for (int i = 0; i < 4; ++i) lam[i] = 0.25; n = 4; return finish(true); }
    V3 nv;
    for (int i = 0; i < n; ++i) nv += lam[i] * sv[i].w;
    if (dot(nv, nv) >= vv) {
      // the simplex sub-algorithm lost precision (thin simplex): restart from the segment [current closest point,
      // new vertex] -- a Frank-Wolfe step with exact line search (same rule as the product's GJK)
      SVert cp;
      cp.w = v;
      for (int i = 0; i < prev_n; ++i) { cp.a += prev_lam[i] * prev_sv[i].a; cp.b += prev_lam[i] * prev_sv[i].b; }
      sv[0] = cp; sv[1] = nw; n = 2;
      closest_on_segment(sv, n, lam);
      nv = V3();
      for (int i = 0; i < n; ++i) nv += lam[i] * sv[i].w;
      if (dot(nv, nv) >= vv) {  // no progress: numerical floor reached, keep the previous simplex
        n = prev_n;
        for (int i = 0; i < n; ++i) { sv[i] = prev_sv[i]; lam[i] = prev_lam[i]; }
        return finish(false);
      }
    }
    v = nv;
  }
  return finish(false);
}

// ---------------------------------------------------------------- EPA
struct EpaFace { int v[3]; V3 n; double d; bool alive; };

inline DistResult epa(const Shape& A, const Shape& B, const GjkResult& g, const GeomParams& gp) {
  std::vector<SVert> P(g.simplex, g.simplex + g.nsimplex);
  auto sup = [&](const V3& d) { SVert s; s.a = support(A, d); s.b = support(B, -d); s.w = s.a - s.b; return s; };
  // grow the GJK simplex to a non-degenerate tetrahedron that contains the origin
  auto far_enough = [&](const SVert& s) {
    for (auto& p : P) { V3 e = p.w - s.w; if (dot(e, e) < 1e-20) return false; }
    return true;
  };
  static const V3 axes[6] = {V3(1, 0, 0), V3(-1, 0, 0), V3(0, 1, 0), V3(0, -1, 0), V3(0, 0, 1), V3(0, 0, -1)};
  if (P.size() == 1) {
    for (auto& ax : axes) { SVert s = sup(ax); if (far_enough(s)) { P.push_back(s); break; } }
  }
  if (P.size() == 2) {
    V3 e = P[1].w - P[0].w;
    V3 best; double bl = -1;
    for (auto& ax : axes) {
      V3 c = cross(e, ax);
      if (dot(c, c) <= 1e-20) continue;
      for (int sgn = -1; sgn <= 1; sgn += 2) {
        SVert s = sup(double(sgn) * c);
        double area = norm(cross(e, s.w - P[0].w));
        if (area > bl) { bl = area; best = double(sgn) * c; }
      }
    }
    P.push_back(sup(best));
  }
  if (P.size() == 3) {
    V3 nrm = cross(P[1].w - P[0].w, P[2].w - P[0].w);
    SVert s1 = sup(nrm), s2 = sup(-nrm);
    double h1 = std::fabs(dot(s1.w - P[0].w, nrm)), h2 = std::fabs(dot(s2.w - P[0].w, nrm));
    P.push_back(h1 >= h2 ? s1 : s2);
  }
  DistResult out;
  out.d = 0; out.pa = g.pa; out.pb = g.pb;
  if (P.size() < 4) return out;
  // Robustness rules (shared with the product's EPA, csrc/drc_geom.h): removed faces found by flood fill from the
  // closest face across shared edges (connected region, closed horizon even with coplanar faces); an expansion
  // that would leave a hole / degenerate face / exceed the capacity is not committed; new face i reuses the slot
  // of the i-th removed face, the last two are appended.
  const int kMaxVert = 104, kMaxFace = 208, kMaxEdge = 96;
  const double kVisEps = 1e-12, kMinArea2 = 1e-28;
  std::vector<EpaFace> F;
  auto make_face = [&](int slot, int a, int b, int c) {
    EpaFace f;
    f.v[0] = a; f.v[1] = b; f.v[2] = c;
    V3 nrm = cross(P[b].w - P[a].w, P[c].w - P[a].w);
    double l = norm(nrm);
    f.n = l > 0 ? (1.0 / l) * nrm : V3(0, 0, 1);
    f.d = dot(f.n, P[a].w);
    f.alive = true;
    if (slot >= int(F.size())) F.resize(slot + 1);
    F[slot] = f;
  };
  // orient the tetrahedron so that every face normal points away from the 4th vertex
  if (dot(cross(P[1].w - P[0].w, P[2].w - P[0].w), P[3].w - P[0].w) > 0) std::swap(P[1], P[2]);
  make_face(0, 0, 1, 2); make_face(1, 0, 3, 1); make_face(2, 0, 2, 3); make_face(3, 1, 3, 2);
  int bestf = -1;
  for (int it = 0; it < gp.epa_max_iter; ++it) {
    out.epa_iters = it + 1;
    bestf = -1;
    double bd = std::numeric_limits<double>::max();
    for (size_t i = 0; i < F.size(); ++i)
      if (F[i].alive && F[i].d < bd) { bd = F[i].d; bestf = int(i); }
    if (bestf < 0) break;
    SVert s = sup(F[bestf].n);
    double ext = dot(F[bestf].n, s.w) - F[bestf].d;
    if (ext <= gp.epa_tol) break;
    if (int(P.size()) >= kMaxVert) break;
    const int nf = int(F.size());
    std::vector<int> mark(nf, 0), killed, stack;   // mark: 0 untested, 1 visible, 2 tested: not visible
    std::vector<std::pair<int, int>> edges;
    bool bad = false;
    mark[bestf] = 1; killed.push_back(bestf); stack.push_back(bestf);
    while (!stack.empty() && !bad) {
      const int f = stack.back();
      stack.pop_back();
      for (int e = 0; e < 3 && !bad; ++e) {
        const int a = F[f].v[e], b = F[f].v[(e + 1) % 3];
        int gn = -1;
        for (int i = 0; i < nf; ++i) {
          if (!F[i].alive) continue;
          const int* v = F[i].v;
          if ((v[0] == b && v[1] == a) || (v[1] == b && v[2] == a) || (v[2] == b && v[0] == a)) { gn = i; break; }
        }
        if (gn < 0) { bad = true; break; }
        if (mark[gn] == 1) continue;
        if (mark[gn] == 0) {
          const bool vis = dot(F[gn].n, s.w - P[F[gn].v[0]].w) > kVisEps;
          mark[gn] = vis ? 1 : 2;
          if (vis) {
            if (int(killed.size()) >= kMaxEdge - 2) { bad = true; break; }
            killed.push_back(gn); stack.push_back(gn);
            continue;
          }
        }
        if (int(edges.size()) >= kMaxEdge) { bad = true; break; }
        edges.emplace_back(a, b);
      }
    }
    const int nk = int(killed.size()), ne = int(edges.size());
    if (bad || ne != nk + 2 || nf + 2 > kMaxFace) break;
    for (auto& e : edges) {
      V3 c = cross(P[e.second].w - P[e.first].w, s.w - P[e.first].w);
      if (dot(c, c) <= kMinArea2) bad = true;
    }
    if (bad) break;
    const int idx = int(P.size());
    P.push_back(s);
    for (int k = 0; k < ne; ++k) make_face(k < nk ? killed[k] : nf + (k - nk), edges[k].first, edges[k].second, idx);
  }
  if (bestf < 0) return out;
  // witness points: barycentric coordinates of the origin's projection on the closest face
  const EpaFace& f = F[bestf];
  SVert t[3] = {P[f.v[0]], P[f.v[1]], P[f.v[2]]};
  V3 proj = f.d * f.n;
  V3 a = t[0].w, b = t[1].w, c = t[2].w;
  V3 v0 = b - a, v1 = c - a, v2 = proj - a;
  double d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
  double den = d00 * d11 - d01 * d01;
  double l1 = den != 0 ? (d11 * d20 - d01 * d21) / den : 0, l2 = den != 0 ? (d00 * d21 - d01 * d20) / den : 0;
  double l0 = 1 - l1 - l2;
  out.pa = l0 * t[0].a + l1 * t[1].a + l2 * t[2].a;
  out.pb = l0 * t[0].b + l1 * t[1].b + l2 * t[2].b;
  out.d = -f.d;
  return out;
}

// ---------------------------------------------------------------- dispatcher
inline DistResult swap_ab(DistResult r) { std::swap(r.pa, r.pb); return r; }

inline DistResult shape_distance(const Shape& A, const Shape& B, const GeomParams& gp) {
  const bool hull = A.type == GEOM_CONVEX || B.type == GEOM_CONVEX;  // mesh hulls: always GJK / EPA
  if (!hull && A.type == GEOM_SPHERE) {
    switch (B.type) {
      case GEOM_SPHERE: return sphere_sphere(A.T.p, A.prm.x, B.T.p, B.prm.x);
      case GEOM_CYLINDER: return sphere_cylinder(A.T.p, A.prm.x, B.T, B.prm.x, B.prm.y);
      case GEOM_BOX: return sphere_box(A.T.p, A.prm.x, B.T, B.prm);
      default: return sphere_capsule(A.T.p, A.prm.x, B.T, B.prm.x, B.prm.y);
    }
  }
  if (!hull && B.type == GEOM_SPHERE) return swap_ab(shape_distance(B, A, gp));
  if (A.type == GEOM_CAPSULE && B.type == GEOM_CAPSULE)
    return capsule_capsule(A.T, A.prm.x, A.prm.y, B.T, B.prm.x, B.prm.y);
  GjkResult g = gjk(A, B, gp);
  if (!g.intersect) {
    DistResult r;
    r.d = g.dist; r.pa = g.pa; r.pb = g.pb; r.gjk_iters = g.iters;
    return r;
  }
  DistResult r = epa(A, B, g, gp);
  r.gjk_iters = g.iters;
  return r;
}

// reference Manipulator::RobotData::getMinDistance (robot_data.cpp:424-517)
struct MinDistResult {
  double d;
  double grad[MAXV], grad_dot[MAXV];
  int pair;
  V3 pa, pb;
};

inline void min_distance(const Model& m, const State& s, bool with_grad, bool with_graddot, const GeomParams& gp,
                         MinDistResult& out) {
  const int n = m.nv;
  out.d = std::numeric_limits<double>::max();
  out.pair = -1;
  for (int i = 0; i < MAXV; ++i) out.grad[i] = out.grad_dot[i] = 0;
  // updateGeometryPlacements
  std::vector<Shape> sh(m.ng);
  for (int g = 0; g < m.ng; ++g) sh[g] = make_shape(m, g, m.geom_parent[g] < 0 ? m.geom_place[g] : s.oMi[m.geom_parent[g]] * m.geom_place[g]);
  DistResult best;
  for (size_t k = 0; k < m.pair_a.size(); ++k) {
    DistResult r = shape_distance(sh[m.pair_a[k]], sh[m.pair_b[k]], gp);
    if (r.d < out.d) { out.d = r.d; out.pair = int(k); best = r; }
  }
  if (out.pair < 0) { out.d = 0; return; }
  out.pa = best.pa; out.pb = best.pb;
  if (!with_grad && !with_graddot) return;
  const int jA = m.geom_parent[m.pair_a[out.pair]], jB = m.geom_parent[m.pair_b[out.pair]];
  const V3 pA = best.pa, pB = best.pb;
  V3 dn = pB - pA;
  double ln = norm(dn);
  V3 nrm = ln > 1e-12 ? (1.0 / ln) * dn : V3();  // Q5 guard (reference: NaN when pA==pB)
  // getJointJacobian(LOCAL_WORLD_ALIGNED) of the two parent joints (zero for the universe)
  std::vector<double> JjA(6 * n), JjB(6 * n);
  V3 oA = jA < 0 ? V3() : s.oMi[jA].p, oB = jB < 0 ? V3() : s.oMi[jB].p;
  translate_lwa(m, jA, oA, s.J, JjA.data());
  translate_lwa(m, jB, oB, s.J, JjB.data());
  const V3 rA = pA - oA, rB = pB - oB;
  // point Jacobians JP = Jv - skew(r) Jw  (robot_data.cpp:489-490)
  auto point_jac = [&](const std::vector<double>& Jj, const V3& r, int col) {
    V3 lin(Jj[0 * n + col], Jj[1 * n + col], Jj[2 * n + col]), ang(Jj[3 * n + col], Jj[4 * n + col], Jj[5 * n + col]);
    return lin - cross(r, ang);
  };
  for (int j = 0; j < n; ++j) {
    V3 dj = point_jac(JjB, rB, j) - point_jac(JjA, rA, j);
    out.grad[j] = dot(nrm, dj);
    if (out.d < 0) out.grad[j] = -out.grad[j];
  }
  if (!with_graddot) return;
  std::vector<double> dJjA(6 * n), dJjB(6 * n);
  lwa_time_variation(m, jA, oA, s.ov, s.J, s.dJ, dJjA.data());
  lwa_time_variation(m, jB, oB, s.ov, s.J, s.dJ, dJjB.data());
  V3 pA_dot, pB_dot, vA, vB;
  for (int j = 0; j < n; ++j) {
    pA_dot += s.qd[j] * point_jac(JjA, rA, j);
    pB_dot += s.qd[j] * point_jac(JjB, rB, j);
    vA += s.qd[j] * V3(JjA[0 * n + j], JjA[1 * n + j], JjA[2 * n + j]);
    vB += s.qd[j] * V3(JjB[0 * n + j], JjB[1 * n + j], JjB[2 * n + j]);
  }
  const V3 rA_dot = pA_dot - vA, rB_dot = pB_dot - vB;
  auto point_jac_dot = [&](const std::vector<double>& Jj, const std::vector<double>& dJj, const V3& r, const V3& rd,
                           int col) {
    V3 ang(Jj[3 * n + col], Jj[4 * n + col], Jj[5 * n + col]);
    V3 dlin(dJj[0 * n + col], dJj[1 * n + col], dJj[2 * n + col]), dang(dJj[3 * n + col], dJj[4 * n + col], dJj[5 * n + col]);
    return dlin - (cross(rd, ang) + cross(r, dang));
  };
  for (int j = 0; j < n; ++j) {
    V3 dj = point_jac_dot(JjB, dJjB, rB, rB_dot, j) - point_jac_dot(JjA, dJjA, rA, rA_dot, j);
    out.grad_dot[j] = dot(nrm, dj);  // n_dot neglected; NOT sign-flipped (robot_data.cpp:513)
  }
}

}  // namespace orc
