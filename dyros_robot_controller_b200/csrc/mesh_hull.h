// drc_b200 -- mesh collision geometry for the model compiler (host only).
//
// The reference hands URDF <mesh> collision elements to pinocchio::urdf::buildGeom, which resolves `package://` through
// the packages_path argument (src/manipulator/robot_data.cpp:24-34) and loads the file with assimp.  Here a mesh becomes
// the CONVEX HULL of its vertices (GeomType kConvex): the narrow phase needs a support function only, i.e. the set of hull
// vertices.  For a convex mesh that is the mesh; for a non-convex one the hull distance is a lower bound of the mesh
// distance (conservative for a collision-avoidance constraint) -- the one documented difference to hpp-fcl's BVH models.
//
// Readers: STL (binary and ASCII), Wavefront OBJ (`v` records), COLLADA (.dae: every <float_array> that feeds a POSITION
// input, with the asset's unit scale).  Hull: extreme vertices along a direction fan, then an exact repair loop -- every
// vertex that lies outside the hull of the selection by more than 1e-9 of the mesh size (GJK point-to-hull distance, the
// narrow phase's own routine) is added until none is left; the selection then spans the hull to that tolerance.
#pragma once
#include <algorithm>
#include <array>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "drc_geom.h"
#include "xml_mini.h"

namespace drc {
namespace mesh {

inline std::string lower_ext(const std::string& path) {
  const size_t dot = path.find_last_of('.');
  std::string e = dot == std::string::npos ? "" : path.substr(dot + 1);
  for (char& c : e) c = (char)std::tolower((unsigned char)c);
  return e;
}

// URDF mesh filename -> file system path.  package://<pkg>/<rest> -> <packages_path>/<pkg>/<rest> (pinocchio's package
// directories are the PARENT directories of the packages); file://<abs>; anything else relative to the URDF's directory.
inline std::string resolve(const std::string& filename, const std::string& urdf_dir, const std::string& packages_path) {
  const std::string pk = "package://", fl = "file://";
  if (filename.compare(0, pk.size(), pk) == 0) {
    if (packages_path.empty()) throw std::runtime_error("urdf: mesh '" + filename + "' needs a packages_path");
    return packages_path + "/" + filename.substr(pk.size());
  }
  if (filename.compare(0, fl.size(), fl) == 0) return filename.substr(fl.size());
  if (!filename.empty() && filename[0] == '/') return filename;
  return urdf_dir.empty() ? filename : urdf_dir + "/" + filename;
}

inline void read_stl(const std::string& data, std::vector<double>& out) {
  // binary: 80-byte header, uint32 triangle count, 50 bytes per triangle -- recognised by its exact size
  if (data.size() >= 84) {
    uint32_t n;
    std::memcpy(&n, data.data() + 80, 4);
    if (data.size() == 84 + (size_t)n * 50) {
      for (uint32_t t = 0; t < n; ++t) {
        const char* tri = data.data() + 84 + (size_t)t * 50 + 12;
        for (int v = 0; v < 9; ++v) { float f; std::memcpy(&f, tri + 4 * v, 4); out.push_back((double)f); }
      }
      return;
    }
  }
  std::istringstream is(data);
  std::string tok;
  while (is >> tok) {
    if (tok == "vertex") { double x, y, z; if (!(is >> x >> y >> z)) throw std::runtime_error("stl: malformed vertex"); out.push_back(x); out.push_back(y); out.push_back(z); }
  }
}
inline void read_obj(const std::string& data, std::vector<double>& out) {
  std::istringstream is(data);
  std::string line;
  while (std::getline(is, line)) {
    if (line.size() > 2 && line[0] == 'v' && (line[1] == ' ' || line[1] == '\t')) {
      std::istringstream ls(line.substr(2));
      double x, y, z;
      if (ls >> x >> y >> z) { out.push_back(x); out.push_back(y); out.push_back(z); }
    }
  }
}
inline void collect(const xml::Node* n, const char* tag, std::vector<const xml::Node*>& out) {
  if (n->tag == tag) out.push_back(n);
  for (const auto& c : n->children) collect(c.get(), tag, out);
}
inline void read_dae(const std::string& data, std::vector<double>& out) {
  std::unique_ptr<xml::Node> root = xml::parse(data);
  double unit = 1.0;
  std::vector<const xml::Node*> units, verts, sources;
  collect(root.get(), "unit", units);
  if (!units.empty() && units[0]->attr("meter")) unit = std::atof(units[0]->attr("meter")->c_str());
  collect(root.get(), "vertices", verts);
  collect(root.get(), "source", sources);
  for (const xml::Node* v : verts)
    for (const xml::Node* in : v->all("input")) {
      if (in->attr_or("semantic", "") != "POSITION") continue;
      std::string ref = in->attr_or("source", "");
      if (!ref.empty() && ref[0] == '#') ref = ref.substr(1);
      for (const xml::Node* s : sources) {
        if (s->attr_or("id", "") != ref) continue;
        const xml::Node* fa = s->child("float_array");
        if (!fa) continue;
        std::istringstream is(fa->text);
        double x;
        while (is >> x) out.push_back(unit * x);
      }
    }
  if (out.size() % 3) out.resize(out.size() - out.size() % 3);
}

// vertex coordinates (xyz triples) of a mesh file
inline std::vector<double> read_vertices(const std::string& path, const std::string& data) {
  std::vector<double> v;
  const std::string e = lower_ext(path);
  if (e == "stl") read_stl(data, v);
  else if (e == "obj") read_obj(data, v);
  else if (e == "dae") read_dae(data, v);
  else throw std::runtime_error("urdf: unsupported mesh format '" + path + "' (stl, obj, dae)");
  if (v.size() < 12) throw std::runtime_error("urdf: mesh '" + path + "' has fewer than 4 vertices");
  return v;
}

struct Hull {
  std::vector<double> verts;  // about `centre`
  double centre[3];
  double radius;              // max |v|
};

// Hull vertex set of a point cloud (see the header comment).  `pts` are xyz triples, already scaled.
inline Hull convex_hull(const std::vector<double>& pts) {
  const int n0 = (int)pts.size() / 3;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int i = 0; i < n0; ++i)
    for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], pts[3 * i + a]); hi[a] = std::max(hi[a], pts[3 * i + a]); }
  Hull H;
  for (int a = 0; a < 3; ++a) H.centre[a] = 0.5 * (lo[a] + hi[a]);
  const double size = std::max({hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]});
  if (!(size > 0)) throw std::runtime_error("urdf: degenerate mesh (zero extent)");
  // unique vertices about the centre (duplicates are the rule: every triangle repeats its corners)
  std::vector<std::array<double, 3>> P;
  P.reserve(n0);
  for (int i = 0; i < n0; ++i) P.push_back({pts[3 * i] - H.centre[0], pts[3 * i + 1] - H.centre[1], pts[3 * i + 2] - H.centre[2]});
  std::sort(P.begin(), P.end());
  P.erase(std::unique(P.begin(), P.end()), P.end());
  const int n = (int)P.size();
  std::vector<char> sel(n, 0);
  auto extreme = [&](double dx, double dy, double dz) {
    int bi = 0;
    double bv = -1e300;
    for (int i = 0; i < n; ++i) { const double t = dx * P[i][0] + dy * P[i][1] + dz * P[i][2]; if (t > bv) { bv = t; bi = i; } }
    sel[bi] = 1;
  };
  const int fan = 512;
  for (int k = 0; k < fan; ++k) {  // Fibonacci sphere
    const double z = 1.0 - 2.0 * (k + 0.5) / fan, r = std::sqrt(std::max(0.0, 1.0 - z * z)), phi = k * 2.399963229728653;
    extreme(r * std::cos(phi), r * std::sin(phi), z);
  }
  for (int a = 0; a < 3; ++a) { double d[3] = {0, 0, 0}; d[a] = 1; extreme(d[0], d[1], d[2]); extreme(-d[0], -d[1], -d[2]); }
  // repair: while some vertex lies outside the hull of the selection, add the vertex that is extreme along the direction in
  // which it sticks out (always a vertex of the true hull, so interior points never enter the selection)
  const double tol = 1e-9 * size;
  for (int pass = 0; pass < 256; ++pass) {
    std::vector<double> S;
    for (int i = 0; i < n; ++i) if (sel[i]) { S.push_back(P[i][0]); S.push_back(P[i][1]); S.push_back(P[i][2]); }
    Prim hull;
    hull.type = kConvex; hull.r = size; hull.h = 0; hull.hb = v3(0, 0, 0); hull.c = v3(0, 0, 0); hull.a = v3(0, 0, 1); hull.R = identity3();
    hull.verts = S.data(); hull.nvert = (int)S.size() / 3;
    int added = 0;
    for (int i = 0; i < n; ++i) {
      if (sel[i]) continue;
      Prim pt;
      pt.type = kSphere; pt.r = 0; pt.h = 0; pt.hb = v3(0, 0, 0); pt.c = v3(P[i][0], P[i][1], P[i][2]); pt.a = v3(0, 0, 1); pt.R = identity3();
      pt.verts = nullptr; pt.nvert = 0;
      GjkOut g;
      gjk_distance(pt, hull, 1e-12, 256, g);
      if (g.intersect || !(g.dist > tol)) continue;
      const Vec3 d = g.pa - g.pb;  // from the hull's closest point to the vertex
      int bi = i;
      double bv = -1e300;
      for (int k = 0; k < n; ++k) { const double t = d.x * P[k][0] + d.y * P[k][1] + d.z * P[k][2]; if (t > bv) { bv = t; bi = k; } }
      if (!sel[bi]) { sel[bi] = 1; ++added; }
    }
    if (!added) break;
  }
  H.radius = 0;
  for (int i = 0; i < n; ++i)
    if (sel[i]) {
      H.verts.push_back(P[i][0]); H.verts.push_back(P[i][1]); H.verts.push_back(P[i][2]);
      H.radius = std::max(H.radius, std::sqrt(P[i][0] * P[i][0] + P[i][1] * P[i][1] + P[i][2] * P[i][2]));
    }
  return H;
}

}  // namespace mesh
}  // namespace drc
