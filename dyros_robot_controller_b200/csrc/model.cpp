// drc_b200 -- URDF/SRDF model compiler (host).  See model.h.
#include "model.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>

#include "mesh_hull.h"
#include "xml_mini.h"

namespace drc {
namespace {

struct Rot {
  double m[9];
};
struct Tf {
  Rot R;
  double p[3];
};

Rot rot_identity() { Rot r = {{1, 0, 0, 0, 1, 0, 0, 0, 1}}; return r; }
Rot rot_mul(const Rot& a, const Rot& b) {
  Rot c;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c.m[3 * i + j] = a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j] + a.m[3 * i + 2] * b.m[6 + j];
  return c;
}
void rot_apply(const Rot& a, const double* v, double* o) {
  for (int i = 0; i < 3; ++i) o[i] = a.m[3 * i] * v[0] + a.m[3 * i + 1] * v[1] + a.m[3 * i + 2] * v[2];
}
// URDF fixed-axis roll-pitch-yaw: R = Rz(yaw) Ry(pitch) Rx(roll)
Rot rot_rpy(double r, double p, double y) {
  const double cr = std::cos(r), sr = std::sin(r), cp = std::cos(p), sp = std::sin(p), cy = std::cos(y), sy = std::sin(y);
  Rot R = {{cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
            sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
            -sp,     cp * sr,                cp * cr}};
  return R;
}
Tf tf_identity() { Tf t; t.R = rot_identity(); t.p[0] = t.p[1] = t.p[2] = 0; return t; }
Tf tf_mul(const Tf& a, const Tf& b) {
  Tf c;
  c.R = rot_mul(a.R, b.R);
  rot_apply(a.R, b.p, c.p);
  for (int i = 0; i < 3; ++i) c.p[i] += a.p[i];
  return c;
}

std::vector<double> numbers(const std::string& s) {
  std::vector<double> v;
  std::istringstream is(s);
  double d;
  while (is >> d) v.push_back(d);
  return v;
}
double number(const xml::Node* n, const char* key, double dflt, bool required = false) {
  const std::string* a = n ? n->attr(key) : nullptr;
  if (!a) {
    if (required) throw std::runtime_error(std::string("urdf: missing attribute '") + key + "'");
    return dflt;
  }
  std::vector<double> v = numbers(*a);
  if (v.size() != 1) throw std::runtime_error(std::string("urdf: attribute '") + key + "' is not a number");
  return v[0];
}
Tf origin_of(const xml::Node* parent) {
  Tf t = tf_identity();
  const xml::Node* o = parent ? parent->child("origin") : nullptr;
  if (!o) return t;
  std::vector<double> xyz = numbers(o->attr_or("xyz", "0 0 0")), rpy = numbers(o->attr_or("rpy", "0 0 0"));
  if (xyz.size() != 3 || rpy.size() != 3) throw std::runtime_error("urdf: malformed <origin>");
  t.R = rot_rpy(rpy[0], rpy[1], rpy[2]);
  for (int i = 0; i < 3; ++i) t.p[i] = xyz[i];
  return t;
}

struct Builder {
  HostModel hm;
  MeshSource meshes;
  std::map<std::string, const xml::Node*> links;
  std::map<std::string, std::vector<const xml::Node*>> child_joints;
  // accumulated body inertia about each joint origin
  std::vector<double> mass;
  std::vector<std::array<double, 3>> first_moment;
  std::vector<std::array<double, 9>> inertia_o;
  std::vector<std::string> geom_link;

  void add_inertia(int j, const Tf& T, const xml::Node* link) {
    const xml::Node* ine = link->child("inertial");
    if (!ine || j < 0) return;  // bodies welded to the universe carry no dynamics
    const Tf Ti = tf_mul(T, origin_of(ine));
    const double mval = number(ine->child("mass"), "value", 0.0, true);
    const xml::Node* it = ine->child("inertia");
    if (!it) throw std::runtime_error("urdf: <inertial> without <inertia>");
    const double ixx = number(it, "ixx", 0), ixy = number(it, "ixy", 0), ixz = number(it, "ixz", 0),
                 iyy = number(it, "iyy", 0), iyz = number(it, "iyz", 0), izz = number(it, "izz", 0);
    const Rot I = {{ixx, ixy, ixz, ixy, iyy, iyz, ixz, iyz, izz}};
    Rot Rt;
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) Rt.m[3 * a + b] = Ti.R.m[3 * b + a];
    const Rot Ic = rot_mul(rot_mul(Ti.R, I), Rt);
    const double* c = Ti.p;
    const double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    mass[j] += mval;
    for (int a = 0; a < 3; ++a) first_moment[j][a] += mval * c[a];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) inertia_o[j][3 * a + b] += Ic.m[3 * a + b] + mval * ((a == b ? cc : 0.0) - c[a] * c[b]);
  }

  void add_link(int j, const Tf& T, const std::string& lname) {
    auto it = links.find(lname);
    if (it == links.end()) throw std::runtime_error("urdf: joint references unknown link '" + lname + "'");
    const xml::Node* link = it->second;
    HostFrame f;
    f.name = lname; f.parent = j;
    std::memcpy(f.R, T.R.m, sizeof f.R);
    std::memcpy(f.p, T.p, sizeof f.p);
    hm.frames.push_back(f);
    add_inertia(j, T, link);
    int k = 0;
    for (const xml::Node* c : link->all("collision")) {
      const int idx = k++;
      const xml::Node* geo = c->child("geometry");
      if (!geo || geo->children.empty()) continue;
      const xml::Node* g = geo->children[0].get();
      int type = -1;
      double prm[3] = {0, 0, 0};
      if (g->tag == "sphere") { type = kSphere; prm[0] = number(g, "radius", 0, true); }
      else if (g->tag == "cylinder") { type = kCylinder; prm[0] = number(g, "radius", 0, true); prm[1] = 0.5 * number(g, "length", 0, true); }
      else if (g->tag == "capsule") { type = kCapsule; prm[0] = number(g, "radius", 0, true); prm[1] = 0.5 * number(g, "length", 0, true); }
      else if (g->tag == "box") {
        std::vector<double> s = numbers(g->attr_or("size", ""));
        if (s.size() != 3) throw std::runtime_error("urdf: malformed <box size>");
        type = kBox;
        for (int a = 0; a < 3; ++a) prm[a] = 0.5 * s[a];
      }
      mesh::Hull hull;
      if (g->tag == "mesh") {  // the convex hull of the mesh's vertices
        const std::string path = mesh::resolve(g->attr_or("filename", ""), meshes.urdf_dir, meshes.packages_path);
        std::vector<double> pts = mesh::read_vertices(path, read_text_file(path));
        std::vector<double> sc = numbers(g->attr_or("scale", "1 1 1"));
        if (sc.size() != 3) throw std::runtime_error("urdf: malformed <mesh scale>");
        for (size_t i = 0; i < pts.size(); ++i) pts[i] *= sc[i % 3];
        hull = mesh::convex_hull(pts);
        type = kConvex;
        ++hm.mesh_geoms;
      } else if (g->tag != "sphere" && g->tag != "cylinder" && g->tag != "capsule" && g->tag != "box") { ++hm.skipped_geoms; continue; }
      DrcModelDev& d = hm.dev;
      if (d.ngeom >= kMaxGeom) throw std::runtime_error("urdf: too many collision primitives (max 64)");
      Tf Tg = tf_mul(T, origin_of(c));
      const int gi = d.ngeom++;
      d.geom.vert_off[gi] = 0; d.geom.vert_n[gi] = 0;
      if (type == kConvex) {  // vertices are kept about the hull's box centre: move the placement point there
        double off[3];
        rot_apply(Tg.R, hull.centre, off);
        for (int a = 0; a < 3; ++a) Tg.p[a] += off[a];
        d.geom.vert_off[gi] = (int)hm.hull.size() / 3;
        d.geom.vert_n[gi] = (int)hull.verts.size() / 3;
        hm.hull.insert(hm.hull.end(), hull.verts.begin(), hull.verts.end());
        prm[0] = hull.radius;
      }
      d.geom.brad[gi] = type == kSphere ? prm[0] : type == kCylinder ? std::sqrt(prm[0] * prm[0] + prm[1] * prm[1])
                        : type == kCapsule ? prm[0] + prm[1] : type == kBox ? std::sqrt(prm[0] * prm[0] + prm[1] * prm[1] + prm[2] * prm[2]) : hull.radius;
      d.geom.type[gi] = type;
      d.geom.parent[gi] = j;
      std::memcpy(d.geom.prm[gi], prm, sizeof prm);
      std::memcpy(d.geom.R[gi], Tg.R.m, sizeof Tg.R.m);
      std::memcpy(d.geom.p[gi], Tg.p, sizeof Tg.p);
      hm.geom_names.push_back(lname + "_" + std::to_string(idx));
      geom_link.push_back(lname);
    }
  }

  void visit(const std::string& lname, int j, const Tf& T) {
    add_link(j, T, lname);
    for (const xml::Node* jn : child_joints[lname]) {
      const Tf Tn = tf_mul(T, origin_of(jn));
      const std::string type = jn->attr_or("type", "");
      const xml::Node* ch = jn->child("child");
      if (!ch || !ch->attr("link")) throw std::runtime_error("urdf: joint without <child link>");
      const std::string child = *ch->attr("link");
      if (type == "fixed") { visit(child, j, Tn); continue; }
      if (type != "revolute" && type != "continuous" && type != "prismatic")
        throw std::runtime_error("urdf: unsupported joint type '" + type + "'");
      DrcModelDev& d = hm.dev;
      if (d.nv >= kMaxV) throw std::runtime_error("urdf: too many joints (max 16)");
      const int nj = d.nv++;
      hm.joint_names.push_back(jn->attr_or("name", ""));
      d.parent[nj] = j;
      d.jtype[nj] = type == "prismatic" ? kPrismatic : kRevolute;
      std::vector<double> ax = numbers(jn->child("axis") ? jn->child("axis")->attr_or("xyz", "1 0 0") : "1 0 0");
      if (ax.size() != 3) throw std::runtime_error("urdf: malformed <axis>");
      const double an = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
      if (an == 0) throw std::runtime_error("urdf: zero joint axis");
      for (int a = 0; a < 3; ++a) d.axis[nj][a] = ax[a] / an;
      std::memcpy(d.jR[nj], Tn.R.m, sizeof Tn.R.m);
      std::memcpy(d.jp[nj], Tn.p, sizeof Tn.p);
      const xml::Node* lim = jn->child("limit");
      const double inf = 1e300;
      const bool bounded = lim && type != "continuous";
      d.q_lo[nj] = bounded ? number(lim, "lower", -inf) : -inf;
      d.q_hi[nj] = bounded ? number(lim, "upper", inf) : inf;
      d.v_lim[nj] = lim ? number(lim, "velocity", inf) : inf;
      hm.effort.push_back(lim ? number(lim, "effort", inf) : inf);
      mass.push_back(0.0);
      first_moment.push_back({0, 0, 0});
      inertia_o.push_back({0, 0, 0, 0, 0, 0, 0, 0, 0});
      visit(child, nj, tf_identity());
    }
  }
};

}  // namespace

std::string read_text_file(const std::string& path) {
  std::ifstream f(path.c_str(), std::ios::in | std::ios::binary);
  if (!f) throw std::runtime_error("cannot open '" + path + "'");
  std::ostringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

HostModel compile_model(const std::string& urdf_text, const std::string& srdf_text, const MeshSource& meshes) {
  std::unique_ptr<xml::Node> root = xml::parse(urdf_text);
  if (root->tag != "robot") throw std::runtime_error("urdf: root element is not <robot>");
  Builder B;
  B.meshes = meshes;
  std::memset(&B.hm.dev, 0, sizeof(DrcModelDev));
  B.hm.name = root->attr_or("name", "");
  std::vector<std::string> order;
  for (const xml::Node* l : root->all("link")) {
    const std::string n = l->attr_or("name", "");
    B.links[n] = l;
    order.push_back(n);
  }
  std::set<std::string> has_parent;
  for (const xml::Node* j : root->all("joint")) {
    const xml::Node *p = j->child("parent"), *c = j->child("child");
    if (!p || !c || !p->attr("link") || !c->attr("link")) throw std::runtime_error("urdf: joint without parent/child link");
    B.child_joints[*p->attr("link")].push_back(j);
    has_parent.insert(*c->attr("link"));
  }
  std::vector<std::string> roots;
  for (auto& n : order) if (!has_parent.count(n)) roots.push_back(n);
  if (roots.size() != 1) throw std::runtime_error("urdf: expected exactly one root link");
  B.visit(roots[0], -1, tf_identity());

  DrcModelDev& d = B.hm.dev;
  if (d.nv == 0) throw std::runtime_error("urdf: model has no moving joints");
  for (int i = 0; i < d.nv; ++i) {
    d.mass[i] = B.mass[i];
    double c[3] = {0, 0, 0};
    if (B.mass[i] > 0) for (int a = 0; a < 3; ++a) c[a] = B.first_moment[i][a] / B.mass[i];
    const double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    double Ic[9];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) Ic[3 * a + b] = B.inertia_o[i][3 * a + b] - B.mass[i] * ((a == b ? cc : 0.0) - c[a] * c[b]);
    for (int a = 0; a < 3; ++a) d.com[i][a] = c[a];
    d.inertia[i][0] = Ic[0]; d.inertia[i][1] = Ic[1]; d.inertia[i][2] = Ic[2];
    d.inertia[i][3] = Ic[4]; d.inertia[i][4] = Ic[5]; d.inertia[i][5] = Ic[8];
    unsigned mask = 0;
    for (int k = i; k >= 0; k = d.parent[k]) mask |= 1u << k;
    d.anc_mask[i] = mask;
    if (d.parent[i] != i - 1) B.hm.chain = false;
  }
  d.gravity[0] = 0; d.gravity[1] = 0; d.gravity[2] = -9.81;  // Pinocchio's default model gravity
  d.drive_type = kNoBase;

  // collision pairs: all (i<j) with different parent joints, minus SRDF-disabled link pairs
  std::set<std::pair<std::string, std::string>> disabled;
  if (!srdf_text.empty()) {
    std::unique_ptr<xml::Node> sroot = xml::parse(srdf_text);
    for (const xml::Node* dc : sroot->all("disable_collisions")) {
      std::string a = dc->attr_or("link1", ""), b = dc->attr_or("link2", "");
      disabled.insert({a, b});
      disabled.insert({b, a});
    }
  }
  struct P { int a, b, id; };
  std::vector<P> pairs;
  int id = 0, ngjk = 0;
  for (int i = 0; i < d.ngeom; ++i)
    for (int j = i + 1; j < d.ngeom; ++j) {
      if (d.geom.parent[i] == d.geom.parent[j]) continue;
      if (disabled.count({B.geom_link[i], B.geom_link[j]})) continue;
      pairs.push_back({i, j, id++});
      const int ta = d.geom.type[i], tb = d.geom.type[j];
      if (!has_closed_form(ta, tb)) ++ngjk;
    }
  if ((int)pairs.size() > kMaxPair) throw std::runtime_error("model: too many collision pairs (max 512)");
  if (ngjk > 64) throw std::runtime_error("model: more than 64 collision pairs need GJK (cylinder / box / mesh hull against each other)");
  // group by (parent joint A, parent joint B), stable in reference order
  // ... and inside a group by the second geometry, so that the narrow phase places it once for all its partners
  std::stable_sort(pairs.begin(), pairs.end(), [&](const P& x, const P& y) {
    const int xa = d.geom.parent[x.a], xb = d.geom.parent[x.b], ya = d.geom.parent[y.a], yb = d.geom.parent[y.b];
    if (xa != ya) return xa < ya;
    if (xb != yb) return xb < yb;
    return x.b < y.b;
  });
  d.npair = (int)pairs.size();
  d.ngroup = 0;
  for (int k = 0; k < d.npair; ++k) {
    d.geom.pair_a[k] = (unsigned char)pairs[k].a;
    d.geom.pair_b[k] = (unsigned char)pairs[k].b;
    d.geom.pair_id[k] = (short)pairs[k].id;
    const int ja = d.geom.parent[pairs[k].a], jb = d.geom.parent[pairs[k].b];
    if (d.ngroup == 0 || d.group_ja[d.ngroup - 1] != ja || d.group_jb[d.ngroup - 1] != jb) {
      if (d.ngroup >= kMaxGroup) throw std::runtime_error("model: too many link pairs (max 64)");
      d.group_ja[d.ngroup] = (short)ja; d.group_jb[d.ngroup] = (short)jb;
      d.group_first[d.ngroup] = (short)k; d.group_count[d.ngroup] = 0;
      ++d.ngroup;
    }
    d.group_count[d.ngroup - 1]++;
  }
  d.ngjk = 0;
  for (int k = 0; k < d.npair; ++k) {
    const int ta = d.geom.type[d.geom.pair_a[k]], tb = d.geom.type[d.geom.pair_b[k]];
    if (!has_closed_form(ta, tb)) d.gjk_pair[d.ngjk++] = (unsigned short)k;
  }
  B.hm.bind_hull();
  return B.hm;
}

// Configuration-independent base Jacobians (w <= 8 wheels):
//   forward  J_fk (3 x w): DifferentialFKJacobian / MecanumFKJacobian = PinvCOD(J_inv)   mobile/robot_data.cpp:138-177
//   inverse  J_ik (w x 3): DifferentialIKJacobian / MecanumIKJacobian                    mobile/robot_controller.cpp:65-102
void mobile_constant_jacobians(const MobileParam& p, int w, double J_fk[3][8], double J_ik[8][3]) {
  std::memset(J_fk, 0, sizeof(double) * 3 * 8);
  std::memset(J_ik, 0, sizeof(double) * 8 * 3);
  if (p.drive_type == kDifferential) {
    J_fk[0][0] = p.wheel_radius / 2; J_fk[0][1] = p.wheel_radius / 2;
    J_fk[2][0] = -p.wheel_radius / p.base_width; J_fk[2][1] = p.wheel_radius / p.base_width;
    J_ik[0][0] = 1 / p.wheel_radius; J_ik[0][2] = -p.base_width / (2 * p.wheel_radius);
    J_ik[1][0] = 1 / p.wheel_radius; J_ik[1][2] = p.base_width / (2 * p.wheel_radius);
  } else if (p.drive_type == kMecanum) {
    for (int i = 0; i < w; ++i) {
      const double r = p.wheel_radius, g = p.roller_angles[i], px = p.b2w_x[i], py = p.b2w_y[i], pt = p.b2w_angles[i];
      const double a3[2] = {1.0, std::tan(g)};
      const double a2[2][2] = {{std::cos(pt), std::sin(pt)}, {-std::sin(pt), std::cos(pt)}};
      const double a1[2][3] = {{1, 0, -py}, {0, 1, px}};
      for (int c = 0; c < 3; ++c) {
        double s = 0;
        for (int k = 0; k < 2; ++k) for (int l = 0; l < 2; ++l) s += a3[k] * a2[k][l] * a1[l][c];
        J_ik[i][c] = s / r;
      }
    }
    // full column rank (3): pinv = (Ji' Ji)^-1 Ji'
    double G[9] = {0};
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int i = 0; i < w; ++i) G[3 * a + b] += J_ik[i][a] * J_ik[i][b];
    const double det = G[0] * (G[4] * G[8] - G[5] * G[7]) - G[1] * (G[3] * G[8] - G[5] * G[6]) + G[2] * (G[3] * G[7] - G[4] * G[6]);
    if (std::fabs(det) < 1e-12) throw std::runtime_error("mobile base: mecanum inverse Jacobian is rank deficient");
    double Gi[9];
    Gi[0] = (G[4] * G[8] - G[5] * G[7]) / det; Gi[1] = (G[2] * G[7] - G[1] * G[8]) / det; Gi[2] = (G[1] * G[5] - G[2] * G[4]) / det;
    Gi[3] = (G[5] * G[6] - G[3] * G[8]) / det; Gi[4] = (G[0] * G[8] - G[2] * G[6]) / det; Gi[5] = (G[2] * G[3] - G[0] * G[5]) / det;
    Gi[6] = (G[3] * G[7] - G[4] * G[6]) / det; Gi[7] = (G[1] * G[6] - G[0] * G[7]) / det; Gi[8] = (G[0] * G[4] - G[1] * G[3]) / det;
    for (int a = 0; a < 3; ++a)
      for (int i = 0; i < w; ++i) {
        double s = 0;
        for (int b = 0; b < 3; ++b) s += Gi[3 * a + b] * J_ik[i][b];
        J_fk[a][i] = s;
      }
  }
}

void attach_mobile_base(HostModel& m, const MobileParam& p, int virtual_start, int mani_start, int mobi_start,
                        int act_mani_start, int act_mobi_start) {
  DrcModelDev& d = m.dev;
  int w = 0;
  if (p.drive_type == kDifferential) w = 2;
  else if (p.drive_type == kMecanum) w = (int)p.roller_angles.size();
  else if (p.drive_type == kCaster) w = 2 * (int)p.b2w_x.size();
  else throw std::runtime_error("mobile base: unknown drive type");
  if (w <= 0 || w > 8) throw std::runtime_error("mobile base: unsupported wheel count");
  if (p.drive_type == kMecanum && (p.b2w_x.size() != (size_t)w || p.b2w_y.size() != (size_t)w || p.b2w_angles.size() != (size_t)w))
    throw std::runtime_error("mobile base: mecanum parameter arrays differ in length");
  const int mani = d.nv - 3 - w;  // mobile_manipulator/robot_data.cpp:19
  if (mani <= 0) throw std::runtime_error("mobile base: URDF has no manipulator joints left (dof - 3 - wheels <= 0)");
  d.drive_type = p.drive_type; d.wheel_num = w; d.mani_dof = mani;
  d.virtual_start = virtual_start; d.mani_start = mani_start; d.mobi_start = mobi_start;
  d.act_mani_start = act_mani_start; d.act_mobi_start = act_mobi_start;
  d.wheel_radius = p.wheel_radius; d.wheel_offset = p.wheel_offset;
  std::memset(d.J_mobile, 0, sizeof d.J_mobile);
  if (p.drive_type != kCaster) {
    double Jik[8][3];
    mobile_constant_jacobians(p, w, d.J_mobile, Jik);
  } else {
    // powered casters: J_mobile depends on the steering angles and is evaluated per state (drc_mobile.h).  One caster
    // alone leaves Jp' Jp rank deficient (the reference would return a reduced-rank pseudo-inverse): not supported.
    if (p.b2w_x.size() < 2 || p.b2w_x.size() > 4 || p.b2w_y.size() != p.b2w_x.size())
      throw std::runtime_error("mobile base: caster drives need 2..4 casters with base2wheel_positions for each");
    if (!(p.wheel_offset > 0) || !(p.wheel_radius > 0)) throw std::runtime_error("mobile base: caster drives need wheel_offset > 0 and wheel_radius > 0");
    for (size_t i = 0; i < p.b2w_x.size(); ++i) { d.b2w_x[i] = p.b2w_x[i]; d.b2w_y[i] = p.b2w_y[i]; }
  }
}

}  // namespace drc
