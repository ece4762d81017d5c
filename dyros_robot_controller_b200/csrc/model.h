// drc_b200 -- host-side model compiler: URDF (+SRDF) text -> flat DrcModelDev blob.
//
// Replaces what the reference's RobotData constructor obtains from pinocchio::urdf::buildModel,
// buildGeom, addAllCollisionPairs and srdf::removeCollisionPairs
// (reference src/manipulator/robot_data.cpp:7-70); conventions listed in SURVEY.md 8(a) row a1.
#pragma once
#include <string>
#include <vector>

#include "drc_common.h"

namespace drc {

struct HostFrame {
  std::string name;
  int parent;  // joint index, -1 = universe
  double R[9], p[3];
};

struct HostModel {
  DrcModelDev dev;
  std::string name;
  std::vector<std::string> joint_names;
  std::vector<HostFrame> frames;
  std::vector<std::string> geom_names;
  std::vector<double> effort;
  bool chain = true;  // parent[i] == i-1 for every joint
  int skipped_geoms = 0;  // collision elements that are not geometry this library knows (none of the URDF primitive / mesh tags)
  int mesh_geoms = 0;     // <mesh> collision elements, turned into convex hulls (kConvex)
  std::vector<double> hull;  // their hull vertices (xyz triples, geometry frame); dev.geom.hull points here on the host
  void bind_hull() { dev.geom.hull = hull.empty() ? nullptr : hull.data(); }  // call after the model has reached its final address
  int frame_id(const std::string& n) const {
    for (size_t i = 0; i < frames.size(); ++i) if (frames[i].name == n) return (int)i;
    return -1;
  }
};

// Where <mesh filename="..."> collision elements are looked up (robot_data.cpp:24-34): `package://` under packages_path,
// relative names under the URDF's directory.
struct MeshSource {
  std::string urdf_dir, packages_path;
};
// Throws std::runtime_error with a readable message on malformed or unsupported input.
HostModel compile_model(const std::string& urdf_text, const std::string& srdf_text, const MeshSource& meshes = MeshSource());

// Mobile-base extension (reference MobileManipulator::RobotData ctor, mobile_manipulator/robot_data.cpp:7-44
// and Mobile::RobotData, mobile/robot_data.cpp:7-32,138-177).
struct MobileParam {
  int drive_type;  // DriveType
  double wheel_radius, base_width, wheel_offset;
  std::vector<double> roller_angles, b2w_x, b2w_y, b2w_angles;
  double max_lin_speed = 0, max_ang_speed = 0, max_lin_acc = 0, max_ang_acc = 0;
};
// differential / mecanum forward (3 x w) and inverse (w x 3) base Jacobians; zero for caster bases (state dependent)
void mobile_constant_jacobians(const MobileParam& p, int w, double J_fk[3][8], double J_ik[8][3]);
void attach_mobile_base(HostModel& m, const MobileParam& p, int virtual_start, int mani_start, int mobi_start,
                        int act_mani_start, int act_mobi_start);

std::string read_text_file(const std::string& path);  // throws if missing

}  // namespace drc
