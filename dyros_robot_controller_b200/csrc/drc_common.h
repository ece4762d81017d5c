// drc_b200 -- shared definitions of the B200 batched control-cycle engine.
//
// Every routine that runs inside a kernel is written as a DRC_HD (host+device) inline so that
//   * nvcc compiles it for sm_100a (the only product path), and
//   * tests/kernel_emu compiles the very same bodies with g++ to unit-test kernel logic on the
//     CPU-only build container (test infrastructure; never linked into libdrc_b200.so).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define DRC_HD __host__ __device__ __forceinline__
#define DRC_HD_NOINLINE static __host__ __device__ __noinline__  // static: one private copy per translation unit
#else
#define DRC_HD inline
#define DRC_HD_NOINLINE static inline
#endif

// phase markers for the flop-counting build of the kernel bodies (tools/flopcount); no-ops everywhere else
#ifndef DRC_PHASE
#define DRC_PHASE(x) ((void)0)
#endif

namespace drc {

constexpr int kMaxV = 16;      // max model dof
constexpr int kMaxGeom = 64;   // collision primitives
constexpr int kMaxPair = 512;  // enabled collision pairs
constexpr int kMaxGroup = 64;  // (jointA, jointB) groups of pairs

enum JointType : int { kRevolute = 0, kPrismatic = 1 };
enum GeomType : int { kSphere = 0, kCylinder = 1, kBox = 2, kCapsule = 3, kConvex = 4 };  // kConvex: convex hull of a mesh (vertex set)
enum DriveType : int { kNoBase = -1, kDifferential = 0, kMecanum = 1, kCaster = 2 };

// QP solver / formulation constants.  Defaults are the reference's hard-coded values
// (QP_IK.cpp:81-86,101,122,130; robot_controller.cpp:12-15) and OSQP 0.6 defaults (QP_base.h:146-149).
struct DrcParams {
  // controller
  double alpha = 50.0, slack_weight = 1000.0, ik_reg = 1.0, moma_ik_reg = 0.01;
  double mani_thresh = 0.01, dist_thresh = 0.05;
  double Kp_task[6] = {100, 100, 100, 100, 100, 100};
  double Kv_task[6] = {20, 20, 20, 20, 20, 20};
  double Kp_joint[kMaxV], Kv_joint[kMaxV];
  // OSQP
  double rho = 0.1, sigma = 1e-6, osqp_alpha = 1.6;
  double eps_abs = 1e-3, eps_rel = 1e-3, eps_prim_inf = 1e-4, eps_dual_inf = 1e-4;
  int max_iter = 4000, check_termination = 25, scaling = 10;
  int adaptive_rho = 1, adaptive_rho_interval = 50;
  double adaptive_rho_tolerance = 5.0;
  // narrow phase
  double gjk_tol = 1e-10, epa_tol = 1e-6;  // EPA: hpp-fcl's default; curved pairs converge like 1/k^2
  int gjk_max_iter = 128, epa_max_iter = 96;
  double pinv_threshold = 1e-6;  // COD rank threshold (math_type_define.h:7)
  // scheduling only (no effect on results): order the ADMM launch by the previous tick's iteration counts
  int schedule_hint = 1;
  // closed-loop rollouts: 0 = the multi-stream pipeline of the fused cycle per tick (fastest), 1 = two launches per tick (k_tick_front + k_admm)
  int rollout_fused = 0;
  // closed-loop rollouts: 1 = warm start every tick's QP from the previous tick's primal / dual solution (osqp_warm_start
  // semantics).  An extension: the reference never warm starts (fresh solver per cycle, QP_base.h:146), so 0 = its iterates.
  int rollout_warm_start = 0;
};

// Collision primitives (kept as one block so a kernel can stage it into shared memory).
struct GeomTable {
  int type[kMaxGeom];
  int parent[kMaxGeom];
  double prm[kMaxGeom][3];  // sphere r | cylinder/capsule r, half length | box half extents
  double R[kMaxGeom][9];    // placement in the parent joint frame
  double p[kMaxGeom][3];
  unsigned char pair_a[kMaxPair], pair_b[kMaxPair];  // geometry indices of the enabled pairs, sorted by group
  short pair_id[kMaxPair];                           // index in the reference's pair order (tie-break)
  // mesh collision geometry (robot_data.cpp:24-34, packages_path branch): the convex hull's vertices, in the geometry's own
  // frame about its placement point `p`; `hull` is host memory in the host model and device memory in a context's copy
  double brad[kMaxGeom];                             // bounding radius about p (every type)
  int vert_off[kMaxGeom], vert_n[kMaxGeom];          // kConvex: vertices hull[3*vert_off ..], vert_n of them
  const double* hull;
};

// Flat, fixed-topology robot model.  Passed to kernels BY VALUE as a __grid_constant__ parameter
// (constant bank: uniform broadcast reads, private to each launch, no cross-stream hazards).
struct DrcModelDev {
  int nv;
  int parent[kMaxV];         // -1 = universe
  int jtype[kMaxV];
  unsigned anc_mask[kMaxV];  // bit j: joint j is i itself or an ancestor of i
  double axis[kMaxV][3];     // joint axis in the joint frame
  double jR[kMaxV][9];       // placement in the parent joint frame (row-major R, then p)
  double jp[kMaxV][3];
  double mass[kMaxV];
  double com[kMaxV][3];      // joint frame
  double inertia[kMaxV][6];  // about the com, joint-frame axes: xx xy xz yy yz zz
  double q_lo[kMaxV], q_hi[kMaxV], v_lim[kMaxV];
  double gravity[3];
  // collision geometry, grouped by (parent joint A, parent joint B)
  int ngeom;
  GeomTable geom;
  int npair, ngroup, ngjk;
  unsigned short gjk_pair[64];   // i-th pair (sorted order) that needs GJK (cylinder/box vs cylinder/box)
  short group_ja[kMaxGroup], group_jb[kMaxGroup], group_first[kMaxGroup], group_count[kMaxGroup];
  // mobile manipulator (drive_type == kNoBase for a plain arm)
  int drive_type, wheel_num, virtual_start, mani_start, mobi_start, act_mani_start, act_mobi_start, mani_dof;
  double J_mobile[3][8];  // constant base Jacobian for differential / mecanum drives (3 x wheel_num)
  double wheel_radius, wheel_offset;
  double b2w_x[4], b2w_y[4];  // caster steering-axis positions
};

// Target frame of a controller call (a BODY frame of the URDF: parent joint + fixed placement).
struct DrcFrame {
  int parent;  // joint index, -1 = universe
  double R[9];
  double p[3];
};

// Addressing of a batched array of K-vectors: element (b, k) at ptr[b*sb + k*sk].
//   SoA  [K][B]: sb = 1, sk = B      (coalesced for thread-per-robot kernels; device entry points)
//   AoS  [B][K]: sb = K, sk = 1      (numpy-natural; host entry points)
struct Strided {
  long long sb, sk;
};
DRC_HD Strided soa(long long B) { return Strided{1, B}; }
DRC_HD Strided aos(long long K) { return Strided{K, 1}; }

enum QpStatus : int {
  kQpUnsolved = 0, kQpSolved = 1, kQpMaxIter = 2, kQpPrimalInfeasible = 3, kQpDualInfeasible = 4, kQpNonConvex = 5,
  kQpSolvedInaccurate = 6
};

}  // namespace drc
