/* drc_b200 -- C ABI of the B200 batched control-cycle engine.
 *
 * Drop-in boundary for the per-cycle hot path of YoungWook0533/dyros_robot_controller: each entry
 * point names the reference interface it replaces (file:line under the reference checkout).  The
 * reference's C++ classes / Python `drc` package keep their method names and forward here
 * (INTEGRATION.md shows the bindings).  No C++/torch types cross this boundary: plain pointers,
 * sizes and opaque handles; every function returns 0 on success or a negative DRC_E_* code, with a
 * message available from drc_last_error().  No exceptions propagate.
 *
 * Batched arrays: element (b, k) of a (B, K) array
 *     DRC_LAYOUT_AOS  -> ptr[b*K + k]     batch-major, what numpy (B, K) gives
 *     DRC_LAYOUT_SOA  -> ptr[k*B + b]     component-major, coalesced for one-robot-per-thread kernels
 * Matrices are flattened row-major into K (M: n*n, J: 6*n).  Poses are K = 12: the top three rows of
 * the 4x4 homogeneous matrix, row-major (R00 R01 R02 px R10 ...).  All reals are fp64.
 *
 * drc_batch_* : DEVICE pointers, asynchronous on `stream` (a cudaStream_t passed as void*, NULL = the
 *               context's own stream).  The context owns only handles and scratch.
 * drc_host_*  : HOST pointers (AoS), synchronous: H2D copy, the same kernels, D2H copy.
 */
#ifndef DRC_B200_H
#define DRC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct drc_model drc_model_t;
typedef struct drc_ctx drc_ctx_t;

enum { DRC_OK = 0, DRC_E_INVALID = -1, DRC_E_IO = -2, DRC_E_PARSE = -3, DRC_E_UNSUPPORTED = -4, DRC_E_CUDA = -5,
       DRC_E_NOMEM = -6 };
enum { DRC_LAYOUT_AOS = 0, DRC_LAYOUT_SOA = 1 };
enum { DRC_DRIVE_DIFFERENTIAL = 0, DRC_DRIVE_MECANUM = 1, DRC_DRIVE_CASTER = 2 };
/* per-robot QP status (OSQP's vocabulary); only DRC_QP_SOLVED counts as success, as in QP_base.h:167 */
enum { DRC_QP_UNSOLVED = 0, DRC_QP_SOLVED = 1, DRC_QP_MAX_ITER = 2, DRC_QP_PRIMAL_INFEASIBLE = 3,
       DRC_QP_DUAL_INFEASIBLE = 4, DRC_QP_NON_CONVEX = 5, DRC_QP_SOLVED_INACCURATE = 6 };

/* Constants the reference hard-codes (src/manipulator/QP_IK.cpp:81-86,101,122,130,
 * src/manipulator/robot_controller.cpp:12-15, include/math_type_define.h:7) and the OSQP settings it
 * leaves at their defaults (include/dyros_robot_controller/QP_base.h:146-149). */
typedef struct drc_params {
  double alpha, slack_weight, ik_reg, moma_ik_reg, mani_thresh, dist_thresh;
  double Kp_task[6], Kv_task[6];
  double Kp_joint[16], Kv_joint[16];
  double rho, sigma, osqp_alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
  int max_iter, check_termination, scaling, adaptive_rho, adaptive_rho_interval;
  double adaptive_rho_tolerance;
  double gjk_tol, epa_tol;
  int gjk_max_iter, epa_max_iter;
  double pinv_threshold;
  /* scheduling only, results unaffected: order the ADMM launch by each robot's iteration count of the previous call on
   * this context (longest first, similar counts share a warp) -- consecutive control ticks solve nearly the same QP */
  int schedule_hint;
  /* drc_batch_rollout_qpik: 0 (default) = per tick, the multi-stream pipeline of drc_batch_cycle_qpik_step (priority pipeline for the
   * predicted-slow robots, integrate step fused into the solver launches); 1 = TWO launches per tick (k_tick_front: schedule
   * scatter + cubic profile + FK + narrow phase + EPA + QPIK record; k_admm: solve + integrate + next tick's schedule).  Same results. */
  int rollout_fused;
  /* drc_batch_rollout_qpik: 1 = every tick's QP is warm started from the robot's previous tick (x0, y0 = its primal / dual
   * solution, osqp_warm_start semantics; tick 0 and robots whose previous QP was infeasible start cold).  An EXTENSION for
   * simulation rollouts: the reference builds a fresh solver per cycle and never warm starts (QP_base.h:133-177, warm_start is
   * left at OSQP's cold path), so results with 1 are not the reference's iterates; 0 (default) = the reference's. */
  int rollout_warm_start;
} drc_params_t;

const char* drc_last_error(void);
int drc_version(void);
/* number of CUDA devices visible (0 when no GPU / driver) */
int drc_device_count(void);

/* ---- model: replaces Manipulator::RobotData::RobotData (src/manipulator/robot_data.cpp:7-70), i.e.
 * pinocchio::urdf::buildModel/buildGeom + addAllCollisionPairs + srdf::removeCollisionPairs.  A missing
 * URDF is an error code here (the reference calls std::exit, :15-19); a missing SRDF enables every pair
 * (:45-49).  <mesh> collision elements (:24-34): `package://pkg/file` is looked up under packages_path, other names next
 * to the URDF; STL / OBJ / DAE vertices become the mesh's CONVEX HULL (GJK / EPA support set) -- exact for convex meshes,
 * a lower bound of the distance for non-convex ones (hpp-fcl walks the triangles of a BVH model).  A mesh that cannot be
 * read is an error code.  drc_model_create_from_text has no file context: relative mesh names resolve against the
 * working directory and package:// names are errors. */
int drc_model_create_from_urdf(const char* urdf_path, const char* srdf_path, const char* packages_path,
                               drc_model_t** out);
int drc_model_create_from_text(const char* urdf_text, const char* srdf_text, drc_model_t** out);
void drc_model_destroy(drc_model_t* m);
int drc_model_dof(const drc_model_t* m);                         /* getDof() */
int drc_model_frame_id(const drc_model_t* m, const char* link);  /* model_.getFrameId(); -1 = unknown link */
int drc_model_num_frames(const drc_model_t* m);
const char* drc_model_frame_name(const drc_model_t* m, int frame);
const char* drc_model_joint_name(const drc_model_t* m, int joint);
/* getJointPositionLimit / getJointVelocityLimit (robot_data.cpp:59-62); any pointer may be NULL */
int drc_model_limits(const drc_model_t* m, double* q_lo, double* q_hi, double* v_lim, double* effort);
/* sizes: [0] dof [1] collision geometries [2] enabled pairs [3] link-pair groups [4] frames [5] skipped (unknown) geometry tags */
int drc_model_info(const drc_model_t* m, int* sizes6);
/* mesh collision geometry: [0] geometries built from <mesh> elements [1] hull vertices kept in total */
int drc_model_mesh_info(const drc_model_t* m, int* sizes2);
/* getVerbose() (robot_data.cpp:72-89); returns a string owned by the model */
const char* drc_model_verbose(const drc_model_t* m);

/* ---- context: device state cache + scratch for up to max_batch robots on one GPU */
int drc_ctx_create(const drc_model_t* m, int device, int max_batch, drc_ctx_t** out);
void drc_ctx_destroy(drc_ctx_t* c);
int drc_ctx_get_params(const drc_ctx_t* c, drc_params_t* p);
int drc_ctx_set_params(drc_ctx_t* c, const drc_params_t* p);
int drc_ctx_max_batch(const drc_ctx_t* c);
int drc_ctx_synchronize(drc_ctx_t* c);
void* drc_ctx_stream(drc_ctx_t* c); /* the context's cudaStream_t */

/* ---- RobotData::updateState (robot_data.cpp:91-124): FK, joint Jacobians, CRBA M, M^-1, g, nle -> cache */
int drc_batch_update_state(drc_ctx_t* c, int B, const double* q, const double* qdot, int layout, void* stream);
/* getPose / getJacobian / getJacobianTimeVariation / getVelocity (robot_data.cpp:378-422); NULL = skip.
 * Always the fresh pose oMi[parent]*placement. */
int drc_batch_get_frame(drc_ctx_t* c, int B, int frame, double* pose12, double* J, double* Jdot, double* vel,
                        int layout, void* stream);
/* getMassMatrix / getMassMatrixInv / getGravity / getCoriolis / getNonlinearEffects; NULL = skip */
int drc_batch_get_dynamics(drc_ctx_t* c, int B, double* M, double* Minv, double* g, double* coriolis, double* nle,
                           int layout, void* stream);
/* getManipulability(with_grad, with_graddot, link) (robot_data.cpp:519-573) */
int drc_batch_get_manipulability(drc_ctx_t* c, int B, int frame, int with_graddot, double* mani, double* grad,
                                 double* grad_dot, int layout, void* stream);
/* getMinDistance(with_grad, with_graddot, verbose) (robot_data.cpp:424-517); pair (int, B) may be NULL */
int drc_batch_get_min_distance(drc_ctx_t* c, int B, int with_graddot, double* dist, double* grad, double* grad_dot,
                               int* pair, int layout, void* stream);

/* ---- RobotController (src/manipulator/robot_controller.cpp), all on the cached state
 * QPIK(xdot_target, link) :277-290  -> qdot* (zeros when the QP is not Solved)
 * QPIKStep(x_target, xdot_target, link) :292-300
 * QPID(xddot_target, link) :319-333 -> tau* (gravity when the QP is not Solved); qddot_out optional
 * QPIDStep :335-343 */
int drc_batch_qpik(drc_ctx_t* c, int B, const double* xdot_des, int frame, double* qdot_out, int* status, int* iters,
                   int layout, void* stream);
int drc_batch_qpik_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, int frame,
                        double* qdot_out, int* status, int* iters, int layout, void* stream);
int drc_batch_qpid(drc_ctx_t* c, int B, const double* xddot_des, int frame, double* tau_out, double* qddot_out,
                   int* status, int* iters, int layout, void* stream);
int drc_batch_qpid_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, int frame,
                        double* tau_out, double* qddot_out, int* status, int* iters, int layout, void* stream);
/* CLIKStep :156-171 (null_qdot may be NULL), OSF :208-225, OSFStep :232-240 (null_torque may be NULL),
 * moveJointTorqueStep(q_target, qdot_target) :115-125 */
int drc_batch_clik_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, const double* null_qdot,
                        int frame, double* qdot_out, int layout, void* stream);
int drc_batch_osf(drc_ctx_t* c, int B, const double* xddot_target, const double* null_torque, int frame, double* tau_out,
                  int layout, void* stream);
int drc_batch_osf_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, const double* null_torque,
                       int frame, double* tau_out, int layout, void* stream);
int drc_batch_joint_torque_step(drc_ctx_t* c, int B, const double* q_target, const double* qdot_target, double* tau_out,
                                int layout, void* stream);
/* BASELINE config 2 fused: updateState (src/manipulator/robot_data.cpp:91-124) + CLIKStep (robot_controller.cpp:156-171, no
 * null-space velocity) + OSFStep (:232-240, no null-space torque) in ONE launch; leaves the same state cache as drc_batch_update_state */
int drc_batch_cycle_clik_osf_step(drc_ctx_t* c, int B, const double* q, const double* qdot, const double* x_target,
                                  const double* xdot_target, int frame, double* qdot_out, double* tau_out, int layout, void* stream);
/* DyrosMath::getTaskSpaceCubic (include/math_type_define.h:647-685) for per-robot targets and a common time */
int drc_batch_task_space_cubic(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, const double* x_init,
                               const double* xdot_init, double t, double t0, double duration, double* x_des,
                               double* xdot_des, int layout, void* stream);

/* ---- fused control cycle: updateState + QPIKStep / QPIDStep in one call (the BASELINE.json metric).
 * The state cache is refreshed exactly as by drc_batch_update_state. */
int drc_batch_cycle_qpik_step(drc_ctx_t* c, int B, const double* q, const double* qdot, const double* x_target,
                              const double* xdot_target, int frame, double* qdot_out, int* status, int* iters, int layout,
                              void* stream);
int drc_batch_cycle_qpid_step(drc_ctx_t* c, int B, const double* q, const double* qdot, const double* x_target,
                              const double* xdot_target, int frame, double* tau_out, int* status, int* iters, int layout,
                              void* stream);

/* ---- host-buffer variants (AoS, synchronous): what a CPU caller of the reference classes uses */
/* ---- closed-loop rollout (the caller of the path: examples/C++/src/fr3_controller.cpp:116-131).  T control ticks of
 * updateState + QPIKCubic (duration > 0: DyrosMath::getTaskSpaceCubic from (x_init, xdot_init) at t0 to (x_target,
 * xdot_target) at t0 + duration, evaluated at t_start + k dt; robot_controller.cpp:302-317) or QPIKStep (duration <= 0,
 * x_init / xdot_init may be NULL), each followed by the example's integrate step q_desired = q + qdot_desired dt with ideal
 * tracking (q <- q_desired, qdot <- qdot_desired).  q, qdot are updated IN PLACE; nothing returns to the host between
 * ticks.  fail_ticks / iters_total (int, B; may be NULL): ticks whose QP was not Solved (command 0), total ADMM iterations. */
int drc_batch_rollout_qpik(drc_ctx_t* c, int B, int T, double dt, double* q, double* qdot, const double* x_target,
                           const double* xdot_target, const double* x_init, const double* xdot_init, double t_start, double t0,
                           double duration, int frame, int* fail_ticks, int* iters_total, int layout, void* stream);

int drc_host_update_state(drc_ctx_t* c, int B, const double* q, const double* qdot);
int drc_host_get_frame(drc_ctx_t* c, int B, int frame, double* pose12, double* J, double* Jdot, double* vel);
int drc_host_get_dynamics(drc_ctx_t* c, int B, double* M, double* Minv, double* g, double* coriolis, double* nle);
int drc_host_get_manipulability(drc_ctx_t* c, int B, int frame, int with_graddot, double* mani, double* grad,
                                double* grad_dot);
int drc_host_get_min_distance(drc_ctx_t* c, int B, int with_graddot, double* dist, double* grad, double* grad_dot, int* pair);
int drc_host_qpik(drc_ctx_t* c, int B, const double* xdot_des, int frame, double* qdot_out, int* status, int* iters);
int drc_host_qpik_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, int frame, double* qdot_out,
                       int* status, int* iters);
int drc_host_qpid(drc_ctx_t* c, int B, const double* xddot_des, int frame, double* tau_out, double* qddot_out, int* status,
                  int* iters);
int drc_host_qpid_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, int frame, double* tau_out,
                       double* qddot_out, int* status, int* iters);
int drc_host_clik_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, const double* null_qdot,
                       int frame, double* qdot_out);
int drc_host_osf(drc_ctx_t* c, int B, const double* xddot_target, const double* null_torque, int frame, double* tau_out);
int drc_host_osf_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, const double* null_torque,
                      int frame, double* tau_out);
int drc_host_joint_torque_step(drc_ctx_t* c, int B, const double* q_target, const double* qdot_target, double* tau_out);
int drc_host_task_space_cubic(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, const double* x_init,
                              const double* xdot_init, double t, double t0, double duration, double* x_des, double* xdot_des);
int drc_host_rollout_qpik(drc_ctx_t* c, int B, int T, double dt, double* q, double* qdot, const double* x_target,
                          const double* xdot_target, const double* x_init, const double* xdot_init, double t_start, double t0,
                          double duration, int frame, int* fail_ticks, int* iters_total);
int drc_host_cycle_qpik_step(drc_ctx_t* c, int B, const double* q, const double* qdot, const double* x_target,
                             const double* xdot_target, int frame, double* qdot_out, int* status, int* iters);
int drc_host_cycle_qpid_step(drc_ctx_t* c, int B, const double* q, const double* qdot, const double* x_target,
                             const double* xdot_target, int frame, double* tau_out, int* status, int* iters);
int drc_host_cycle_clik_osf_step(drc_ctx_t* c, int B, const double* q, const double* qdot, const double* x_target,
                                 const double* xdot_target, int frame, double* qdot_out, double* tau_out);

/* ---- mobile base + mobile manipulator (reference src/mobile/robot_data.cpp, src/mobile_manipulator/*.cpp).
 * Attach the base to a compiled URDF model BEFORE creating contexts: replaces the Mobile::RobotData(KinematicParam) and
 * MobileManipulator::RobotData(KinematicParam, JointIndex, ActuatorIndex, urdf...) constructors
 * (mobile/robot_data.cpp:7-32, mobile_manipulator/robot_data.cpp:7-44).  The URDF must hold three virtual joints
 * (prismatic x, prismatic y, revolute z at virtual_start..+2), n_wheels wheel joints at mobi_start and the manipulator
 * joints at mani_start (dof - 3 - n_wheels of them, :19).  Arrays follow KinematicParam (type_define.h:58-72):
 * roller_angles / b2w_angles (n_wheels, mecanum), b2w_x / b2w_y (n_wheels for mecanum, n_wheels/2 for caster). */
int drc_model_attach_mobile_base(drc_model_t* m, int drive_type, double wheel_radius, double base_width, double wheel_offset,
                                 int n_wheels, const double* roller_angles, const double* b2w_x, const double* b2w_y,
                                 const double* b2w_angles, int virtual_start, int mani_start, int mobi_start,
                                 int act_mani_start, int act_mobi_start);
/* sizes: [0] drive type (-1 none) [1] wheels [2] manipulator dof [3] actuated dof */
int drc_model_moma_info(const drc_model_t* m, int* sizes4);
/* Mobile::RobotData::getFKJacobian (mobile/robot_data.cpp:122-177), 3 x n_wheels row-major; differential / mecanum */
int drc_model_base_jacobian(const drc_model_t* m, double* J);
/* MobileManipulator::RobotData::updateState (mobile_manipulator/robot_data.cpp:83-144) on the joint-ordered vectors
 * q = getJointVector(q_virtual, q_mobile, q_mani), qdot likewise (dof each, :417-427). */
int drc_batch_moma_update_state(drc_ctx_t* c, int B, const double* q, const double* qdot, int layout, void* stream);
/* getPose / getJacobianActuated / getJacobianActuatedTimeVariation / getVelocity (:407-415), getMassMatrixActuated,
 * getMassMatrixActuatedInv, getGravityActuated, getNonlinearEffectsActuated (:138-142), getManipulability (manipulator
 * columns, :439-496).  J: 6 x act, M: act x act, mani_grad: manipulator dof.  NULL = skip. */
int drc_batch_moma_get_state(drc_ctx_t* c, int B, int frame, double* pose12, double* J_act, double* Jdot_act, double* vel,
                             double* M_act, double* Minv_act, double* g_act, double* nle_act, double* mani, double* mani_grad,
                             double* mani_graddot, int layout, void* stream);
/* MobileManipulator::RobotController::QPIK / QPIKStep / QPID / QPIDStep (mobile_manipulator/robot_controller.cpp:147-231).
 * eta_out / tau_out / etadot_out hold the ACTUATED vector (act per robot); the reference's (mobile, manipulator) pair is
 * its ActuatorIndex split.  Failures: zeros (QPIK), actuated gravity + zero eta_dot (QPID; the reference slices the
 * full-dof gravity with actuator indices, :208-218, a latent bug that is not reproduced). */
int drc_batch_moma_qpik(drc_ctx_t* c, int B, const double* xdot_des, int frame, double* eta_out, int* status, int* iters,
                        int layout, void* stream);
int drc_batch_moma_qpik_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, int frame,
                             double* eta_out, int* status, int* iters, int layout, void* stream);
int drc_batch_moma_qpid(drc_ctx_t* c, int B, const double* xddot_des, int frame, double* tau_out, double* etadot_out,
                        int* status, int* iters, int layout, void* stream);
int drc_batch_moma_qpid_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, int frame,
                             double* tau_out, double* etadot_out, int* status, int* iters, int layout, void* stream);
/* fused updateState + QPIKStep / QPIDStep (BASELINE configs 4-5) */
int drc_batch_moma_cycle_qpik_step(drc_ctx_t* c, int B, const double* q, const double* qdot, const double* x_target,
                                   const double* xdot_target, int frame, double* eta_out, int* status, int* iters,
                                   int layout, void* stream);
int drc_batch_moma_cycle_qpid_step(drc_ctx_t* c, int B, const double* q, const double* qdot, const double* x_target,
                                   const double* xdot_target, int frame, double* tau_out, double* etadot_out, int* status,
                                   int* iters, int layout, void* stream);
int drc_host_moma_update_state(drc_ctx_t* c, int B, const double* q, const double* qdot);
int drc_host_moma_get_state(drc_ctx_t* c, int B, int frame, double* pose12, double* J_act, double* Jdot_act, double* vel,
                            double* M_act, double* Minv_act, double* g_act, double* nle_act, double* mani, double* mani_grad,
                            double* mani_graddot);
int drc_host_moma_qpik(drc_ctx_t* c, int B, const double* xdot_des, int frame, double* eta_out, int* status, int* iters);
int drc_host_moma_qpik_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, int frame,
                            double* eta_out, int* status, int* iters);
int drc_host_moma_qpid(drc_ctx_t* c, int B, const double* xddot_des, int frame, double* tau_out, double* etadot_out,
                       int* status, int* iters);
int drc_host_moma_qpid_step(drc_ctx_t* c, int B, const double* x_target, const double* xdot_target, int frame,
                            double* tau_out, double* etadot_out, int* status, int* iters);
int drc_host_moma_cycle_qpik_step(drc_ctx_t* c, int B, const double* q, const double* qdot, const double* x_target,
                                  const double* xdot_target, int frame, double* eta_out, int* status, int* iters);
int drc_host_moma_cycle_qpid_step(drc_ctx_t* c, int B, const double* q, const double* qdot, const double* x_target,
                                  const double* xdot_target, int frame, double* tau_out, double* etadot_out, int* status,
                                  int* iters);

/* ---- mobile base alone (reference src/mobile/robot_data.cpp, src/mobile/robot_controller.cpp).  The base has no URDF:
 * the handle holds the KinematicParam (type_define.h:58-72).  n_wheels: 2 (differential), roller_angles.size() (mecanum),
 * 2 * base2wheel_positions.size() (caster: steer, roll per caster).  Replaces Mobile::RobotData::RobotData
 * (mobile/robot_data.cpp:7-32) and Mobile::RobotController::RobotController (mobile/robot_controller.cpp:7-12). */
typedef struct drc_mobile drc_mobile_t;
int drc_mobile_create(int drive_type, double wheel_radius, double base_width, double wheel_offset, double max_lin_speed,
                      double max_ang_speed, double max_lin_acc, double max_ang_acc, int n_wheels, const double* roller_angles,
                      const double* b2w_x, const double* b2w_y, const double* b2w_angles, int device, drc_mobile_t** out);
void drc_mobile_destroy(drc_mobile_t* h);
int drc_mobile_wheel_num(const drc_mobile_t* h);
int drc_mobile_synchronize(drc_mobile_t* h);
long long drc_mobile_launch_count(const drc_mobile_t* h);
/* Mobile::RobotData::updateState / computeBaseVel / computeFKJacobian (mobile/robot_data.cpp:103-204): per base
 * J_fk (3 x w row-major; NULL = skip) and base_vel = J_fk * wheel_vel (3; NULL = skip).  wheel_pos is read for caster
 * bases only (steering angles at the even indices).  DEVICE pointers. */
int drc_batch_mobile_fk(drc_mobile_t* h, int B, const double* wheel_pos, const double* wheel_vel, double* J_fk, double* base_vel,
                        int layout, void* stream);
/* Mobile::RobotController::computeIKJacobian / computeWheelVel / VelocityCommand (mobile/robot_controller.cpp:14-124):
 * J_ik (w x 3 row-major; NULL = skip), wheel_vel = J_ik * v with v = base_vel_des, saturated as VelocityCommand does
 * (planar speed to max_lin_speed along its direction, yaw rate to max_ang_speed) when saturate != 0. */
int drc_batch_mobile_ik(drc_mobile_t* h, int B, const double* wheel_pos, const double* base_vel_des, int saturate, double* J_ik,
                        double* wheel_vel, int layout, void* stream);
/* HOST-pointer siblings (batch-major arrays, synchronous) */
int drc_host_mobile_fk(drc_mobile_t* h, int B, const double* wheel_pos, const double* wheel_vel, double* J_fk, double* base_vel);
int drc_host_mobile_ik(drc_mobile_t* h, int B, const double* wheel_pos, const double* base_vel_des, int saturate, double* J_ik,
                       double* wheel_vel);

/* ---- instrumentation (replaces QP::TimeDuration / SuhanBenchmark, QP_base.h:19-43) */
/* device time in ms of the stages of the LAST cycle/QP call on this context (CUDA events on its stream):
 * [0] joint placements + self-collision narrow phase  [1] state / QP build (EPA pass of the collision stage runs next to
 * it on a side stream)  [2] ADMM solve  [3] total; requires drc_ctx_enable_timing(c,1) */
int drc_ctx_enable_timing(drc_ctx_t* c, int on);
int drc_ctx_last_timing(drc_ctx_t* c, float* ms4);
/* stage trace of the LAST fused cycle call: up to max_marks named events (names joined by ';'), times in ms relative to the first
 * mark, recorded on the streams of the main and the priority pipeline; returns the number of marks; requires timing enabled */
int drc_ctx_last_trace(drc_ctx_t* c, int max_marks, float* ms, char* names, int names_len);
/* debug outputs of the QP solves (test instrumentation for the active-set gate; the reference exposes nothing comparable):
 * after enabling, every QP call also stores, per robot, the primal vector [core x (NC) | unit slacks (KU*NC) | row singletons (NR)]
 * and the UNSCALED dual vector [core bound rows (NC) | unit rows (KU*NC) | unit-slack bound rows (KU*NC) | dense/equality rows (NR)
 * | row-singleton bound rows (NR)] of OSQP's final iterate (y = E y_scaled / c); rows that do not exist in a formulation hold 0.
 * NC = core variables (dof / actuated dof), KU = 2 (QPIK) or 4 (QPID), NR = 2 (QPIK) or 2 + NC (QPID). */
int drc_ctx_enable_qp_debug(drc_ctx_t* c, int on);
int drc_host_get_qp_debug(drc_ctx_t* c, int B, int x_per_robot, int y_per_robot, double* x, double* y);
/* number of kernels this library launched on the context since creation */
long long drc_ctx_launch_count(const drc_ctx_t* c);
/* measured FP64 FMA throughput of the device (TFLOP/s) -- the roofline denominator of this fp64 path */
int drc_bench_fp64_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
