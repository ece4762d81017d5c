// drc_b200 -- pybind11 extension module `dyros_robot_controller_cpp_wrapper`.
//
// Same module name, class names and method names as the reference's Boost.Python module (reference src/bindings.cpp:219-447), so
// that the reference's Python package `drc` (which subclasses these classes, drc/manipulator/robot_data.py:6) and user code written
// against it import and run unmodified.  Implementation: every method forwards to the C ABI of libdrc_b200.so
// (include/drc_b200.h) with a batch of ONE robot -- the reference's single-robot semantics; numpy arrays in and out where the
// reference converts Eigen types (VectorXd / MatrixXd <-> ndarray, Affine3d <-> 4x4 ndarray :104-112,192-217, pair<VectorXd,
// VectorXd> -> tuple :93-102, MM QP results -> (mobile, manipulator) tuples :34-90).  No Eigen, no Pinocchio, no OSQP: the
// arithmetic runs in the CUDA kernels behind the C ABI (there is no CPU fallback).
// The batched siblings live in the Python engine (dyros_robot_controller_b200.engine); this module is the compat face.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <array>
#include <cmath>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/drc_b200.h"

namespace py = pybind11;
using arr = py::array_t<double, py::array::c_style | py::array::forcecast>;
using vecd = std::vector<double>;

static void chk(int rc, const char* where) {
  if (rc != DRC_OK) throw std::runtime_error(std::string(where) + ": " + drc_last_error());
}
static arr to_vec(const vecd& v) { return arr(v.size(), v.data()); }
static arr to_mat(const vecd& v, int r, int c) { return arr({r, c}, v.data()); }
static vecd from(const arr& a, size_t n, const char* what) {
  if ((size_t)a.size() != n) throw std::runtime_error(std::string(what) + ": expected " + std::to_string(n) + " values, got " + std::to_string(a.size()));
  return vecd(a.data(), a.data() + n);
}
static vecd pose12(const arr& T) {  // 4x4 (or 3x4) homogeneous matrix -> top three rows
  if (T.ndim() != 2 || T.shape(1) != 4 || T.shape(0) < 3) throw std::runtime_error("pose must be a 4x4 matrix");
  return vecd(T.data(), T.data() + 12);
}
static arr pose44(const vecd& p) {
  vecd T(16, 0.0);
  std::copy(p.begin(), p.begin() + 12, T.begin());
  T[15] = 1.0;
  return to_mat(T, 4, 4);
}
static vecd cubic(double t, double t0, double tf, const vecd& x0, const vecd& xf, const vecd& v0, const vecd& vf, bool dot) {
  // DyrosMath::cubicVector / cubicDotVector (include/math_type_define.h:62-144)
  vecd out(x0.size());
  for (size_t i = 0; i < x0.size(); ++i) {
    if (t < t0) { out[i] = dot ? v0[i] : x0[i]; continue; }
    if (t > tf) { out[i] = dot ? vf[i] : xf[i]; continue; }
    const double e = t - t0, T = tf - t0;
    const double a2 = 3 * (xf[i] - x0[i]) / (T * T) - 2 * v0[i] / T - vf[i] / T, a3 = -2 * (xf[i] - x0[i]) / (T * T * T) + (v0[i] + vf[i]) / (T * T);
    out[i] = dot ? v0[i] + 2 * a2 * e + 3 * a3 * e * e : x0[i] + v0[i] * e + a2 * e * e + a3 * e * e * e;
  }
  return out;
}

// ------------------------------------------------------------------------------------------------ records (type_define.h:13-171)
enum class DriveType : int { Differential = 0, Mecanum = 1, Caster = 2 };
struct KinematicParam {
  DriveType type = DriveType::Differential;
  double wheel_radius = 0, max_lin_speed = 2, max_ang_speed = 2, max_lin_acc = 2, max_ang_acc = 2, base_width = 0, wheel_offset = 0;
  vecd roller_angles, base2wheel_angles;
  std::vector<std::array<double, 2>> base2wheel_positions;
  int wheels() const {
    return type == DriveType::Differential ? 2 : (type == DriveType::Mecanum ? (int)roller_angles.size() : 2 * (int)base2wheel_positions.size());
  }
};
struct MinDistResult {
  double distance = 0;
  arr grad, grad_dot;
  void setZero(int n) { distance = 0; grad = to_vec(vecd(n, 0.0)); grad_dot = to_vec(vecd(n, 0.0)); }
};
struct ManipulabilityResult {
  double manipulability = 0;
  arr grad, grad_dot;
  void setZero(int n) { manipulability = 0; grad = to_vec(vecd(n, 0.0)); grad_dot = to_vec(vecd(n, 0.0)); }
};
struct JointIndex { int virtual_start = 0, mani_start = 0, mobi_start = 0; };
struct ActuatorIndex { int mani_start = 0, mobi_start = 0; };

// ------------------------------------------------------------------------------------------------ Mobile::RobotData
class MobileRD {
 public:
  explicit MobileRD(const KinematicParam& p) : param_(p), w_(p.wheels()) {
    vecd bx, by;
    for (auto& v : p.base2wheel_positions) { bx.push_back(v[0]); by.push_back(v[1]); }
    const vecd ra = p.roller_angles.empty() ? vecd(w_, 0.0) : p.roller_angles, ba = p.base2wheel_angles.empty() ? vecd(w_, 0.0) : p.base2wheel_angles;
    if (bx.empty()) { bx.assign(w_, 0.0); by.assign(w_, 0.0); }
    chk(drc_mobile_create((int)p.type, p.wheel_radius, p.base_width, p.wheel_offset, p.max_lin_speed, p.max_ang_speed, p.max_lin_acc, p.max_ang_acc,
                          w_, ra.data(), bx.data(), by.data(), ba.data(), 0, &h_), "Mobile::RobotData");
    wp_.assign(w_, 0.0); wv_.assign(w_, 0.0); bv_.assign(3, 0.0); J_.assign(3 * w_, 0.0);
    fk(wp_, wv_, &J_, &bv_);
  }
  virtual ~MobileRD() { if (h_) drc_mobile_destroy(h_); }
  MobileRD(const MobileRD&) = delete;
  void fk(const vecd& wp, const vecd& wv, vecd* J, vecd* bv) const {
    chk(drc_host_mobile_fk(h_, 1, wp.data(), wv.data(), J ? J->data() : nullptr, bv ? bv->data() : nullptr), "drc_host_mobile_fk");
  }
  void ik(const vecd& wp, const vecd* v, bool saturate, vecd* J, vecd* wv) const {
    chk(drc_host_mobile_ik(h_, 1, wp.data(), v ? v->data() : nullptr, saturate ? 1 : 0, J ? J->data() : nullptr, wv ? wv->data() : nullptr),
        "drc_host_mobile_ik");
  }
  bool updateMobile(const arr& wheel_pos, const arr& wheel_vel) {  // mobile/robot_data.cpp:104-112
    wp_ = from(wheel_pos, w_, "wheel_pos"); wv_ = from(wheel_vel, w_, "wheel_vel");
    fk(wp_, wv_, &J_, &bv_);
    return true;
  }
  arr computeBaseVel(const arr& wheel_pos, const arr& wheel_vel) const {
    vecd bv(3);
    fk(from(wheel_pos, w_, "wheel_pos"), from(wheel_vel, w_, "wheel_vel"), nullptr, &bv);
    return to_vec(bv);
  }
  arr computeFKJacobian(const arr& wheel_pos) const {
    vecd J(3 * w_);
    fk(from(wheel_pos, w_, "wheel_pos"), vecd(w_, 0.0), &J, nullptr);
    return to_mat(J, 3, w_);
  }
  std::string mobileVerbose() const {
    return "Mobile base: type " + std::to_string((int)param_.type) + ", " + std::to_string(w_) + " wheels, wheel radius " + std::to_string(param_.wheel_radius) + "\n";
  }
  KinematicParam param_;
  int w_;
  drc_mobile_t* h_ = nullptr;
  vecd wp_, wv_, bv_, J_;
};

// ------------------------------------------------------------------------------------------------ Manipulator::RobotData
class ManipRD {
 public:
  ManipRD(const std::string& urdf, const std::string& srdf, const std::string& packages) {
    load(urdf, srdf, packages);
    make_ctx();
  }
  virtual ~ManipRD() {
    if (sc_) drc_ctx_destroy(sc_);
    if (c_) drc_ctx_destroy(c_);
    if (m_) drc_model_destroy(m_);
  }
  ManipRD(const ManipRD&) = delete;

  std::string getVerbose() const { return drc_model_verbose(m_); }
  bool updateState(const arr& q, const arr& qdot) {  // robot_data.cpp:91-99
    q_ = from(q, n_, "q"); qd_ = from(qdot, n_, "qdot");
    chk(drc_host_update_state(c_, 1, q_.data(), qd_.data()), "updateState");
    return true;
  }
  int getDof() const { return n_; }
  arr getJointPosition() const { return to_vec(q_); }
  arr getJointVelocity() const { return to_vec(qd_); }
  py::tuple getJointPositionLimit() const { return py::make_tuple(to_vec(lo_), to_vec(hi_)); }
  py::tuple getJointVelocityLimit() const {
    vecd neg(n_);
    for (int i = 0; i < n_; ++i) neg[i] = -vl_[i];
    return py::make_tuple(to_vec(neg), to_vec(vl_));
  }
  // cached getters (robot_data.h:248-280)
  arr getMassMatrix() { return dyn(c_, 0); }
  arr getMassMatrixInv() { return dyn(c_, 1); }
  arr getGravity() { return dyn(c_, 2); }
  arr getCoriolis() { return dyn(c_, 3); }
  arr getNonlinearEffects() { return dyn(c_, 4); }
  py::object getPose(const std::string& link) { return frame(c_, link, 0); }
  py::object getJacobian(const std::string& link) { return frame(c_, link, 1); }
  py::object getJacobianTimeVariation(const std::string& link) { return frame(c_, link, 2); }
  py::object getVelocity(const std::string& link) { return frame(c_, link, 3); }
  MinDistResult getMinDistance(bool with_grad, bool with_graddot, bool /*verbose*/ = false) { return mindist(c_, with_grad, with_graddot); }
  virtual ManipulabilityResult getManipulability(bool with_grad, bool with_graddot, const std::string& link) { return mani(c_, with_grad, with_graddot, link); }
  // stateless twins (robot_data.cpp:128-374): evaluated on a private scratch context, the cache stays untouched
  arr computeMassMatrix(const arr& q) { return dyn(at(q, nullptr), 0); }
  arr computeGravity(const arr& q) { return dyn(at(q, nullptr), 2); }
  arr computeCoriolis(const arr& q, const arr& qd) { return dyn(at(q, &qd), 3); }
  arr computeNonlinearEffects(const arr& q, const arr& qd) { return dyn(at(q, &qd), 4); }
  py::object computePose(const arr& q, const std::string& link) { return frame(at(q, nullptr), link, 0); }
  py::object computeJacobian(const arr& q, const std::string& link) { return frame(at(q, nullptr), link, 1); }
  py::object computeJacobianTimeVariation(const arr& q, const arr& qd, const std::string& link) { return frame(at(q, &qd), link, 2); }
  py::object computeVelocity(const arr& q, const arr& qd, const std::string& link) { return frame(at(q, &qd), link, 3); }
  MinDistResult computeMinDistance(const arr& q, const arr& qd, bool with_grad, bool with_graddot, bool /*verbose*/ = false) {
    return mindist(at(q, &qd), with_grad, with_graddot);
  }
  ManipulabilityResult computeManipulability(const arr& q, const arr& qd, bool with_grad, bool with_graddot, const std::string& link) {
    return mani(at(q, &qd), with_grad, with_graddot, link);
  }

  // ---- plumbing shared with the controllers
  int frame_id(const std::string& link, bool loud = true) const {
    const int f = drc_model_frame_id(m_, link.c_str());
    if (f < 0 && loud) std::cerr << "\033[1;31mError: Link name " << link << " not found in URDF.\033[0m" << std::endl;  // robot_data.cpp:380-384
    return f;
  }
  drc_ctx_t* ctx() const { return c_; }
  drc_model_t* model() const { return m_; }

 protected:
  struct Deferred {};
  explicit ManipRD(Deferred) {}
  void load(const std::string& urdf, const std::string& srdf, const std::string& packages) {
    chk(drc_model_create_from_urdf(urdf.c_str(), srdf.c_str(), packages.c_str(), &m_), "RobotData");
    n_ = drc_model_dof(m_);
    lo_.assign(n_, 0.0); hi_.assign(n_, 0.0); vl_.assign(n_, 0.0);
    chk(drc_model_limits(m_, lo_.data(), hi_.data(), vl_.data(), nullptr), "drc_model_limits");
    q_.assign(n_, 0.0); qd_.assign(n_, 0.0);
  }
  void make_ctx() { chk(drc_ctx_create(m_, 0, 1, &c_), "RobotData (context)"); }
  virtual void set_state(drc_ctx_t* c, const vecd& q, const vecd& qd) { chk(drc_host_update_state(c, 1, q.data(), qd.data()), "updateState"); }
  drc_ctx_t* at(const arr& q, const arr* qd) {
    if (!sc_) chk(drc_ctx_create(m_, 0, 1, &sc_), "RobotData (scratch context)");
    set_state(sc_, from(q, n_, "q"), qd ? from(*qd, n_, "qdot") : vecd(n_, 0.0));
    return sc_;
  }
  arr dyn(drc_ctx_t* c, int what) {
    vecd out(what < 2 ? n_ * n_ : n_);
    double* p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    p[what] = out.data();
    chk(drc_host_get_dynamics(c, 1, p[0], p[1], p[2], p[3], p[4]), "drc_host_get_dynamics");
    return what < 2 ? to_mat(out, n_, n_) : to_vec(out);
  }
  py::object frame(drc_ctx_t* c, const std::string& link, int what) {
    const int f = frame_id(link);
    if (f < 0) {  // neutral values of the reference (robot_data.cpp:380-399)
      if (what == 0) return pose44({1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0});
      if (what == 3) return to_vec(vecd(6, 0.0));
      return to_mat(vecd(6 * n_, 0.0), 6, n_);
    }
    vecd out(what == 0 ? 12 : (what == 3 ? 6 : 6 * n_));
    double* p[4] = {nullptr, nullptr, nullptr, nullptr};
    p[what] = out.data();
    chk(drc_host_get_frame(c, 1, f, p[0], p[1], p[2], p[3]), "drc_host_get_frame");
    if (what == 0) return pose44(out);
    if (what == 3) return to_vec(out);
    return to_mat(out, 6, n_);
  }
  MinDistResult mindist(drc_ctx_t* c, bool with_grad, bool with_graddot) {
    MinDistResult r;
    r.setZero(n_);
    vecd g(n_), gd(n_);
    chk(drc_host_get_min_distance(c, 1, with_graddot ? 1 : 0, &r.distance, g.data(), gd.data(), nullptr), "drc_host_get_min_distance");
    if (with_grad || with_graddot) r.grad = to_vec(g);
    if (with_graddot) r.grad_dot = to_vec(gd);
    return r;
  }
  ManipulabilityResult mani(drc_ctx_t* c, bool with_grad, bool with_graddot, const std::string& link) {
    ManipulabilityResult r;
    r.setZero(n_);
    const int f = frame_id(link);
    if (f < 0) return r;
    vecd g(n_), gd(n_);
    chk(drc_host_get_manipulability(c, 1, f, with_graddot ? 1 : 0, &r.manipulability, g.data(), gd.data()), "drc_host_get_manipulability");
    if (with_grad || with_graddot) r.grad = to_vec(g);
    if (with_graddot) r.grad_dot = to_vec(gd);
    return r;
  }
  drc_model_t* m_ = nullptr;
  drc_ctx_t *c_ = nullptr, *sc_ = nullptr;
  int n_ = 0;
  vecd q_, qd_, lo_, hi_, vl_;
};

// ------------------------------------------------------------------------------------------------ MobileManipulator::RobotData
class MomaRD : public ManipRD, public MobileRD {
 public:
  MomaRD(const KinematicParam& p, const JointIndex& ji, const ActuatorIndex& ai, const std::string& urdf, const std::string& srdf, const std::string& packages)
      : ManipRD(Deferred{}), MobileRD(p), ji_(ji), ai_(ai) {
    load(urdf, srdf, packages);
    vecd bx, by;
    for (auto& v : p.base2wheel_positions) { bx.push_back(v[0]); by.push_back(v[1]); }
    const vecd ra = p.roller_angles.empty() ? vecd(w_, 0.0) : p.roller_angles, ba = p.base2wheel_angles.empty() ? vecd(w_, 0.0) : p.base2wheel_angles;
    if (bx.empty()) { bx.assign(w_, 0.0); by.assign(w_, 0.0); }
    chk(drc_model_attach_mobile_base(m_, (int)p.type, p.wheel_radius, p.base_width, p.wheel_offset, w_, ra.data(), bx.data(), by.data(), ba.data(),
                                     ji.virtual_start, ji.mani_start, ji.mobi_start, ai.mani_start, ai.mobi_start), "MobileManipulator::RobotData");
    int s[4];
    chk(drc_model_moma_info(m_, s), "drc_model_moma_info");
    k_ = s[2]; act_ = s[3];
    make_ctx();
    qa_.assign(act_, 0.0); qda_.assign(act_, 0.0);
  }
  std::string getVerboseMM() const { return mobileVerbose() + ManipRD::getVerbose(); }
  vecd joint_vector(const vecd& v, const vecd& mo, const vecd& ma) const {  // robot_data.cpp:417-427
    vecd q(n_, 0.0);
    for (int i = 0; i < 3; ++i) q[ji_.virtual_start + i] = v[i];
    for (int i = 0; i < w_; ++i) q[ji_.mobi_start + i] = mo[i];
    for (int i = 0; i < k_; ++i) q[ji_.mani_start + i] = ma[i];
    return q;
  }
  vecd actuator_vector(const vecd& mo, const vecd& ma) const {
    vecd q(act_, 0.0);
    for (int i = 0; i < w_; ++i) q[ai_.mobi_start + i] = mo[i];
    for (int i = 0; i < k_; ++i) q[ai_.mani_start + i] = ma[i];
    return q;
  }
  bool updateState6(const arr& qv, const arr& qmo, const arr& qma, const arr& dv, const arr& dmo, const arr& dma) {  // :83-105
    const vecd mo = from(qmo, w_, "q_mobile"), ma = from(qma, k_, "q_mani"), dmo_ = from(dmo, w_, "qdot_mobile"), dma_ = from(dma, k_, "qdot_mani");
    q_ = joint_vector(from(qv, 3, "q_virtual"), mo, ma);
    qd_ = joint_vector(from(dv, 3, "qdot_virtual"), dmo_, dma_);
    qa_ = actuator_vector(mo, ma); qda_ = actuator_vector(dmo_, dma_);
    updateMobile(qmo, dmo);
    chk(drc_host_moma_update_state(c_, 1, q_.data(), qd_.data()), "updateState");
    return true;
  }
  // ---- full-dof stateless twins with the six-vector signatures (:146-285)
  arr jv(const arr& v, const arr& mo, const arr& ma) const { return to_vec(joint_vector(from(v, 3, "virtual"), from(mo, w_, "mobile"), from(ma, k_, "mani"))); }
  arr cMass(const arr& v, const arr& mo, const arr& ma) { return computeMassMatrix(jv(v, mo, ma)); }
  arr cGrav(const arr& v, const arr& mo, const arr& ma) { return computeGravity(jv(v, mo, ma)); }
  arr cCor(const arr& v, const arr& mo, const arr& ma, const arr& dv, const arr& dmo, const arr& dma) { return computeCoriolis(jv(v, mo, ma), jv(dv, dmo, dma)); }
  arr cNle(const arr& v, const arr& mo, const arr& ma, const arr& dv, const arr& dmo, const arr& dma) { return computeNonlinearEffects(jv(v, mo, ma), jv(dv, dmo, dma)); }
  py::object cPose(const arr& v, const arr& mo, const arr& ma, const std::string& l) { return computePose(jv(v, mo, ma), l); }
  py::object cJac(const arr& v, const arr& mo, const arr& ma, const std::string& l) { return computeJacobian(jv(v, mo, ma), l); }
  py::object cJdot(const arr& v, const arr& mo, const arr& ma, const arr& dv, const arr& dmo, const arr& dma, const std::string& l) {
    return computeJacobianTimeVariation(jv(v, mo, ma), jv(dv, dmo, dma), l);
  }
  py::object cVel(const arr& v, const arr& mo, const arr& ma, const arr& dv, const arr& dmo, const arr& dma, const std::string& l) {
    return computeVelocity(jv(v, mo, ma), jv(dv, dmo, dma), l);
  }
  MinDistResult cMinDist(const arr& v, const arr& mo, const arr& ma, const arr& dv, const arr& dmo, const arr& dma, bool wg, bool wgd, bool verbose = false) {
    return computeMinDistance(jv(v, mo, ma), jv(dv, dmo, dma), wg, wgd, verbose);
  }
  // ---- selection matrix (robot_data.cpp:22-25, 115-120; computeSelectionMatrix :360-376 with the correct wheel block, quirk Q8)
  vecd selection(const vecd& qv, const vecd& qmo) const {
    vecd S((size_t)n_ * act_, 0.0), J(3 * w_);
    fk(qmo, vecd(w_, 0.0), &J, nullptr);
    for (int i = 0; i < k_; ++i) S[(ji_.mani_start + i) * act_ + ai_.mani_start + i] = 1.0;
    for (int i = 0; i < w_; ++i) S[(ji_.mobi_start + i) * act_ + ai_.mobi_start + i] = 1.0;
    const double c = std::cos(qv[2]), s = std::sin(qv[2]);
    for (int kk = 0; kk < w_; ++kk) {
      S[(ji_.virtual_start + 0) * act_ + ai_.mobi_start + kk] = c * J[0 * w_ + kk] - s * J[1 * w_ + kk];
      S[(ji_.virtual_start + 1) * act_ + ai_.mobi_start + kk] = s * J[0 * w_ + kk] + c * J[1 * w_ + kk];
      S[(ji_.virtual_start + 2) * act_ + ai_.mobi_start + kk] = J[2 * w_ + kk];
    }
    return S;
  }
  arr computeSelectionMatrix(const arr& qv, const arr& qmo) const { return to_mat(selection(from(qv, 3, "q_virtual"), from(qmo, w_, "q_mobile")), n_, act_); }
  arr getSelectionMatrix() const {
    return to_mat(selection(vecd(q_.begin() + ji_.virtual_start, q_.begin() + ji_.virtual_start + 3), vecd(q_.begin() + ji_.mobi_start, q_.begin() + ji_.mobi_start + w_)), n_, act_);
  }
  // ---- actuated twins: S(q_virtual)' X(q with q_virtual = 0) (S), as the reference writes them (:185-232, 382-405)
  arr act_twin(int what, const arr& v, const arr& mo, const arr& ma, const arr* dmo, const arr* dma, const std::string* link) {
    const vecd qv = from(v, 3, "q_virtual"), qmo = from(mo, w_, "q_mobile");
    const vecd z3(3, 0.0);
    const vecd q0 = joint_vector(z3, qmo, from(ma, k_, "q_mani"));
    const vecd d0 = joint_vector(z3, dmo ? from(*dmo, w_, "qdot_mobile") : vecd(w_, 0.0), dma ? from(*dma, k_, "qdot_mani") : vecd(k_, 0.0));
    drc_ctx_t* c = at(to_vec(q0), nullptr);
    set_state(c, q0, d0);
    const vecd S = selection(qv, qmo);
    auto St_v = [&](const vecd& x) { vecd o(act_, 0.0); for (int a = 0; a < act_; ++a) for (int i = 0; i < n_; ++i) o[a] += S[i * act_ + a] * x[i]; return o; };
    if (what == 0) {  // S' M S
      vecd M(n_ * n_), T((size_t)n_ * act_, 0.0), out((size_t)act_ * act_, 0.0);
      chk(drc_host_get_dynamics(c, 1, M.data(), nullptr, nullptr, nullptr, nullptr), "get_dynamics");
      for (int i = 0; i < n_; ++i) for (int a = 0; a < act_; ++a) for (int j = 0; j < n_; ++j) T[i * act_ + a] += M[i * n_ + j] * S[j * act_ + a];
      for (int a = 0; a < act_; ++a) for (int b = 0; b < act_; ++b) for (int i = 0; i < n_; ++i) out[a * act_ + b] += S[i * act_ + a] * T[i * act_ + b];
      return to_mat(out, act_, act_);
    }
    if (what <= 3) {  // 1 gravity, 2 coriolis, 3 nle
      vecd g(n_), nle(n_);
      chk(drc_host_get_dynamics(c, 1, nullptr, nullptr, g.data(), nullptr, nle.data()), "get_dynamics");
      if (what == 1) return to_vec(St_v(g));
      if (what == 3) return to_vec(St_v(nle));
      for (int i = 0; i < n_; ++i) nle[i] -= g[i];
      return to_vec(St_v(nle));
    }
    const int f = frame_id(*link);  // 4 J S, 5 Jdot S
    vecd out(6 * act_, 0.0);
    if (f < 0) return to_mat(out, 6, act_);
    vecd J(6 * n_);
    chk(drc_host_get_frame(c, 1, f, nullptr, what == 4 ? J.data() : nullptr, what == 5 ? J.data() : nullptr, nullptr), "get_frame");
    for (int r = 0; r < 6; ++r) for (int a = 0; a < act_; ++a) for (int i = 0; i < n_; ++i) out[r * act_ + a] += J[r * n_ + i] * S[i * act_ + a];
    return to_mat(out, 6, act_);
  }
  arr cMassAct(const arr& v, const arr& mo, const arr& ma) { return act_twin(0, v, mo, ma, nullptr, nullptr, nullptr); }
  arr cGravAct(const arr& v, const arr& mo, const arr& ma) { return act_twin(1, v, mo, ma, nullptr, nullptr, nullptr); }
  arr cCorAct(const arr& v, const arr& mo, const arr& ma, const arr& dmo, const arr& dma) { return act_twin(2, v, mo, ma, &dmo, &dma, nullptr); }
  arr cNleAct(const arr& v, const arr& mo, const arr& ma, const arr& dmo, const arr& dma) { return act_twin(3, v, mo, ma, &dmo, &dma, nullptr); }
  arr cJacAct(const arr& v, const arr& mo, const arr& ma, const std::string& l) { return act_twin(4, v, mo, ma, nullptr, nullptr, &l); }
  arr cJdotAct(const arr& v, const arr& mo, const arr& ma, const arr& /*dv*/, const arr& dmo, const arr& dma, const std::string& l) {
    return act_twin(5, v, mo, ma, &dmo, &dma, &l);
  }
  ManipulabilityResult cMani(const arr& qma, const arr& dma, bool wg, bool wgd, const std::string& link) {  // base at the origin (:287-351)
    const vecd z3(3, 0.0), zw(w_, 0.0);
    drc_ctx_t* c = at(to_vec(joint_vector(z3, zw, from(qma, k_, "q_mani"))), nullptr);
    set_state(c, joint_vector(z3, zw, from(qma, k_, "q_mani")), joint_vector(z3, zw, from(dma, k_, "qdot_mani")));
    return moma_mani(c, wg, wgd, link);
  }
  arr computeMobileFKJacobian(const arr& qmo) const { return computeFKJacobian(qmo); }
  arr computeMobileBaseVel(const arr& qmo, const arr& dmo) const { return computeBaseVel(qmo, dmo); }
  // ---- sizes, indices, joint state
  int getActuatordDof() const { return act_; }
  int getManipulatorDof() const { return k_; }
  int getMobileDof() const { return w_; }
  JointIndex getJointIndex() const { return ji_; }
  ActuatorIndex getActuatorIndex() const { return ai_; }
  arr seg(const vecd& v, int s, int n) const { return to_vec(vecd(v.begin() + s, v.begin() + s + n)); }
  arr getMobileJointPosition() const { return seg(q_, ji_.mobi_start, w_); }
  arr getVirtualJointPosition() const { return seg(q_, ji_.virtual_start, 3); }
  arr getManiJointPosition() const { return seg(q_, ji_.mani_start, k_); }
  arr getMobileJointVelocity() const { return seg(qd_, ji_.mobi_start, w_); }
  arr getVirtualJointVelocity() const { return seg(qd_, ji_.virtual_start, 3); }
  arr getManiJointVelocity() const { return seg(qd_, ji_.mani_start, k_); }
  arr getJointPositionActuated() const { return to_vec(qa_); }
  arr getJointVelocityActuated() const { return to_vec(qda_); }
  // ---- actuated-space cached getters (:126-144, 407-415)
  arr moma_get(int what, const std::string* link) {
    const int f = link ? frame_id(*link) : 0;
    const bool mat = what == 0 || what == 1;  // 0 M~ 1 M~^-1 2 g~ 3 nle~ 4 J~ 5 J~dot
    vecd out(what >= 4 ? 6 * act_ : (mat ? act_ * act_ : act_), 0.0);
    if (f >= 0)
      chk(drc_host_moma_get_state(c_, 1, f, nullptr, what == 4 ? out.data() : nullptr, what == 5 ? out.data() : nullptr, nullptr, what == 0 ? out.data() : nullptr,
                                  what == 1 ? out.data() : nullptr, what == 2 ? out.data() : nullptr, what == 3 ? out.data() : nullptr, nullptr, nullptr, nullptr),
          "drc_host_moma_get_state");
    return what >= 4 ? to_mat(out, 6, act_) : (mat ? to_mat(out, act_, act_) : to_vec(out));
  }
  arr getMassMatrixActuated() { return moma_get(0, nullptr); }
  arr getMassMatrixActuatedInv() { return moma_get(1, nullptr); }
  arr getGravityActuated() { return moma_get(2, nullptr); }
  arr getNonlinearEffectsActuated() { return moma_get(3, nullptr); }
  arr getCoriolisActuated() {
    vecd g(act_), nle(act_);
    chk(drc_host_moma_get_state(c_, 1, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, g.data(), nle.data(), nullptr, nullptr, nullptr), "drc_host_moma_get_state");
    for (int i = 0; i < act_; ++i) nle[i] -= g[i];
    return to_vec(nle);
  }
  arr getJacobianActuated(const std::string& l) { return moma_get(4, &l); }
  arr getJacobianActuatedTimeVariation(const std::string& l) { return moma_get(5, &l); }
  ManipulabilityResult moma_mani(drc_ctx_t* c, bool wg, bool wgd, const std::string& link) {  // manipulator columns only (:439-496)
    ManipulabilityResult r;
    r.setZero(k_);
    const int f = frame_id(link);
    if (f < 0) return r;
    vecd g(k_), gd(k_);
    chk(drc_host_moma_get_state(c, 1, f, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &r.manipulability, g.data(), gd.data()), "drc_host_moma_get_state");
    if (wg || wgd) r.grad = to_vec(g);
    if (wgd) r.grad_dot = to_vec(gd);
    return r;
  }
  ManipulabilityResult getManipulability(bool wg, bool wgd, const std::string& link) override { return moma_mani(c_, wg, wgd, link); }
  arr getMobileFKJacobian() const { return to_mat(J_, 3, w_); }
  arr getMobileBaseVel() const { return to_vec(bv_); }
  int act() const { return act_; }
  int mani_dof() const { return k_; }
  const JointIndex& ji() const { return ji_; }
  const ActuatorIndex& ai() const { return ai_; }
  const vecd& q() const { return q_; }
  const vecd& qd() const { return qd_; }

 protected:
  void set_state(drc_ctx_t* c, const vecd& q, const vecd& qd) override { chk(drc_host_moma_update_state(c, 1, q.data(), qd.data()), "updateState"); }
  JointIndex ji_;
  ActuatorIndex ai_;
  int k_ = 0, act_ = 0;
  vecd qa_, qda_;
};

// ------------------------------------------------------------------------------------------------ controllers
class MobileRC {  // mobile/robot_controller.cpp:7-124
 public:
  MobileRC(double /*dt*/, std::shared_ptr<MobileRD> rd) : rd_(std::move(rd)) {}
  arr computeWheelVel(const arr& base_vel) { vecd v = from(base_vel, 3, "base_vel"), wv(rd_->w_); rd_->ik(rd_->wp_, &v, false, nullptr, &wv); return to_vec(wv); }
  arr computeIKJacobian() { vecd J(3 * rd_->w_); rd_->ik(rd_->wp_, nullptr, false, &J, nullptr); return to_mat(J, rd_->w_, 3); }
  arr VelocityCommand(const arr& desired_base_vel) { vecd v = from(desired_base_vel, 3, "base_vel"), wv(rd_->w_); rd_->ik(rd_->wp_, &v, true, nullptr, &wv); return to_vec(wv); }
 private:
  std::shared_ptr<MobileRD> rd_;
};

class ManipRC {  // manipulator/robot_controller.cpp:7-360
 public:
  ManipRC(double /*dt*/, std::shared_ptr<ManipRD> rd) : rd_(std::move(rd)), n_(rd_->getDof()) {
    kpj_.assign(n_, 400.0); kvj_.assign(n_, 40.0); kpt_.assign(6, 100.0); kvt_.assign(6, 20.0);  // :12-15
    push();
  }
  void setJointGain(const arr& kp, const arr& kv) { kpj_ = from(kp, n_, "Kp"); kvj_ = from(kv, n_, "Kv"); push(); }
  void setJointKpGain(const arr& kp) { kpj_ = from(kp, n_, "Kp"); push(); }
  void setJointKvGain(const arr& kv) { kvj_ = from(kv, n_, "Kv"); push(); }
  void setTaskGain(const arr& kp, const arr& kv) { kpt_ = from(kp, 6, "Kp"); kvt_ = from(kv, 6, "Kv"); push(); }
  void setTaskKpGain(const arr& kp) { kpt_ = from(kp, 6, "Kp"); push(); }
  void setTaskKvGain(const arr& kv) { kvt_ = from(kv, 6, "Kv"); push(); }
  arr moveJointPositionCubic(const arr& qt, const arr& qdt, const arr& qi, const arr& qdi, double t, double t0, double T) {
    return to_vec(cubic(t, t0, t0 + T, from(qi, n_, "q_init"), from(qt, n_, "q_target"), from(qdi, n_, "qdot_init"), from(qdt, n_, "qdot_target"), false));
  }
  arr moveJointVelocityCubic(const arr& qt, const arr& qdt, const arr& qi, const arr& qdi, double t, double t0, double T) {
    return to_vec(cubic(t, t0, t0 + T, from(qi, n_, "q_init"), from(qt, n_, "q_target"), from(qdi, n_, "qdot_init"), from(qdt, n_, "qdot_target"), true));
  }
  arr torqueFromAcc(const arr& qddot) {  // M qddot + g (:108-113)
    vecd a = from(qddot, n_, "qddot_target"), M(n_ * n_), g(n_), tau(n_);
    chk(drc_host_get_dynamics(rd_->ctx(), 1, M.data(), nullptr, g.data(), nullptr, nullptr), "get_dynamics");
    for (int i = 0; i < n_; ++i) { tau[i] = g[i]; for (int j = 0; j < n_; ++j) tau[i] += M[i * n_ + j] * a[j]; }
    return to_vec(tau);
  }
  arr torqueStep(const arr& qt, const arr& qdt) {  // :115-125
    vecd a = from(qt, n_, "q_target"), b = from(qdt, n_, "qdot_target"), tau(n_);
    chk(drc_host_joint_torque_step(rd_->ctx(), 1, a.data(), b.data(), tau.data()), "moveJointTorqueStep");
    return to_vec(tau);
  }
  arr moveJointTorqueCubic(const arr& qt, const arr& qdt, const arr& qi, const arr& qdi, double t, double t0, double T) {
    return torqueStep(moveJointPositionCubic(qt, qdt, qi, qdi, t, t0, T), moveJointVelocityCubic(qt, qdt, qi, qdi, t, t0, T));
  }
  int fid(const std::string& link) const {
    const int f = rd_->frame_id(link);
    if (f < 0) throw std::runtime_error("Link name " + link + " not found in URDF.");
    return f;
  }
  std::pair<vecd, vecd> task_cubic(const arr& xt, const arr& xdt, const arr& xi, const arr& xdi, double t, double t0, double T) {
    vecd a = pose12(xt), b = from(xdt, 6, "xdot_target"), c = pose12(xi), d = from(xdi, 6, "xdot_init"), xd(12), xdd(6);
    chk(drc_host_task_space_cubic(rd_->ctx(), 1, a.data(), b.data(), c.data(), d.data(), t, t0, T, xd.data(), xdd.data()), "getTaskSpaceCubic");
    return {xd, xdd};
  }
  arr clik(const vecd& xt, const vecd& xd, const vecd* nul, const std::string& link) {
    vecd out(n_);
    chk(drc_host_clik_step(rd_->ctx(), 1, xt.data(), xd.data(), nul ? nul->data() : nullptr, fid(link), out.data()), "CLIKStep");
    return to_vec(out);
  }
  arr CLIKStep1(const arr& xt, const arr& xdt, const arr& nul, const std::string& link) { vecd nv = from(nul, n_, "null_qdot"); return clik(pose12(xt), from(xdt, 6, "xdot_target"), &nv, link); }
  arr CLIKStep2(const arr& xt, const arr& xdt, const std::string& link) { return clik(pose12(xt), from(xdt, 6, "xdot_target"), nullptr, link); }
  arr CLIKCubic1(const arr& xt, const arr& xdt, const arr& xi, const arr& xdi, double t, double t0, double T, const arr& nul, const std::string& link) {
    auto d = task_cubic(xt, xdt, xi, xdi, t, t0, T); vecd nv = from(nul, n_, "null_qdot"); return clik(d.first, d.second, &nv, link);
  }
  arr CLIKCubic2(const arr& xt, const arr& xdt, const arr& xi, const arr& xdi, double t, double t0, double T, const std::string& link) {
    auto d = task_cubic(xt, xdt, xi, xdi, t, t0, T); return clik(d.first, d.second, nullptr, link);
  }
  arr osf(const vecd& xdd, const vecd* nul, const std::string& link) {
    vecd out(n_);
    chk(drc_host_osf(rd_->ctx(), 1, xdd.data(), nul ? nul->data() : nullptr, fid(link), out.data()), "OSF");
    return to_vec(out);
  }
  arr OSF1(const arr& xdd, const arr& nul, const std::string& link) { vecd nv = from(nul, n_, "null_torque"); return osf(from(xdd, 6, "xddot_target"), &nv, link); }
  arr OSF2(const arr& xdd, const std::string& link) { return osf(from(xdd, 6, "xddot_target"), nullptr, link); }
  arr osfstep(const vecd& xt, const vecd& xd, const vecd* nul, const std::string& link) {
    vecd out(n_);
    chk(drc_host_osf_step(rd_->ctx(), 1, xt.data(), xd.data(), nul ? nul->data() : nullptr, fid(link), out.data()), "OSFStep");
    return to_vec(out);
  }
  arr OSFStep1(const arr& xt, const arr& xdt, const arr& nul, const std::string& link) { vecd nv = from(nul, n_, "null_torque"); return osfstep(pose12(xt), from(xdt, 6, "xdot_target"), &nv, link); }
  arr OSFStep2(const arr& xt, const arr& xdt, const std::string& link) { return osfstep(pose12(xt), from(xdt, 6, "xdot_target"), nullptr, link); }
  arr OSFCubic1(const arr& xt, const arr& xdt, const arr& xi, const arr& xdi, double t, double t0, double T, const arr& nul, const std::string& link) {
    auto d = task_cubic(xt, xdt, xi, xdi, t, t0, T); vecd nv = from(nul, n_, "null_torque"); return osfstep(d.first, d.second, &nv, link);
  }
  arr OSFCubic2(const arr& xt, const arr& xdt, const arr& xi, const arr& xdi, double t, double t0, double T, const std::string& link) {
    auto d = task_cubic(xt, xdt, xi, xdi, t, t0, T); return osfstep(d.first, d.second, nullptr, link);
  }
  arr report(const vecd& out, int status, const char* what) {
    if (status != DRC_QP_SOLVED) std::cerr << "QP " << what << " failed to compute optimal solution" << std::endl;   // :283-287, 326-330
    return to_vec(out);
  }
  arr QPIK(const arr& xdot, const std::string& link) {
    vecd x = from(xdot, 6, "xdot_target"), out(n_); int st = 0, it = 0;
    chk(drc_host_qpik(rd_->ctx(), 1, x.data(), fid(link), out.data(), &st, &it), "QPIK");
    return report(out, st, "IK");
  }
  arr qpikstep(const vecd& xt, const vecd& xd, const std::string& link) {
    vecd out(n_); int st = 0, it = 0;
    chk(drc_host_qpik_step(rd_->ctx(), 1, xt.data(), xd.data(), fid(link), out.data(), &st, &it), "QPIKStep");
    return report(out, st, "IK");
  }
  arr QPIKStep(const arr& xt, const arr& xdt, const std::string& link) { return qpikstep(pose12(xt), from(xdt, 6, "xdot_target"), link); }
  arr QPIKCubic(const arr& xt, const arr& xdt, const arr& xi, const arr& xdi, double t, double t0, double T, const std::string& link) {
    auto d = task_cubic(xt, xdt, xi, xdi, t, t0, T); return qpikstep(d.first, d.second, link);
  }
  arr QPID(const arr& xddot, const std::string& link) {
    vecd x = from(xddot, 6, "xddot_target"), out(n_); int st = 0, it = 0;
    chk(drc_host_qpid(rd_->ctx(), 1, x.data(), fid(link), out.data(), nullptr, &st, &it), "QPID");
    return report(out, st, "ID");
  }
  arr qpidstep(const vecd& xt, const vecd& xd, const std::string& link) {
    vecd out(n_); int st = 0, it = 0;
    chk(drc_host_qpid_step(rd_->ctx(), 1, xt.data(), xd.data(), fid(link), out.data(), nullptr, &st, &it), "QPIDStep");
    return report(out, st, "ID");
  }
  arr QPIDStep(const arr& xt, const arr& xdt, const std::string& link) { return qpidstep(pose12(xt), from(xdt, 6, "xdot_target"), link); }
  arr QPIDCubic(const arr& xt, const arr& xdt, const arr& xi, const arr& xdi, double t, double t0, double T, const std::string& link) {
    auto d = task_cubic(xt, xdt, xi, xdi, t, t0, T); return qpidstep(d.first, d.second, link);
  }
 private:
  void push() {
    drc_params_t p;
    chk(drc_ctx_get_params(rd_->ctx(), &p), "get_params");
    for (int i = 0; i < n_; ++i) { p.Kp_joint[i] = kpj_[i]; p.Kv_joint[i] = kvj_[i]; }
    for (int i = 0; i < 6; ++i) { p.Kp_task[i] = kpt_[i]; p.Kv_task[i] = kvt_[i]; }
    chk(drc_ctx_set_params(rd_->ctx(), &p), "set_params");
  }
  std::shared_ptr<ManipRD> rd_;
  int n_;
  vecd kpj_, kvj_, kpt_, kvt_;
};

class MomaRC {  // mobile_manipulator/robot_controller.cpp:7-250
 public:
  MomaRC(double /*dt*/, std::shared_ptr<MomaRD> rd) : rd_(std::move(rd)), k_(rd_->mani_dof()), act_(rd_->act()) {
    kpj_.assign(k_, 400.0); kvj_.assign(k_, 40.0); kpt_.assign(6, 400.0); kvt_.assign(6, 40.0);  // :15-18
    push();
  }
  void setManipulatorJointGain(const arr& kp, const arr& kv) { kpj_ = from(kp, k_, "Kp"); kvj_ = from(kv, k_, "Kv"); }
  void setManipulatorJointKpGain(const arr& kp) { kpj_ = from(kp, k_, "Kp"); }
  void setManipulatorJointKvGain(const arr& kv) { kvj_ = from(kv, k_, "Kv"); }
  void setTaskGain(const arr& kp, const arr& kv) { kpt_ = from(kp, 6, "Kp"); kvt_ = from(kv, 6, "Kv"); push(); }
  void setTaskKpGain(const arr& kp) { kpt_ = from(kp, 6, "Kp"); push(); }
  void setTaskKvGain(const arr& kv) { kvt_ = from(kv, 6, "Kv"); push(); }
  arr moveManipulatorJointPositionCubic(const arr& qt, const arr& qdt, const arr& qi, const arr& qdi, double t, double t0, double T) {
    return to_vec(cubic(t, t0, t0 + T, from(qi, k_, "q_init"), from(qt, k_, "q_target"), from(qdi, k_, "qdot_init"), from(qdt, k_, "qdot_target"), false));
  }
  arr torqueFromAcc(const arr& qddot) {  // M_mani qddot + g_mani (:104-110)
    const int n = rd_->getDof(), ms = rd_->ji().mani_start;
    vecd a = from(qddot, k_, "qddot_mani_target"), M(n * n), g(n), tau(k_);
    chk(drc_host_get_dynamics(rd_->ctx(), 1, M.data(), nullptr, g.data(), nullptr, nullptr), "get_dynamics");
    for (int i = 0; i < k_; ++i) { tau[i] = g[ms + i]; for (int j = 0; j < k_; ++j) tau[i] += M[(ms + i) * n + ms + j] * a[j]; }
    return to_vec(tau);
  }
  arr torqueStep(const arr& qt, const arr& qdt) {  // :112-119
    const int ms = rd_->ji().mani_start;
    vecd a = from(qt, k_, "q_mani_target"), b = from(qdt, k_, "qdot_mani_target"), acc(k_);
    for (int i = 0; i < k_; ++i) acc[i] = kpj_[i] * (a[i] - rd_->q()[ms + i]) + kvj_[i] * (b[i] - rd_->qd()[ms + i]);
    return torqueFromAcc(to_vec(acc));
  }
  arr moveManipulatorJointTorqueCubic(const arr& qt, const arr& qdt, const arr& qi, const arr& qdi, double t, double t0, double T) {
    const vecd q0 = from(qi, k_, "q_init"), q1 = from(qt, k_, "q_target"), v0 = from(qdi, k_, "qdot_init"), v1 = from(qdt, k_, "qdot_target");
    return torqueStep(to_vec(cubic(t, t0, t0 + T, q0, q1, v0, v1, false)), to_vec(cubic(t, t0, t0 + T, q0, q1, v0, v1, true)));
  }
  int fid(const std::string& link) const {
    const int f = rd_->frame_id(link);
    if (f < 0) throw std::runtime_error("Link name " + link + " not found in URDF.");
    return f;
  }
  py::tuple split(const vecd& mobile_src, const vecd& mani_src, int status, const char* what) {  // ActuatorIndex split (:162-165, 215-218)
    if (status != DRC_QP_SOLVED) std::cerr << "QP " << what << " failed to compute optimal solution" << std::endl;
    const int w = rd_->getMobileDof();
    vecd mo(mobile_src.begin() + rd_->ai().mobi_start, mobile_src.begin() + rd_->ai().mobi_start + w);
    vecd ma(mani_src.begin() + rd_->ai().mani_start, mani_src.begin() + rd_->ai().mani_start + k_);
    return py::make_tuple(to_vec(mo), to_vec(ma));
  }
  std::pair<vecd, vecd> task_cubic(const arr& xt, const arr& xdt, const arr& xi, const arr& xdi, double t, double t0, double T) {
    vecd a = pose12(xt), b = from(xdt, 6, "xdot_target"), c = pose12(xi), d = from(xdi, 6, "xdot_init"), xd(12), xdd(6);
    chk(drc_host_task_space_cubic(rd_->ctx(), 1, a.data(), b.data(), c.data(), d.data(), t, t0, T, xd.data(), xdd.data()), "getTaskSpaceCubic");
    return {xd, xdd};
  }
  py::tuple QPIK(const arr& xdot, const std::string& link) {
    vecd x = from(xdot, 6, "xdot_target"), out(act_); int st = 0, it = 0;
    chk(drc_host_moma_qpik(rd_->ctx(), 1, x.data(), fid(link), out.data(), &st, &it), "QPIK");
    return split(out, out, st, "IK");
  }
  py::tuple qpikstep(const vecd& xt, const vecd& xd, const std::string& link) {
    vecd out(act_); int st = 0, it = 0;
    chk(drc_host_moma_qpik_step(rd_->ctx(), 1, xt.data(), xd.data(), fid(link), out.data(), &st, &it), "QPIKStep");
    return split(out, out, st, "IK");
  }
  py::tuple QPIKStep(const arr& xt, const arr& xdt, const std::string& link) { return qpikstep(pose12(xt), from(xdt, 6, "xdot_target"), link); }
  py::tuple QPIKCubic(const arr& xt, const arr& xdt, const arr& xi, const arr& xdi, double t, double t0, double T, const std::string& link) {
    auto d = task_cubic(xt, xdt, xi, xdi, t, t0, T); return qpikstep(d.first, d.second, link);
  }
  py::tuple QPID(const arr& xddot, const std::string& link) {
    vecd x = from(xddot, 6, "xddot_target"), tau(act_), ed(act_); int st = 0, it = 0;
    chk(drc_host_moma_qpid(rd_->ctx(), 1, x.data(), fid(link), tau.data(), ed.data(), &st, &it), "QPID");
    return split(ed, tau, st, "ID");
  }
  py::tuple qpidstep(const vecd& xt, const vecd& xd, const std::string& link) {
    vecd tau(act_), ed(act_); int st = 0, it = 0;
    chk(drc_host_moma_qpid_step(rd_->ctx(), 1, xt.data(), xd.data(), fid(link), tau.data(), ed.data(), &st, &it), "QPIDStep");
    return split(ed, tau, st, "ID");
  }
  py::tuple QPIDStep(const arr& xt, const arr& xdt, const std::string& link) { return qpidstep(pose12(xt), from(xdt, 6, "xdot_target"), link); }
  py::tuple QPIDCubic(const arr& xt, const arr& xdt, const arr& xi, const arr& xdi, double t, double t0, double T, const std::string& link) {
    auto d = task_cubic(xt, xdt, xi, xdi, t, t0, T); return qpidstep(d.first, d.second, link);
  }
 private:
  void push() {
    drc_params_t p;
    chk(drc_ctx_get_params(rd_->ctx(), &p), "get_params");
    for (int i = 0; i < 6; ++i) { p.Kp_task[i] = kpt_[i]; p.Kv_task[i] = kvt_[i]; }
    chk(drc_ctx_set_params(rd_->ctx(), &p), "set_params");
  }
  std::shared_ptr<MomaRD> rd_;
  int k_, act_;
  vecd kpj_, kvj_, kpt_, kvt_;
};

// ------------------------------------------------------------------------------------------------ module (bindings.cpp:219-447)
PYBIND11_MODULE(dyros_robot_controller_cpp_wrapper, m) {
  m.doc() = "dyros_robot_controller_cpp_wrapper on libdrc_b200 (B200 batched control-cycle engine, C ABI include/drc_b200.h)";
  py::enum_<DriveType>(m, "DriveType").value("Differential", DriveType::Differential).value("Mecanum", DriveType::Mecanum).value("Caster", DriveType::Caster);
  py::class_<KinematicParam>(m, "KinematicParam")
      .def(py::init<>())
      .def_readwrite("type", &KinematicParam::type).def_readwrite("wheel_radius", &KinematicParam::wheel_radius)
      .def_readwrite("max_lin_speed", &KinematicParam::max_lin_speed).def_readwrite("max_ang_speed", &KinematicParam::max_ang_speed)
      .def_readwrite("max_lin_acc", &KinematicParam::max_lin_acc).def_readwrite("max_ang_acc", &KinematicParam::max_ang_acc)
      .def_readwrite("base_width", &KinematicParam::base_width).def_readwrite("roller_angles", &KinematicParam::roller_angles)
      .def_readwrite("base2wheel_positions", &KinematicParam::base2wheel_positions).def_readwrite("base2wheel_angles", &KinematicParam::base2wheel_angles)
      .def_readwrite("wheel_offset", &KinematicParam::wheel_offset);
  py::class_<MinDistResult>(m, "MinDistResult").def(py::init<>()).def_readwrite("distance", &MinDistResult::distance)
      .def_readwrite("grad", &MinDistResult::grad).def_readwrite("grad_dot", &MinDistResult::grad_dot).def("setZero", &MinDistResult::setZero);
  py::class_<ManipulabilityResult>(m, "ManipulabilityResult").def(py::init<>()).def_readwrite("manipulability", &ManipulabilityResult::manipulability)
      .def_readwrite("grad", &ManipulabilityResult::grad).def_readwrite("grad_dot", &ManipulabilityResult::grad_dot).def("setZero", &ManipulabilityResult::setZero);
  py::class_<JointIndex>(m, "JointIndex").def(py::init<>()).def_readwrite("virtual_start", &JointIndex::virtual_start)
      .def_readwrite("mani_start", &JointIndex::mani_start).def_readwrite("mobi_start", &JointIndex::mobi_start);
  py::class_<ActuatorIndex>(m, "ActuatorIndex").def(py::init<>()).def_readwrite("mani_start", &ActuatorIndex::mani_start)
      .def_readwrite("mobi_start", &ActuatorIndex::mobi_start);

  py::class_<MobileRD, std::shared_ptr<MobileRD>>(m, "MobileRobotData")
      .def(py::init<const KinematicParam&>())
      .def("getVerbose", &MobileRD::mobileVerbose).def("updateState", &MobileRD::updateMobile)
      .def("computeBaseVel", &MobileRD::computeBaseVel).def("computeFKJacobian", &MobileRD::computeFKJacobian)
      .def("getWheelNum", [](const MobileRD& r) { return r.w_; }).def("getKineParam", [](const MobileRD& r) { return r.param_; })
      .def("getWheelPosition", [](const MobileRD& r) { return to_vec(r.wp_); }).def("getWheelVelocity", [](const MobileRD& r) { return to_vec(r.wv_); })
      .def("getBaseVel", [](const MobileRD& r) { return to_vec(r.bv_); }).def("getFKJacobian", [](const MobileRD& r) { return to_mat(r.J_, 3, r.w_); });

  py::class_<ManipRD, std::shared_ptr<ManipRD>>(m, "ManipulatorRobotData")
      .def(py::init<const std::string&, const std::string&, const std::string&>(), py::arg("urdf_path"), py::arg("srdf_path") = "", py::arg("packages_path") = "")
      .def("getVerbose", &ManipRD::getVerbose).def("updateState", &ManipRD::updateState)
      .def("computeMassMatrix", &ManipRD::computeMassMatrix).def("computeGravity", &ManipRD::computeGravity)
      .def("computeCoriolis", &ManipRD::computeCoriolis).def("computeNonlinearEffects", &ManipRD::computeNonlinearEffects)
      .def("computePose", &ManipRD::computePose).def("computeJacobian", &ManipRD::computeJacobian)
      .def("computeJacobianTimeVariation", &ManipRD::computeJacobianTimeVariation).def("computeVelocity", &ManipRD::computeVelocity)
      .def("computeMinDistance", &ManipRD::computeMinDistance, py::arg("q"), py::arg("qdot"), py::arg("with_grad"), py::arg("with_graddot"), py::arg("verbose") = false)
      .def("computeManipulability", &ManipRD::computeManipulability)
      .def("getDof", &ManipRD::getDof).def("getJointPosition", &ManipRD::getJointPosition).def("getJointVelocity", &ManipRD::getJointVelocity)
      .def("getJointPositionLimit", &ManipRD::getJointPositionLimit).def("getJointVelocityLimit", &ManipRD::getJointVelocityLimit)
      .def("getMassMatrix", &ManipRD::getMassMatrix).def("getMassMatrixInv", &ManipRD::getMassMatrixInv).def("getCoriolis", &ManipRD::getCoriolis)
      .def("getGravity", &ManipRD::getGravity).def("getNonlinearEffects", &ManipRD::getNonlinearEffects)
      .def("getPose", &ManipRD::getPose).def("getJacobian", &ManipRD::getJacobian).def("getJacobianTimeVariation", &ManipRD::getJacobianTimeVariation)
      .def("getVelocity", &ManipRD::getVelocity)
      .def("getMinDistance", &ManipRD::getMinDistance, py::arg("with_grad"), py::arg("with_graddot"), py::arg("verbose") = false)
      .def("getManipulability", &ManipRD::getManipulability);

  py::class_<MomaRD, ManipRD, MobileRD, std::shared_ptr<MomaRD>>(m, "MobileManipulatorRobotData")
      .def(py::init<const KinematicParam&, const JointIndex&, const ActuatorIndex&, const std::string&, const std::string&, const std::string&>(),
           py::arg("mobile_param"), py::arg("joint_idx"), py::arg("actuator_idx"), py::arg("urdf_path"), py::arg("srdf_path") = "", py::arg("packages_path") = "")
      .def("getVerbose", &MomaRD::getVerboseMM).def("updateState", &MomaRD::updateState6)
      .def("computeMassMatrix", &MomaRD::cMass).def("computeGravity", &MomaRD::cGrav).def("computeCoriolis", &MomaRD::cCor)
      .def("computeNonlinearEffects", &MomaRD::cNle)
      .def("computeMassMatrixActuated", &MomaRD::cMassAct).def("computeGravityActuated", &MomaRD::cGravAct)
      .def("computeCoriolisActuated", &MomaRD::cCorAct).def("computeNonlinearEffectsActuated", &MomaRD::cNleAct)
      .def("computePose", &MomaRD::cPose).def("computeJacobian", &MomaRD::cJac).def("computeJacobianTimeVariation", &MomaRD::cJdot)
      .def("computeVelocity", &MomaRD::cVel)
      .def("computeMinDistance", &MomaRD::cMinDist, py::arg("q_virtual"), py::arg("q_mobile"), py::arg("q_mani"), py::arg("qdot_virtual"), py::arg("qdot_mobile"),
           py::arg("qdot_mani"), py::arg("with_grad"), py::arg("with_graddot"), py::arg("verbose") = false)
      .def("computeSelectionMatrix", &MomaRD::computeSelectionMatrix).def("computeJacobianActuated", &MomaRD::cJacAct)
      .def("computeJacobianTimeVariationActuated", &MomaRD::cJdotAct).def("computeManipulability", &MomaRD::cMani)
      .def("computeMobileFKJacobian", &MomaRD::computeMobileFKJacobian).def("computeMobileBaseVel", &MomaRD::computeMobileBaseVel)
      .def("getActuatordDof", &MomaRD::getActuatordDof).def("getManipulatorDof", &MomaRD::getManipulatorDof).def("getMobileDof", &MomaRD::getMobileDof)
      .def("getJointIndex", &MomaRD::getJointIndex).def("getActuatorIndex", &MomaRD::getActuatorIndex)
      .def("getMobileJointPosition", &MomaRD::getMobileJointPosition).def("getVirtualJointPosition", &MomaRD::getVirtualJointPosition)
      .def("getManiJointPosition", &MomaRD::getManiJointPosition).def("getJointVelocityActuated", &MomaRD::getJointVelocityActuated)
      .def("getMobileJointVelocity", &MomaRD::getMobileJointVelocity).def("getVirtualJointVelocity", &MomaRD::getVirtualJointVelocity)
      .def("getManiJointVelocity", &MomaRD::getManiJointVelocity).def("getJointPositionActuated", &MomaRD::getJointPositionActuated)
      .def("getMassMatrixActuated", &MomaRD::getMassMatrixActuated).def("getMassMatrixActuatedInv", &MomaRD::getMassMatrixActuatedInv)
      .def("getGravityActuated", &MomaRD::getGravityActuated).def("getCoriolisActuated", &MomaRD::getCoriolisActuated)
      .def("getNonlinearEffectsActuated", &MomaRD::getNonlinearEffectsActuated)
      .def("getJacobianActuated", &MomaRD::getJacobianActuated).def("getJacobianActuatedTimeVariation", &MomaRD::getJacobianActuatedTimeVariation)
      .def("getSelectionMatrix", &MomaRD::getSelectionMatrix).def("getManipulability", &MomaRD::getManipulability)
      .def("getMobileFKJacobian", &MomaRD::getMobileFKJacobian).def("getMobileBaseVel", &MomaRD::getMobileBaseVel);

  py::class_<MobileRC>(m, "MobileRobotController")
      .def(py::init<double, std::shared_ptr<MobileRD>>())
      .def("computeWheelVel", &MobileRC::computeWheelVel).def("computeIKJacobian", &MobileRC::computeIKJacobian).def("VelocityCommand", &MobileRC::VelocityCommand);

  py::class_<ManipRC>(m, "ManipulatorRobotController")
      .def(py::init<double, std::shared_ptr<ManipRD>>())
      .def("setJointGain", &ManipRC::setJointGain).def("setJointKpGain", &ManipRC::setJointKpGain).def("setJointKvGain", &ManipRC::setJointKvGain)
      .def("setTaskGain", &ManipRC::setTaskGain).def("setTaskKpGain", &ManipRC::setTaskKpGain).def("setTaskKvGain", &ManipRC::setTaskKvGain)
      .def("moveJointPositionCubic", &ManipRC::moveJointPositionCubic).def("moveJointVelocityCubic", &ManipRC::moveJointVelocityCubic)
      .def("moveJointTorqueStep", &ManipRC::torqueFromAcc).def("moveJointTorqueStep", &ManipRC::torqueStep)
      .def("moveJointTorqueCubic", &ManipRC::moveJointTorqueCubic)
      .def("CLIKStep", &ManipRC::CLIKStep1).def("CLIKStep", &ManipRC::CLIKStep2).def("CLIKCubic", &ManipRC::CLIKCubic1).def("CLIKCubic", &ManipRC::CLIKCubic2)
      .def("OSF", &ManipRC::OSF1).def("OSF", &ManipRC::OSF2).def("OSFStep", &ManipRC::OSFStep1).def("OSFStep", &ManipRC::OSFStep2)
      .def("OSFCubic", &ManipRC::OSFCubic1).def("OSFCubic", &ManipRC::OSFCubic2)
      .def("QPIK", &ManipRC::QPIK).def("QPIKStep", &ManipRC::QPIKStep).def("QPIKCubic", &ManipRC::QPIKCubic)
      .def("QPID", &ManipRC::QPID).def("QPIDStep", &ManipRC::QPIDStep).def("QPIDCubic", &ManipRC::QPIDCubic);

  py::class_<MomaRC>(m, "MobileManipulatorRobotController")
      .def(py::init<double, std::shared_ptr<MomaRD>>())
      .def("setManipulatorJointGain", &MomaRC::setManipulatorJointGain).def("setManipulatorJointKpGain", &MomaRC::setManipulatorJointKpGain)
      .def("setManipulatorJointKvGain", &MomaRC::setManipulatorJointKvGain)
      .def("setTaskGain", &MomaRC::setTaskGain).def("setTaskKpGain", &MomaRC::setTaskKpGain).def("setTaskKvGain", &MomaRC::setTaskKvGain)
      .def("moveManipulatorJointPositionCubic", &MomaRC::moveManipulatorJointPositionCubic)
      .def("moveManipulatorJointTorqueStep", &MomaRC::torqueFromAcc).def("moveManipulatorJointTorqueStep", &MomaRC::torqueStep)
      .def("moveManipulatorJointTorqueCubic", &MomaRC::moveManipulatorJointTorqueCubic)
      .def("QPIK", &MomaRC::QPIK).def("QPIKStep", &MomaRC::QPIKStep).def("QPIKCubic", &MomaRC::QPIKCubic)
      .def("QPID", &MomaRC::QPID).def("QPIDStep", &MomaRC::QPIDStep).def("QPIDCubic", &MomaRC::QPIDCubic);
}
