// drc_b200 -- mobile base entry points of the C ABI (include/drc_b200.h, "mobile base").
// Replaces, for batches of base states, Mobile::RobotData (reference src/mobile/robot_data.cpp:7-204) and
// Mobile::RobotController (src/mobile/robot_controller.cpp:7-124).  The base has no URDF: the handle holds the
// KinematicParam only.  One thread per base; the parameter block travels as a __grid_constant__ kernel argument.
// There is NO CPU fallback.
#include "drc_host.h"
#include "drc_mobile.h"

struct drc_mobile {
  MobileDev dev;
  int device;
  cudaStream_t stream;
  double* stage; size_t stage_doubles;   // staging of the drc_host_mobile_* entry points (grown on demand)
  long long launches;
};

namespace {

// One thread per base.  WMAX / EXACT: see drc_mobile.h -- bases with 2, 4 or 8 wheels get predicate-free bodies.
template <int WMAX, bool EXACT>
__global__ void __launch_bounds__(128) k_mobile_fk(const __grid_constant__ MobileDev m, const MobileIO io) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < io.B) mobile_fk_job_t<WMAX, EXACT>(m, io, b);
}
template <int WMAX, bool EXACT>
__global__ void __launch_bounds__(128) k_mobile_ik(const __grid_constant__ MobileDev m, const MobileIO io) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < io.B) mobile_ik_job_t<WMAX, EXACT>(m, io, b);
}
#define DRC_DISPATCH_WHEELS(w, KERNEL, ...)                                             \
  switch (w) {                                                                          \
    case 2: KERNEL<2, true> __VA_ARGS__; break;                                         \
    case 4: KERNEL<4, true> __VA_ARGS__; break;                                         \
    case 8: KERNEL<8, true> __VA_ARGS__; break;                                         \
    default: KERNEL<kMaxWheel, false> __VA_ARGS__; break;                               \
  }

int check_mobile(const drc_mobile* h, int B) {
  if (!h) return fail(DRC_E_INVALID, "null mobile base handle");
  if (B <= 0) return fail(DRC_E_INVALID, "batch size must be positive");
  return DRC_OK;
}

int launch_fk(drc_mobile* h, int B, const double* wheel_pos, const double* wheel_vel, double* J, double* base_vel, int layout,
              cudaStream_t s) {
  const int w = h->dev.wheel_num;
  if (h->dev.drive_type == kCaster && !wheel_pos) return fail(DRC_E_INVALID, "caster bases need wheel_pos (steering angles)");
  if (base_vel && !wheel_vel) return fail(DRC_E_INVALID, "base_vel requested without wheel_vel");
  MobileIO io;
  std::memset(&io, 0, sizeof io);
  io.B = B; io.wheel_pos = wheel_pos; io.swp = lay(layout, w, B); io.wheel_vel = wheel_vel; io.swv = io.swp;
  io.J = J; io.sj = lay(layout, 3 * w, B); io.out = base_vel; io.so = lay(layout, 3, B);
  DRC_DISPATCH_WHEELS(w, k_mobile_fk, <<<(B + 127) / 128, 128, 0, s>>>(h->dev, io))
  CU(cudaGetLastError());
  h->launches += 1;
  return DRC_OK;
}
int launch_ik(drc_mobile* h, int B, const double* wheel_pos, const double* base_vel_des, int saturate, double* J, double* wheel_vel,
              int layout, cudaStream_t s) {
  const int w = h->dev.wheel_num;
  if (h->dev.drive_type == kCaster && !wheel_pos) return fail(DRC_E_INVALID, "caster bases need wheel_pos (steering angles)");
  if (wheel_vel && !base_vel_des) return fail(DRC_E_INVALID, "wheel velocities requested without a base velocity");
  MobileIO io;
  std::memset(&io, 0, sizeof io);
  io.B = B; io.wheel_pos = wheel_pos; io.swp = lay(layout, w, B); io.base_vel = base_vel_des; io.sbv = lay(layout, 3, B);
  io.J = J; io.sj = lay(layout, 3 * w, B); io.out = wheel_vel; io.so = lay(layout, w, B); io.saturate = saturate;
  DRC_DISPATCH_WHEELS(w, k_mobile_ik, <<<(B + 127) / 128, 128, 0, s>>>(h->dev, io))
  CU(cudaGetLastError());
  h->launches += 1;
  return DRC_OK;
}

// host staging: one device buffer, inputs first, outputs behind them
struct MobileStage {
  drc_mobile* h;
  size_t used = 0;
  struct Out { double* host; double* dev; size_t cnt; };
  std::vector<Out> outs;
  int err = DRC_OK;
  int reserve(size_t doubles) {
    if (doubles <= h->stage_doubles) return DRC_OK;
    if (h->stage) { cudaStreamSynchronize(h->stream); cudaFree(h->stage); h->stage = nullptr; h->stage_doubles = 0; }
    if (cudaMalloc(&h->stage, doubles * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return fail(DRC_E_NOMEM, "mobile base: staging allocation failed"); }
    h->stage_doubles = doubles;
    return DRC_OK;
  }
  double* in(const double* p, size_t cnt) {
    if (!p) return nullptr;
    double* d = h->stage + used; used += cnt;
    if (cudaMemcpyAsync(d, p, cnt * sizeof(double), cudaMemcpyHostToDevice, h->stream) != cudaSuccess) err = DRC_E_CUDA;
    return d;
  }
  double* out(double* p, size_t cnt) {
    if (!p) return nullptr;
    double* d = h->stage + used; used += cnt;
    outs.push_back({p, d, cnt});
    return d;
  }
  int finish(int rc) {
    if (err) return fail(err, "mobile base: H2D copy failed");
    if (rc) return rc;
    for (auto& o : outs)
      if (cudaMemcpyAsync(o.host, o.dev, o.cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) return fail(DRC_E_CUDA, "mobile base: D2H copy failed");
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return fail(DRC_E_CUDA, std::string("mobile base kernel failed: ") + cudaGetErrorString(e));
    return DRC_OK;
  }
};

}  // namespace

extern "C" {

// Mobile::RobotData::RobotData(KinematicParam) (mobile/robot_data.cpp:7-32): wheel_num = 2 (differential),
// roller_angles.size() (mecanum), 2 * base2wheel_positions.size() (caster)
int drc_mobile_create(int drive_type, double wheel_radius, double base_width, double wheel_offset, double max_lin_speed,
                      double max_ang_speed, double max_lin_acc, double max_ang_acc, int n_wheels, const double* roller_angles,
                      const double* b2w_x, const double* b2w_y, const double* b2w_angles, int device, drc_mobile_t** out) {
  if (!out) return fail(DRC_E_INVALID, "null output handle");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return fail(DRC_E_CUDA, "no CUDA device: drc_b200 has no CPU fallback");
  }
  if (device < 0 || device >= ndev) return fail(DRC_E_INVALID, "device index out of range");
  if (!(wheel_radius > 0)) return fail(DRC_E_INVALID, "mobile base: wheel_radius must be positive");
  MobileParam p;
  p.drive_type = drive_type; p.wheel_radius = wheel_radius; p.base_width = base_width; p.wheel_offset = wheel_offset;
  int w = 0;
  if (drive_type == kDifferential) {
    w = 2;
    if (!(base_width > 0)) return fail(DRC_E_INVALID, "mobile base: differential drives need base_width > 0");
  } else if (drive_type == kMecanum) {
    w = n_wheels;
    if (w < 3 || w > kMaxWheel || !roller_angles || !b2w_x || !b2w_y || !b2w_angles)
      return fail(DRC_E_INVALID, "mobile base: mecanum drives need 3..8 wheels with roller_angles, base2wheel_positions and base2wheel_angles");
    for (int i = 0; i < w; ++i) { p.roller_angles.push_back(roller_angles[i]); p.b2w_x.push_back(b2w_x[i]); p.b2w_y.push_back(b2w_y[i]); p.b2w_angles.push_back(b2w_angles[i]); }
  } else if (drive_type == kCaster) {
    w = n_wheels;
    if (w < 4 || w > kMaxWheel || (w & 1) || !b2w_x || !b2w_y) return fail(DRC_E_INVALID, "mobile base: caster drives need 2..4 casters (4..8 joints: steer, roll) with base2wheel_positions");
    if (!(wheel_offset > 0)) return fail(DRC_E_INVALID, "mobile base: caster drives need wheel_offset > 0");
    for (int i = 0; i < w / 2; ++i) { p.b2w_x.push_back(b2w_x[i]); p.b2w_y.push_back(b2w_y[i]); }
  } else {
    return fail(DRC_E_INVALID, "mobile base: unknown drive type");
  }
  std::unique_ptr<drc_mobile> h(new drc_mobile());
  std::memset(&h->dev, 0, sizeof h->dev);
  MobileDev& d = h->dev;
  d.drive_type = drive_type; d.wheel_num = w; d.wheel_radius = wheel_radius; d.base_width = base_width; d.wheel_offset = wheel_offset;
  d.max_lin_speed = max_lin_speed; d.max_ang_speed = max_ang_speed; d.max_lin_acc = max_lin_acc; d.max_ang_acc = max_ang_acc;
  for (size_t i = 0; i < p.b2w_x.size(); ++i) { d.b2w_x[i] = p.b2w_x[i]; d.b2w_y[i] = p.b2w_y[i]; }
  try {
    mobile_constant_jacobians(p, w, d.J_fk, d.J_ik);
  } catch (const std::exception& e) {
    return fail(DRC_E_INVALID, e.what());
  }
  h->device = device; h->stage = nullptr; h->stage_doubles = 0; h->launches = 0;
  CU(cudaSetDevice(device));
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  *out = h.release();
  return DRC_OK;
}
void drc_mobile_destroy(drc_mobile_t* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->stage) cudaFree(h->stage);
  cudaStreamDestroy(h->stream);
  delete h;
}
int drc_mobile_wheel_num(const drc_mobile_t* h) { return h ? h->dev.wheel_num : fail(DRC_E_INVALID, "null mobile base handle"); }
long long drc_mobile_launch_count(const drc_mobile_t* h) { return h ? h->launches : 0; }
int drc_mobile_synchronize(drc_mobile_t* h) {
  if (!h) return fail(DRC_E_INVALID, "null mobile base handle");
  CU(cudaStreamSynchronize(h->stream));
  return DRC_OK;
}

int drc_batch_mobile_fk(drc_mobile_t* h, int B, const double* wheel_pos, const double* wheel_vel, double* J_fk, double* base_vel,
                        int layout, void* stream) {
  int rc = check_mobile(h, B); if (rc) return rc;
  CU(cudaSetDevice(h->device));
  return launch_fk(h, B, wheel_pos, wheel_vel, J_fk, base_vel, layout, stream ? (cudaStream_t)stream : h->stream);
}
int drc_batch_mobile_ik(drc_mobile_t* h, int B, const double* wheel_pos, const double* base_vel_des, int saturate, double* J_ik,
                        double* wheel_vel, int layout, void* stream) {
  int rc = check_mobile(h, B); if (rc) return rc;
  CU(cudaSetDevice(h->device));
  return launch_ik(h, B, wheel_pos, base_vel_des, saturate, J_ik, wheel_vel, layout, stream ? (cudaStream_t)stream : h->stream);
}
int drc_host_mobile_fk(drc_mobile_t* h, int B, const double* wheel_pos, const double* wheel_vel, double* J_fk, double* base_vel) {
  int rc = check_mobile(h, B); if (rc) return rc;
  CU(cudaSetDevice(h->device));
  const size_t Bz = (size_t)B, w = (size_t)h->dev.wheel_num;
  MobileStage st{h};
  rc = st.reserve(Bz * (2 * w + 3 * w + 3)); if (rc) return rc;
  const double* d_wp = st.in(wheel_pos, Bz * w);
  const double* d_wv = st.in(wheel_vel, Bz * w);
  double* d_J = st.out(J_fk, Bz * 3 * w);
  double* d_bv = st.out(base_vel, Bz * 3);
  return st.finish(launch_fk(h, B, d_wp, d_wv, d_J, d_bv, DRC_LAYOUT_AOS, h->stream));
}
int drc_host_mobile_ik(drc_mobile_t* h, int B, const double* wheel_pos, const double* base_vel_des, int saturate, double* J_ik,
                       double* wheel_vel) {
  int rc = check_mobile(h, B); if (rc) return rc;
  CU(cudaSetDevice(h->device));
  const size_t Bz = (size_t)B, w = (size_t)h->dev.wheel_num;
  MobileStage st{h};
  rc = st.reserve(Bz * (w + 3 + 3 * w + w)); if (rc) return rc;
  const double* d_wp = st.in(wheel_pos, Bz * w);
  const double* d_bv = st.in(base_vel_des, Bz * 3);
  double* d_J = st.out(J_ik, Bz * 3 * w);
  double* d_wv = st.out(wheel_vel, Bz * w);
  return st.finish(launch_ik(h, B, d_wp, d_bv, saturate, d_J, d_wv, DRC_LAYOUT_AOS, h->stream));
}

}  // extern "C"
