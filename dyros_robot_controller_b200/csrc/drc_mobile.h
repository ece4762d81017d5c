// drc_b200 -- mobile base kinematics (kernel bodies, host+device).
//
// Replaces, for batches of base states,
//   Mobile::RobotData::updateState / computeBaseVel / computeFKJacobian      src/mobile/robot_data.cpp:103-204
//   Mobile::RobotController::VelocityCommand / computeWheelVel / computeIKJacobian
//                                                                           src/mobile/robot_controller.cpp:14-124
// Differential and mecanum Jacobians are configuration independent: the host computes them once (MobileDev::J_fk,
// J_ik); the powered-caster Jacobians (Holmberg & Khatib) depend on the steering angles and are evaluated per state.
#pragma once
#include "drc_common.h"

namespace drc {

constexpr int kMaxWheel = 8;

// KinematicParam (include/dyros_robot_controller/type_define.h:58-72) as the kernels see it.
struct MobileDev {
  int drive_type, wheel_num;
  double wheel_radius, base_width, wheel_offset;
  double max_lin_speed, max_ang_speed, max_lin_acc, max_ang_acc;
  double b2w_x[kMaxWheel], b2w_y[kMaxWheel];  // wheel positions (mecanum: per wheel; caster: per steering axis)
  double J_fk[3][kMaxWheel];                  // differential / mecanum forward Jacobian (3 x w)
  double J_ik[kMaxWheel][3];                  // differential / mecanum inverse Jacobian (w x 3)
};

// CasterFKJacobian (mobile/robot_data.cpp:179-203):  J = pinv(Jp' Jp) Jp' Jq_inv, 3 x w row-major with row stride ld.
// Jp' Jp is 3 x 3 and symmetric; for two or more casters it is positive definite and PinvCOD is the inverse.  Below the
// COD rank threshold (relative pivot 1e-6) the reference returns a reduced-rank pseudo-inverse: that case (a single
// caster) is rejected on the host when the base is created.
template <int W>
DRC_HD void caster_fk_jacobian(double r, double b, const double* px, const double* py, const double* wheel_pos, int w,
                               double* J, int ld) {
  double g00 = 0, g01 = 0, g02 = 0, g11 = 0, g12 = 0, g22 = 0;
  double T[3 * W];  // Jp' Jq_inv
  const int ns = w / 2;
#pragma unroll
  for (int i = 0; i < W / 2; ++i) {
    if (i < ns) {
      double s, c;
      sincos(wheel_pos[2 * i], &s, &c);
      const double a = -(py[i] + b * s), d = px[i] + b * c;  // rows (1, 0, a) and (0, 1, d) of Jp
      g00 += 1.0; g11 += 1.0; g02 += a; g12 += d; g22 += a * a + d * d;
      const double q00 = b * s, q01 = r * c, q10 = -b * c, q11 = r * s;  // Jq_inv block
      T[0 * W + 2 * i] = q00;               T[0 * W + 2 * i + 1] = q01;
      T[1 * W + 2 * i] = q10;               T[1 * W + 2 * i + 1] = q11;
      T[2 * W + 2 * i] = a * q00 + d * q10; T[2 * W + 2 * i + 1] = a * q01 + d * q11;
    }
  }
  // inverse of the symmetric 3 x 3 Gram matrix (adjugate)
  const double c00 = g11 * g22 - g12 * g12, c01 = g02 * g12 - g01 * g22, c02 = g01 * g12 - g02 * g11;
  const double c11 = g00 * g22 - g02 * g02, c12 = g01 * g02 - g00 * g12, c22 = g00 * g11 - g01 * g01;
  const double det = g00 * c00 + g01 * c01 + g02 * c02;
  const double id = 1.0 / det;
#pragma unroll
  for (int k = 0; k < W; ++k) {
    if (k < w) {
      const double t0 = T[0 * W + k], t1 = T[1 * W + k], t2 = T[2 * W + k];
      J[0 * ld + k] = id * (c00 * t0 + c01 * t1 + c02 * t2);
      J[1 * ld + k] = id * (c01 * t0 + c11 * t1 + c12 * t2);
      J[2 * ld + k] = id * (c02 * t0 + c12 * t1 + c22 * t2);
    }
  }
}

// CasterIKJacobian (mobile/robot_controller.cpp:104-123): w x 3 row-major
template <int W>
DRC_HD void caster_ik_jacobian(double r, double b, const double* px, const double* py, const double* wheel_pos, int w,
                               double* Ji) {
  const int ns = w / 2;
  const double ib = 1.0 / b, ir = 1.0 / r;   // two divisions per base instead of six per caster (fp64 division is ~30 instructions)
#pragma unroll
  for (int i = 0; i < W / 2; ++i) {
    if (i < ns) {
      double s, c;
      sincos(wheel_pos[2 * i], &s, &c);
      Ji[(2 * i) * 3 + 0] = -s * ib; Ji[(2 * i) * 3 + 1] = c * ib; Ji[(2 * i) * 3 + 2] = (px[i] * c + py[i] * s) * ib - 1.0;
      Ji[(2 * i + 1) * 3 + 0] = c * ir; Ji[(2 * i + 1) * 3 + 1] = s * ir; Ji[(2 * i + 1) * 3 + 2] = (px[i] * s - py[i] * c) * ir;
    }
  }
}

struct MobileIO {
  int B;
  const double* wheel_pos; Strided swp;   // w per base (needed for caster bases only; may be null otherwise)
  const double* wheel_vel; Strided swv;   // w per base (FK) -- null = Jacobian only
  const double* base_vel; Strided sbv;    // desired (vx, vy, omega) per base (IK)
  double* J; Strided sj;                  // FK: 3 x w row-major | IK: w x 3 row-major (null = skip)
  double* out; Strided so;                // FK: base velocity (3) | IK: wheel velocities (w)
  int saturate;                           // IK: VelocityCommand's speed saturation before the inverse map
};

// Mobile::RobotData::updateState (mobile/robot_data.cpp:103-114): J_mobile = computeFKJacobian(wheel_pos),
// base_vel = J_mobile * wheel_vel.  WMAX = compile-time bound of the wheel count; EXACT: the base has exactly WMAX wheels
// (the loops then carry no predicates and differential / mecanum bases read their constant Jacobian straight from the
// parameter block -- the kernels are HBM bound only when the per-base instruction count stays in the tens).
template <int WMAX, bool EXACT>
DRC_HD void mobile_fk_job_t(const MobileDev& m, const MobileIO& io, int b) {
  const int w = EXACT ? WMAX : m.wheel_num;
  double v0 = 0, v1 = 0, v2 = 0;
  if (m.drive_type == kCaster) {
    double J[3 * WMAX], wp[WMAX];
#pragma unroll
    for (int k = 0; k < WMAX; ++k) wp[k] = k < w ? io.wheel_pos[b * io.swp.sb + k * io.swp.sk] : 0.0;
    caster_fk_jacobian<WMAX>(m.wheel_radius, m.wheel_offset, m.b2w_x, m.b2w_y, wp, w, J, WMAX);
    if (io.J) {
      for (int r = 0; r < 3; ++r)
        for (int k = 0; k < w; ++k) io.J[b * io.sj.sb + (r * w + k) * io.sj.sk] = J[r * WMAX + k];
    }
    if (io.wheel_vel && io.out) {
#pragma unroll
      for (int k = 0; k < WMAX; ++k) {
        if (k < w) {
          const double wv = io.wheel_vel[b * io.swv.sb + k * io.swv.sk];
          v0 += J[0 * WMAX + k] * wv; v1 += J[1 * WMAX + k] * wv; v2 += J[2 * WMAX + k] * wv;
        }
      }
    }
  } else {
    if (io.J) {
      for (int r = 0; r < 3; ++r)
        for (int k = 0; k < w; ++k) io.J[b * io.sj.sb + (r * w + k) * io.sj.sk] = m.J_fk[r][k];
    }
    if (io.wheel_vel && io.out) {
#pragma unroll
      for (int k = 0; k < WMAX; ++k) {
        if (k < w) {
          const double wv = io.wheel_vel[b * io.swv.sb + k * io.swv.sk];
          v0 += m.J_fk[0][k] * wv; v1 += m.J_fk[1][k] * wv; v2 += m.J_fk[2][k] * wv;
        }
      }
    }
  }
  if (io.wheel_vel && io.out) {
    io.out[b * io.so.sb + 0 * io.so.sk] = v0;
    io.out[b * io.so.sb + 1 * io.so.sk] = v1;
    io.out[b * io.so.sb + 2 * io.so.sk] = v2;
  }
}
DRC_HD void mobile_fk_job(const MobileDev& m, const MobileIO& io, int b) { mobile_fk_job_t<kMaxWheel, false>(m, io, b); }

// VelocityCommand's saturation (mobile/robot_controller.cpp:14-41): the planar speed is clipped to max_lin_speed along its
// own direction (directions of speeds below 1e-4 are dropped), the yaw rate to +-max_ang_speed.
DRC_HD void saturate_base_velocity(const MobileDev& m, double* v) {
  const double sp = sqrt(v[0] * v[0] + v[1] * v[1]);
  const double sc = fmin(fmax(sp, -m.max_lin_speed), m.max_lin_speed);
  const double k = fabs(sp) < 1e-4 ? 0.0 : sc / sp;   // direction (v / speed) times the clipped speed, one division
  v[0] *= k; v[1] *= k;
  v[2] = fmin(fmax(v[2], -m.max_ang_speed), m.max_ang_speed);
}

// Mobile::RobotController::VelocityCommand / computeWheelVel / computeIKJacobian (mobile/robot_controller.cpp:14-124)
template <int WMAX, bool EXACT>
DRC_HD void mobile_ik_job_t(const MobileDev& m, const MobileIO& io, int b) {
  const int w = EXACT ? WMAX : m.wheel_num;
  double v[3] = {0, 0, 0};
  const bool cmd = io.base_vel && io.out;
  if (cmd) {
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = io.base_vel[b * io.sbv.sb + c * io.sbv.sk];
    if (io.saturate) saturate_base_velocity(m, v);
  }
  if (m.drive_type == kCaster) {
    double Ji[WMAX * 3], wp[WMAX];
#pragma unroll
    for (int k = 0; k < WMAX; ++k) wp[k] = k < w ? io.wheel_pos[b * io.swp.sb + k * io.swp.sk] : 0.0;
    caster_ik_jacobian<WMAX>(m.wheel_radius, m.wheel_offset, m.b2w_x, m.b2w_y, wp, w, Ji);
    if (io.J) {
      for (int k = 0; k < w; ++k)
        for (int c = 0; c < 3; ++c) io.J[b * io.sj.sb + (k * 3 + c) * io.sj.sk] = Ji[k * 3 + c];
    }
    if (cmd) {
#pragma unroll
      for (int k = 0; k < WMAX; ++k)
        if (k < w) io.out[b * io.so.sb + k * io.so.sk] = Ji[k * 3 + 0] * v[0] + Ji[k * 3 + 1] * v[1] + Ji[k * 3 + 2] * v[2];
    }
  } else {
    if (io.J) {
      for (int k = 0; k < w; ++k)
        for (int c = 0; c < 3; ++c) io.J[b * io.sj.sb + (k * 3 + c) * io.sj.sk] = m.J_ik[k][c];
    }
    if (cmd) {
#pragma unroll
      for (int k = 0; k < WMAX; ++k)
        if (k < w) io.out[b * io.so.sb + k * io.so.sk] = m.J_ik[k][0] * v[0] + m.J_ik[k][1] * v[1] + m.J_ik[k][2] * v[2];
    }
  }
}
DRC_HD void mobile_ik_job(const MobileDev& m, const MobileIO& io, int b) { mobile_ik_job_t<kMaxWheel, false>(m, io, b); }

}  // namespace drc
