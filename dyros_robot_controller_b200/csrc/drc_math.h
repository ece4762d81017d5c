// drc_b200 -- small fixed-size fp64 helpers for per-thread (one robot per thread) kernel code.
// Plain arrays + fully unrolled loops so everything stays in registers.
#pragma once
#include "drc_common.h"

namespace drc {

struct Vec3 {
  double x, y, z;
};
DRC_HD Vec3 v3(double x, double y, double z) { return Vec3{x, y, z}; }
DRC_HD Vec3 operator+(Vec3 a, Vec3 b) { return Vec3{a.x + b.x, a.y + b.y, a.z + b.z}; }
DRC_HD Vec3 operator-(Vec3 a, Vec3 b) { return Vec3{a.x - b.x, a.y - b.y, a.z - b.z}; }
DRC_HD Vec3 operator-(Vec3 a) { return Vec3{-a.x, -a.y, -a.z}; }
DRC_HD Vec3 operator*(double s, Vec3 a) { return Vec3{s * a.x, s * a.y, s * a.z}; }
DRC_HD double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DRC_HD Vec3 cross(Vec3 a, Vec3 b) { return Vec3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
DRC_HD double norm2(Vec3 a) { return dot(a, a); }
DRC_HD double norm(Vec3 a) { return sqrt(dot(a, a)); }
DRC_HD Vec3 axpy(double s, Vec3 a, Vec3 b) { return Vec3{s * a.x + b.x, s * a.y + b.y, s * a.z + b.z}; }
DRC_HD double comp(Vec3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// Row-major 3x3.
struct Mat3 {
  double m[9];
};
DRC_HD Vec3 mul(const Mat3& A, Vec3 v) {
  return Vec3{A.m[0] * v.x + A.m[1] * v.y + A.m[2] * v.z, A.m[3] * v.x + A.m[4] * v.y + A.m[5] * v.z,
              A.m[6] * v.x + A.m[7] * v.y + A.m[8] * v.z};
}
DRC_HD Vec3 tmul(const Mat3& A, Vec3 v) {
  return Vec3{A.m[0] * v.x + A.m[3] * v.y + A.m[6] * v.z, A.m[1] * v.x + A.m[4] * v.y + A.m[7] * v.z,
              A.m[2] * v.x + A.m[5] * v.y + A.m[8] * v.z};
}
DRC_HD Mat3 mul(const Mat3& A, const Mat3& B) {
  Mat3 C;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return C;
}
DRC_HD Mat3 tmul(const Mat3& A, const Mat3& B) {  // A^T B
  Mat3 C;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C.m[3 * i + j] = A.m[i] * B.m[j] + A.m[3 + i] * B.m[3 + j] + A.m[6 + i] * B.m[6 + j];
  return C;
}
DRC_HD Mat3 mat3_from(const double* a) {
  Mat3 R;
#pragma unroll
  for (int i = 0; i < 9; ++i) R.m[i] = a[i];
  return R;
}
DRC_HD Vec3 col(const Mat3& A, int j) { return Vec3{A.m[j], A.m[3 + j], A.m[6 + j]}; }
DRC_HD Mat3 identity3() { Mat3 R = {{1, 0, 0, 0, 1, 0, 0, 0, 1}}; return R; }

// R(axis, angle) for a unit axis (Rodrigues).
DRC_HD Mat3 rot_axis(Vec3 a, double s, double c) {
  const double t = 1.0 - c;
  Mat3 R;
  R.m[0] = c + t * a.x * a.x;       R.m[1] = t * a.x * a.y - s * a.z; R.m[2] = t * a.x * a.z + s * a.y;
  R.m[3] = t * a.x * a.y + s * a.z; R.m[4] = c + t * a.y * a.y;       R.m[5] = t * a.y * a.z - s * a.x;
  R.m[6] = t * a.x * a.z - s * a.y; R.m[7] = t * a.y * a.z + s * a.x; R.m[8] = c + t * a.z * a.z;
  return R;
}

DRC_HD double dmin(double a, double b) { return a < b ? a : b; }
DRC_HD double dmax(double a, double b) { return a > b ? a : b; }
DRC_HD double clampd(double x, double lo, double hi) { return dmin(dmax(x, lo), hi); }

// Symmetric packed index (upper triangle, row-major): (i<=j) -> i*N - i(i-1)/2 + (j-i)
template <int N>
DRC_HD constexpr int symidx(int i, int j) {
  return i <= j ? i * N - (i * (i - 1)) / 2 + (j - i) : j * N - (j * (j - 1)) / 2 + (i - j);
}

// In-place Cholesky of a dense row-major SPD matrix (lower factor in the lower triangle).
// Returns the smallest pivot d_jj (before the square root); <= 0 means not positive definite.
template <int N>
DRC_HD double chol_inplace(double* A) {
  double minpiv = 1e300;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double d = A[j * N + j];
#pragma unroll
    for (int k = 0; k < j; ++k) d -= A[j * N + k] * A[j * N + k];
    minpiv = dmin(minpiv, d);
    const double dj = d > 0 ? sqrt(d) : 1.0;
    const double inv = 1.0 / dj;
    A[j * N + j] = dj;
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      double s = A[i * N + j];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= A[i * N + k] * A[j * N + k];
      A[i * N + j] = s * inv;
    }
  }
  return minpiv;
}
template <int N>
DRC_HD void chol_solve(const double* L, double* b) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= L[i * N + k] * b[k];
    b[i] = s / L[i * N + i];
  }
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    double s = b[i];
#pragma unroll
    for (int k = i + 1; k < N; ++k) s -= L[k * N + i] * b[k];
    b[i] = s / L[i * N + i];
  }
}
// Ainv = (L L^T)^-1, dense symmetric output.
template <int N>
DRC_HD void chol_inverse(const double* L, double* Ainv) {
#pragma unroll
  for (int c = 0; c < N; ++c) {
    double e[N];
#pragma unroll
    for (int i = 0; i < N; ++i) e[i] = (i == c) ? 1.0 : 0.0;
    chol_solve<N>(L, e);
#pragma unroll
    for (int i = 0; i < N; ++i) Ainv[i * N + c] = e[i];
  }
}

// Pseudo-inverse by column-pivoted Householder QR with Eigen's rank rule
// (|R_ii| > threshold * max|R_ii|) and minimum-norm completion -- what
// CompleteOrthogonalDecomposition::pseudoInverse() returns (reference math_type_define.h:563-570).
// Runtime-indexed (local memory); only taken on the rare ill-conditioned path.
// Two phases: (1) the pivoted factorisation of R with the reflectors kept aside, which fixes the rank; (2) Q from the reflectors (the
// same updates in the same order as accumulating it on the fly) and the pseudo-inverse.  `stop_if_full_rank`: return true after
// phase 1 without touching `out` when no pivot is truncated -- the caller then knows that PinvCOD(A) is the plain inverse and takes
// it by a cheaper route (spd_pinv: whole-body mass matrices sit a factor 1.5 above the rank threshold, every robot came here).
template <int M, int N>
DRC_HD_NOINLINE bool pinv_cpqr(const double* Ain, double* out /* N x M */, double threshold, double* abs_det = nullptr,
                               bool stop_if_full_rank = false) {
  double R[M * N], Q[M * M], v[M];
  constexpr int K = M < N ? M : N;
  double V[K * M], vn2s[K];   // reflector k: V[k][k..M), its squared norm (0 = none)
  int perm[N];
  for (int i = 0; i < M * N; ++i) R[i] = Ain[i];
  for (int j = 0; j < N; ++j) perm[j] = j;
  for (int k = 0; k < K; ++k) {
    vn2s[k] = 0.0;
    int piv = k;
    double best = -1.0;
    for (int j = k; j < N; ++j) {
      double s = 0;
      for (int i = k; i < M; ++i) s += R[i * N + j] * R[i * N + j];
      if (s > best) { best = s; piv = j; }
    }
    if (piv != k) {
      for (int i = 0; i < M; ++i) { double t = R[i * N + k]; R[i * N + k] = R[i * N + piv]; R[i * N + piv] = t; }
      int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
    }
    const double nrm = sqrt(best);
    if (nrm == 0) continue;
    const double alpha = R[k * N + k] > 0 ? -nrm : nrm;
    for (int i = k; i < M; ++i) v[i] = R[i * N + k];
    v[k] -= alpha;
    double vn2 = 0;
    for (int i = k; i < M; ++i) vn2 += v[i] * v[i];
    if (vn2 == 0) continue;
    for (int j = k; j < N; ++j) {
      double s = 0;
      for (int i = k; i < M; ++i) s += v[i] * R[i * N + j];
      s = 2 * s / vn2;
      for (int i = k; i < M; ++i) R[i * N + j] -= s * v[i];
    }
    for (int i = k; i < M; ++i) V[k * M + i] = v[i];
    vn2s[k] = vn2;
  }
  double maxpiv = 0;
  for (int k = 0; k < K; ++k) maxpiv = dmax(maxpiv, fabs(R[k * N + k]));
  if (abs_det) {  // |det| of a square input = prod |R_kk|
    double pd = 1.0;
    for (int k = 0; k < K; ++k) pd *= fabs(R[k * N + k]);
    *abs_det = pd;
  }
  int r = 0;
  for (int k = 0; k < K; ++k) if (fabs(R[k * N + k]) > threshold * maxpiv) ++r;
  if (stop_if_full_rank && r == K) return true;
  for (int i = 0; i < N * M; ++i) out[i] = 0.0;
  if (r == 0) return false;
  // phase 2: Q = H_0 H_1 ... applied to the identity
  for (int i = 0; i < M * M; ++i) Q[i] = 0.0;
  for (int i = 0; i < M; ++i) Q[i * M + i] = 1.0;
  for (int k = 0; k < K; ++k) {
    const double vn2 = vn2s[k];
    if (vn2 == 0) continue;
    for (int j = 0; j < M; ++j) {
      double s = 0;
      for (int i = k; i < M; ++i) s += Q[j * M + i] * V[k * M + i];
      s = 2 * s / vn2;
      for (int i = k; i < M; ++i) Q[j * M + i] -= s * V[k * M + i];
    }
  }
  double G[K * K], e[K], Wp[N * K];
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < r; ++j) {
      double s = 0;
      for (int l = 0; l < N; ++l) s += R[i * N + l] * R[j * N + l];
      G[i * r + j] = s;
    }
  // Cholesky of the r x r Gram matrix (runtime size)
  for (int j = 0; j < r; ++j) {
    double d = G[j * r + j];
    for (int k = 0; k < j; ++k) d -= G[j * r + k] * G[j * r + k];
    d = sqrt(d > 0 ? d : 1e-300);
    G[j * r + j] = d;
    for (int i = j + 1; i < r; ++i) {
      double s = G[i * r + j];
      for (int k = 0; k < j; ++k) s -= G[i * r + k] * G[j * r + k];
      G[i * r + j] = s / d;
    }
  }
  for (int c = 0; c < r; ++c) {
    for (int i = 0; i < r; ++i) e[i] = (i == c) ? 1.0 : 0.0;
    for (int i = 0; i < r; ++i) {
      double s = e[i];
      for (int k = 0; k < i; ++k) s -= G[i * r + k] * e[k];
      e[i] = s / G[i * r + i];
    }
    for (int i = r - 1; i >= 0; --i) {
      double s = e[i];
      for (int k = i + 1; k < r; ++k) s -= G[k * r + i] * e[k];
      e[i] = s / G[i * r + i];
    }
    for (int l = 0; l < N; ++l) {
      double s = 0;
      for (int i = 0; i < r; ++i) s += R[i * N + l] * e[i];
      Wp[l * K + c] = s;
    }
  }
  for (int l = 0; l < N; ++l)
    for (int j = 0; j < M; ++j) {
      double s = 0;
      for (int c = 0; c < r; ++c) s += Wp[l * K + c] * Q[j * M + c];
      out[perm[l] * M + j] = s;
    }
  return false;
}

}  // namespace drc
