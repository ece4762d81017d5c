// ORACLE -- test infrastructure only (see oracle/README.md). PARITY UNPINNED (no reference goldens exist).
// Small fixed-size linear algebra used by the CPU restatement.  No Eigen in this image,
// so the Eigen calls of the reference (products, determinant, CompleteOrthogonalDecomposition
// pseudo-inverse with threshold 1e-6: reference include/math_type_define.h:563-570) are
// restated here on plain row-major double arrays.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace orc {

struct V3 {
  double x, y, z;
  V3() : x(0), y(0), z(0) {}
  V3(double a, double b, double c) : x(a), y(b), z(c) {}
  double& operator[](int i) { return (&x)[i]; }
  double operator[](int i) const { return (&x)[i]; }
};
inline V3 operator+(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(const V3& a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(double s, const V3& a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(const V3& a, double s) { return {s * a.x, s * a.y, s * a.z}; }
inline V3& operator+=(V3& a, const V3& b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
inline V3& operator-=(V3& a, const V3& b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
inline double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(const V3& a, const V3& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double norm(const V3& a) { return std::sqrt(dot(a, a)); }

struct M3 {
  double m[9];  // row-major
  M3() { std::memset(m, 0, sizeof m); }
  static M3 identity() { M3 r; r.m[0] = r.m[4] = r.m[8] = 1; return r; }
  double& operator()(int i, int j) { return m[3 * i + j]; }
  double operator()(int i, int j) const { return m[3 * i + j]; }
  V3 col(int j) const { return {m[j], m[3 + j], m[6 + j]}; }
  V3 row(int i) const { return {m[3 * i], m[3 * i + 1], m[3 * i + 2]}; }
};
inline V3 operator*(const M3& A, const V3& v) {
  return {A.m[0] * v.x + A.m[1] * v.y + A.m[2] * v.z, A.m[3] * v.x + A.m[4] * v.y + A.m[5] * v.z,
          A.m[6] * v.x + A.m[7] * v.y + A.m[8] * v.z};
}
inline V3 tmul(const M3& A, const V3& v) {  // A^T v
  return {A.m[0] * v.x + A.m[3] * v.y + A.m[6] * v.z, A.m[1] * v.x + A.m[4] * v.y + A.m[7] * v.z,
          A.m[2] * v.x + A.m[5] * v.y + A.m[8] * v.z};
}
inline M3 operator*(const M3& A, const M3& B) {
  M3 C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += A(i, k) * B(k, j);
      C(i, j) = s;
    }
  return C;
}
inline M3 operator+(const M3& A, const M3& B) { M3 C; for (int i = 0; i < 9; ++i) C.m[i] = A.m[i] + B.m[i]; return C; }
inline M3 operator-(const M3& A, const M3& B) { M3 C; for (int i = 0; i < 9; ++i) C.m[i] = A.m[i] - B.m[i]; return C; }
inline M3 operator*(double s, const M3& A) { M3 C; for (int i = 0; i < 9; ++i) C.m[i] = s * A.m[i]; return C; }
inline M3 transpose(const M3& A) { M3 C; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C(i, j) = A(j, i); return C; }
inline M3 skew(const V3& v) {
  M3 S;
  S(0, 1) = -v.z; S(0, 2) = v.y; S(1, 0) = v.z; S(1, 2) = -v.x; S(2, 0) = -v.y; S(2, 1) = v.x;
  return S;
}
inline M3 outer(const V3& a, const V3& b) {
  M3 C;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C(i, j) = a[i] * b[j];
  return C;
}
// Rodrigues rotation about a unit axis.
inline M3 axis_angle(const V3& a, double th) {
  double c = std::cos(th), s = std::sin(th), t = 1 - c;
  M3 R;
  R(0, 0) = c + t * a.x * a.x;       R(0, 1) = t * a.x * a.y - s * a.z; R(0, 2) = t * a.x * a.z + s * a.y;
  R(1, 0) = t * a.x * a.y + s * a.z; R(1, 1) = c + t * a.y * a.y;       R(1, 2) = t * a.y * a.z - s * a.x;
  R(2, 0) = t * a.x * a.z - s * a.y; R(2, 1) = t * a.y * a.z + s * a.x; R(2, 2) = c + t * a.z * a.z;
  return R;
}

struct SE3 {
  M3 R;
  V3 p;
  SE3() : R(M3::identity()) {}
  SE3(const M3& r, const V3& t) : R(r), p(t) {}
};
inline SE3 operator*(const SE3& a, const SE3& b) { return {a.R * b.R, a.R * b.p + a.p}; }

// ---------------------------------------------------------------------------------------------
// Dense helpers on row-major arrays (runtime sizes, caller-owned storage).
// ---------------------------------------------------------------------------------------------
using Mat = std::vector<double>;

inline void matmul(const double* A, const double* B, double* C, int m, int k, int n) {  // C(m,n)=A(m,k)B(k,n)
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int l = 0; l < k; ++l) s += A[i * k + l] * B[l * n + j];
      C[i * n + j] = s;
    }
}
inline void matmul_tn(const double* A, const double* B, double* C, int k, int m, int n) {  // C(m,n)=A(k,m)^T B(k,n)
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int l = 0; l < k; ++l) s += A[l * m + i] * B[l * n + j];
      C[i * n + j] = s;
    }
}
inline void matmul_nt(const double* A, const double* B, double* C, int m, int k, int n) {  // C(m,n)=A(m,k)B(n,k)^T
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int l = 0; l < k; ++l) s += A[i * k + l] * B[j * k + l];
      C[i * n + j] = s;
    }
}
inline void matvec(const double* A, const double* x, double* y, int m, int n) {
  for (int i = 0; i < m; ++i) {
    double s = 0;
    for (int j = 0; j < n; ++j) s += A[i * n + j] * x[j];
    y[i] = s;
  }
}
inline void matvec_t(const double* A, const double* x, double* y, int m, int n) {  // y(n) = A(m,n)^T x(m)
  for (int j = 0; j < n; ++j) y[j] = 0;
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) y[j] += A[i * n + j] * x[i];
}

// Determinant by LU with partial pivoting (Eigen's MatrixXd::determinant() uses PartialPivLU for n>4).
inline double determinant(const double* Ain, int n) {
  std::vector<double> A(Ain, Ain + n * n);
  double det = 1;
  for (int k = 0; k < n; ++k) {
    int piv = k;
    double best = std::fabs(A[k * n + k]);
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(A[i * n + k]) > best) { best = std::fabs(A[i * n + k]); piv = i; }
    if (best == 0) return 0;
    if (piv != k) {
      for (int j = 0; j < n; ++j) std::swap(A[k * n + j], A[piv * n + j]);
      det = -det;
    }
    det *= A[k * n + k];
    for (int i = k + 1; i < n; ++i) {
      double f = A[i * n + k] / A[k * n + k];
      for (int j = k + 1; j < n; ++j) A[i * n + j] -= f * A[k * n + j];
    }
  }
  return det;
}

// Cholesky (lower) in place; returns false if not positive definite.
inline bool cholesky(double* A, int n) {
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0)) return false;
    d = std::sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
  }
  return true;
}
inline void chol_solve(const double* L, double* b, int n) {  // solves L L^T x = b in place
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[i * n + k] * b[k];
    b[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * b[k];
    b[i] = s / L[i * n + i];
  }
}

// Pseudo-inverse through a rank-revealing (column-pivoted Householder QR) decomposition with
// Eigen's rank rule: rank = #{ |R_ii| > threshold * max|R_ii| } (threshold 1e-6,
// reference math_type_define.h:7,563-570).  For the retained rank-r factor W = [R11 R12] the
// minimum-norm solution W^+ = W^T (W W^T)^-1 equals what the complete orthogonal decomposition
// returns, so pinv(A) = P W^+ Q1^T.
inline void pinv_cod(const double* Ain, int m, int n, double* out /* n x m */, double threshold = 1e-6,
                     int* rank_out = nullptr) {
  std::vector<double> R(Ain, Ain + m * n);       // becomes R (upper-trapezoidal), m x n
  std::vector<double> Q(m * m, 0.0);             // accumulates Q (m x m)
  for (int i = 0; i < m; ++i) Q[i * m + i] = 1;
  std::vector<int> perm(n);
  for (int j = 0; j < n; ++j) perm[j] = j;
  const int kmax = std::min(m, n);
  std::vector<double> v(m);
  for (int k = 0; k < kmax; ++k) {
    // pivot: column with the largest remaining norm
    int piv = k;
    double best = -1;
    for (int j = k; j < n; ++j) {
      double s = 0;
      for (int i = k; i < m; ++i) s += R[i * n + j] * R[i * n + j];
      if (s > best) { best = s; piv = j; }
    }
    if (piv != k) {
      for (int i = 0; i < m; ++i) std::swap(R[i * n + k], R[i * n + piv]);
      std::swap(perm[k], perm[piv]);
    }
    double nrm = std::sqrt(best);
    if (nrm == 0) continue;
    double alpha = R[k * n + k] > 0 ? -nrm : nrm;
    for (int i = k; i < m; ++i) v[i] = R[i * n + k];
    v[k] -= alpha;
    double vn2 = 0;
    for (int i = k; i < m; ++i) vn2 += v[i] * v[i];
    if (vn2 == 0) continue;
    for (int j = k; j < n; ++j) {
      double s = 0;
      for (int i = k; i < m; ++i) s += v[i] * R[i * n + j];
      s = 2 * s / vn2;
      for (int i = k; i < m; ++i) R[i * n + j] -= s * v[i];
    }
    for (int j = 0; j < m; ++j) {  // Q <- Q H
      double s = 0;
      for (int i = k; i < m; ++i) s += Q[j * m + i] * v[i];
      s = 2 * s / vn2;
      for (int i = k; i < m; ++i) Q[j * m + i] -= s * v[i];
    }
  }
  double maxpiv = 0;
  for (int k = 0; k < kmax; ++k) maxpiv = std::max(maxpiv, std::fabs(R[k * n + k]));
  int r = 0;
  for (int k = 0; k < kmax; ++k)
    if (std::fabs(R[k * n + k]) > threshold * maxpiv) ++r;
  if (rank_out) *rank_out = r;
  std::fill(out, out + n * m, 0.0);
  if (r == 0) return;
  // W = R[0:r, 0:n]; G = W W^T (r x r); Wp = W^T G^-1 (n x r)
  std::vector<double> G(r * r), Wp(n * r);
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < r; ++j) {
      double s = 0;
      for (int l = 0; l < n; ++l) s += R[i * n + l] * R[j * n + l];
      G[i * r + j] = s;
    }
  cholesky(G.data(), r);
  std::vector<double> e(r);
  for (int c = 0; c < r; ++c) {  // column c of G^-1
    std::fill(e.begin(), e.end(), 0.0);
    e[c] = 1;
    chol_solve(G.data(), e.data(), r);
    for (int l = 0; l < n; ++l) {
      double s = 0;
      for (int i = 0; i < r; ++i) s += R[i * n + l] * e[i];
      Wp[l * r + c] = s;
    }
  }
  // out[perm[l], :] = Wp[l,:] * Q[:, 0:r]^T
  for (int l = 0; l < n; ++l)
    for (int j = 0; j < m; ++j) {
      double s = 0;
      for (int c = 0; c < r; ++c) s += Wp[l * r + c] * Q[j * m + c];
      out[perm[l] * m + j] = s;
    }
}

}  // namespace orc
