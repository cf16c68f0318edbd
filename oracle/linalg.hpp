// oracle/linalg.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Tiny fixed-size fp64 linear algebra used by the CPU restatement of the
// reference's Nano-GICP / Quatro arithmetic.  The reference gets these from
// Eigen (not installed here):
//   * Eigen::JacobiSVD<Matrix3d>          third_party/nano_gicp/include/nano_gicp/impl/nano_gicp_impl.hpp:332
//   * Eigen::Matrix4d::inverse            ...nano_gicp_impl.hpp:208   (block-diagonal => 3x3 inverse)
//   * Eigen::LDLT<Matrix<double,6,6>>     ...impl/lsq_registration_impl.hpp:147,172
//   * Quaterniond::toRotationMatrix       ...lsq_registration_impl.hpp:176 via so3_exp (gicp/so3.hpp:99-118)
// Written independently of the CUDA product code on purpose: the product uses a
// symmetric Jacobi eigen-solver and an un-pivoted Cholesky, the oracle a
// one-sided (Hestenes) Jacobi SVD and a pivoted LDLT, so a bug in either does
// not cancel in the parity tests.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>

namespace orc {

struct M3 {
  double m[9];  // row-major
  double& operator()(int r, int c) { return m[3 * r + c]; }
  double operator()(int r, int c) const { return m[3 * r + c]; }
};

inline M3 m3_zero() {
  M3 a;
  for (int i = 0; i < 9; i++) a.m[i] = 0;
  return a;
}
inline M3 m3_identity() {
  M3 a = m3_zero();
  a(0, 0) = a(1, 1) = a(2, 2) = 1;
  return a;
}
inline M3 m3_mul(const M3& a, const M3& b) {
  M3 c = m3_zero();
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += a(i, k) * b(k, j);
      c(i, j) = s;
    }
  return c;
}
inline M3 m3_transpose(const M3& a) {
  M3 c;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c(i, j) = a(j, i);
  return c;
}
inline M3 m3_add(const M3& a, const M3& b) {
  M3 c;
  for (int i = 0; i < 9; i++) c.m[i] = a.m[i] + b.m[i];
  return c;
}
inline void m3_vec(const M3& a, const double v[3], double out[3]) {
  for (int i = 0; i < 3; i++) out[i] = a(i, 0) * v[0] + a(i, 1) * v[1] + a(i, 2) * v[2];
}
inline double m3_det(const M3& a) {
  return a(0, 0) * (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) - a(0, 1) * (a(1, 0) * a(2, 2) - a(1, 2) * a(2, 0)) +
         a(0, 2) * (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0));
}
// general 3x3 inverse by cofactors (what Eigen does for fixed sizes <= 4)
inline M3 m3_inverse(const M3& a) {
  M3 c;
  double det = m3_det(a);
  double id = 1.0 / det;
  c(0, 0) = (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) * id;
  c(0, 1) = (a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2)) * id;
  c(0, 2) = (a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1)) * id;
  c(1, 0) = (a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2)) * id;
  c(1, 1) = (a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0)) * id;
  c(1, 2) = (a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2)) * id;
  c(2, 0) = (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0)) * id;
  c(2, 1) = (a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1)) * id;
  c(2, 2) = (a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0)) * id;
  return c;
}

// One-sided Jacobi SVD A = U diag(s) V^T, singular values sorted descending
// (the ordering Eigen::JacobiSVD guarantees).
inline void m3_svd(const M3& A, M3& U, double s[3], M3& V) {
  double B[3][3], Vm[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      B[i][j] = A(i, j);
      Vm[i][j] = (i == j);
    }
  for (int sweep = 0; sweep < 60; sweep++) {
    bool rotated = false;
    for (int i = 0; i < 2; i++)
      for (int j = i + 1; j < 3; j++) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int k = 0; k < 3; k++) {
          alpha += B[k][i] * B[k][i];
          beta += B[k][j] * B[k][j];
          gamma += B[k][i] * B[k][j];
        }
        if (gamma == 0.0 || std::fabs(gamma) <= 1e-17 * std::sqrt(alpha * beta)) continue;
        rotated = true;
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double c = 1.0 / std::sqrt(1.0 + t * t), sn = c * t;
        for (int k = 0; k < 3; k++) {
          double bi = B[k][i], bj = B[k][j];
          B[k][i] = c * bi - sn * bj;
          B[k][j] = sn * bi + c * bj;
          double vi = Vm[k][i], vj = Vm[k][j];
          Vm[k][i] = c * vi - sn * vj;
          Vm[k][j] = sn * vi + c * vj;
        }
      }
    if (!rotated) break;
  }
  double sv[3];
  for (int j = 0; j < 3; j++) sv[j] = std::sqrt(B[0][j] * B[0][j] + B[1][j] * B[1][j] + B[2][j] * B[2][j]);
  int ord[3] = {0, 1, 2};
  std::sort(ord, ord + 3, [&](int a, int b) { return sv[a] > sv[b]; });
  double Um[3][3];
  for (int jj = 0; jj < 3; jj++) {
    int j = ord[jj];
    s[jj] = sv[j];
    for (int k = 0; k < 3; k++) {
      V(k, jj) = Vm[k][j];
      Um[k][jj] = sv[j] > 0 ? B[k][j] / sv[j] : 0.0;
    }
  }
  // complete the left basis where a singular value is (numerically) zero
  const double tiny = 1e-14 * (s[0] > 0 ? s[0] : 1.0);
  if (s[0] <= tiny) {  // zero matrix: U = V
    for (int k = 0; k < 3; k++)
      for (int j = 0; j < 3; j++) Um[k][j] = V(k, j);
  } else {
    if (s[1] <= tiny) {  // rank 1: take v1 projected off u0
      double d = 0;
      for (int k = 0; k < 3; k++) d += Um[k][0] * V(k, 1);
      double nn = 0;
      for (int k = 0; k < 3; k++) {
        Um[k][1] = V(k, 1) - d * Um[k][0];
        nn += Um[k][1] * Um[k][1];
      }
      nn = std::sqrt(nn);
      for (int k = 0; k < 3; k++) Um[k][1] /= nn;
    }
    if (s[2] <= tiny) {  // u2 = +-(u0 x u1), sign matching v2 (A symmetric PSD => u2 == v2)
      double c0 = Um[1][0] * Um[2][1] - Um[2][0] * Um[1][1];
      double c1 = Um[2][0] * Um[0][1] - Um[0][0] * Um[2][1];
      double c2 = Um[0][0] * Um[1][1] - Um[1][0] * Um[0][1];
      double d = c0 * V(0, 2) + c1 * V(1, 2) + c2 * V(2, 2);
      double sg = d >= 0 ? 1.0 : -1.0;
      Um[0][2] = sg * c0;
      Um[1][2] = sg * c1;
      Um[2][2] = sg * c2;
    }
  }
  for (int k = 0; k < 3; k++)
    for (int j = 0; j < 3; j++) U(k, j) = Um[k][j];
}

// Symmetric 6x6 solve (H x = rhs) by LDL^T with diagonal pivoting (Eigen::LDLT's strategy).
inline void ldlt6_solve(const double Hin[36], const double rhs[6], double x[6]) {
  double A[6][6];
  int perm[6];
  for (int i = 0; i < 6; i++) {
    perm[i] = i;
    for (int j = 0; j < 6; j++) A[i][j] = Hin[6 * i + j];
  }
  double L[6][6] = {{0}}, D[6];
  for (int k = 0; k < 6; k++) {
    int p = k;
    double best = std::fabs(A[k][k]);
    for (int i = k + 1; i < 6; i++)
      if (std::fabs(A[i][i]) > best) {
        best = std::fabs(A[i][i]);
        p = i;
      }
    if (p != k) {
      for (int j = 0; j < 6; j++) std::swap(A[k][j], A[p][j]);
      for (int i = 0; i < 6; i++) std::swap(A[i][k], A[i][p]);
      for (int j = 0; j < k; j++) std::swap(L[k][j], L[p][j]);
      std::swap(perm[k], perm[p]);
    }
    D[k] = A[k][k];
    L[k][k] = 1.0;
    for (int i = k + 1; i < 6; i++) L[i][k] = D[k] != 0.0 ? A[i][k] / D[k] : 0.0;
    for (int i = k + 1; i < 6; i++)
      for (int j = k + 1; j < 6; j++) A[i][j] -= L[i][k] * D[k] * L[j][k];
  }
  double y[6], z[6];
  for (int i = 0; i < 6; i++) {
    double s = rhs[perm[i]];
    for (int j = 0; j < i; j++) s -= L[i][j] * y[j];
    y[i] = s;
  }
  for (int i = 0; i < 6; i++) z[i] = D[i] != 0.0 ? y[i] / D[i] : 0.0;
  for (int i = 5; i >= 0; i--) {
    double s = z[i];
    for (int j = i + 1; j < 6; j++) s -= L[j][i] * y[j];
    y[i] = s;
  }
  for (int i = 0; i < 6; i++) x[perm[i]] = y[i];
}

// rigid transform: R row-major + t
struct Iso {
  M3 R;
  double t[3];
};
inline Iso iso_identity() {
  Iso x;
  x.R = m3_identity();
  x.t[0] = x.t[1] = x.t[2] = 0;
  return x;
}
inline Iso iso_mul(const Iso& a, const Iso& b) {  // a * b
  Iso c;
  c.R = m3_mul(a.R, b.R);
  m3_vec(a.R, b.t, c.t);
  for (int i = 0; i < 3; i++) c.t[i] += a.t[i];
  return c;
}
inline void iso_to_rowmajor16(const Iso& x, double T[16]) {
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T[4 * i + j] = x.R(i, j);
    T[4 * i + 3] = x.t[i];
  }
  T[12] = T[13] = T[14] = 0;
  T[15] = 1;
}
inline Iso iso_from_rowmajor16(const double T[16]) {
  Iso x;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) x.R(i, j) = T[4 * i + j];
    x.t[i] = T[4 * i + 3];
  }
  return x;
}

// so3_exp (third_party/nano_gicp/include/nano_gicp/gicp/so3.hpp:99-118) followed by
// Eigen's Quaternion::toRotationMatrix.
inline M3 so3_exp_matrix(const double w[3]) {
  double theta_sq = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double imag, real;
  if (theta_sq < 1e-10) {
    double theta_quad = theta_sq * theta_sq;
    imag = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    double theta = std::sqrt(theta_sq);
    double half = 0.5 * theta;
    imag = std::sin(half) / theta;
    real = std::cos(half);
  }
  double qw = real, qx = imag * w[0], qy = imag * w[1], qz = imag * w[2];
  double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  double twx = tx * qw, twy = ty * qw, twz = tz * qw;
  double txx = tx * qx, txy = ty * qx, txz = tz * qx;
  double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  M3 R;
  R(0, 0) = 1 - (tyy + tzz);
  R(0, 1) = txy - twz;
  R(0, 2) = txz + twy;
  R(1, 0) = txy + twz;
  R(1, 1) = 1 - (txx + tzz);
  R(1, 2) = tyz - twx;
  R(2, 0) = txz - twy;
  R(2, 1) = tyz + twx;
  R(2, 2) = 1 - (txx + tyy);
  return R;
}

}  // namespace orc
