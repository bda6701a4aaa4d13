// ORACLE — TEST INFRASTRUCTURE ONLY.  Parity unpinned (see oracle/README.md).
//
// Tiny fixed-size linear algebra for the CPU restatement of mrg_slam's scan
// registration hot path.  The reference (/root/reference) contains none of this
// arithmetic: it lives in Eigen (un-vendored), used by PCL / ndt_omp / fast_gicp
// which src/mrg_slam/registrations.cpp:46-148 instantiates.  Everything here
// restates the published algorithms without Eigen.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
// arm may use anything under oracle/.
#pragma once
#include <cmath>
#include <cstring>
#include <algorithm>

namespace orc {

// ---- 3x3 (row-major double[9]) ------------------------------------------------
inline void m3_mul(const double* A, const double* B, double* C) {
  double r[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += A[i * 3 + k] * B[k * 3 + j];
      r[i * 3 + j] = s;
    }
  std::memcpy(C, r, sizeof(r));
}
inline void m3_mul_bt(const double* A, const double* B, double* C) {  // C = A * B^T
  double r[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += A[i * 3 + k] * B[j * 3 + k];
      r[i * 3 + j] = s;
    }
  std::memcpy(C, r, sizeof(r));
}
inline void m3_vec(const double* A, const double* v, double* o) {
  double r0 = A[0] * v[0] + A[1] * v[1] + A[2] * v[2];
  double r1 = A[3] * v[0] + A[4] * v[1] + A[5] * v[2];
  double r2 = A[6] * v[0] + A[7] * v[1] + A[8] * v[2];
  o[0] = r0; o[1] = r1; o[2] = r2;
}
inline bool m3_inverse(const double* A, double* inv) {
  double c00 = A[4] * A[8] - A[5] * A[7];
  double c01 = A[5] * A[6] - A[3] * A[8];
  double c02 = A[3] * A[7] - A[4] * A[6];
  double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  double id = 1.0 / det;
  double r[9];
  r[0] = c00 * id;
  r[1] = (A[2] * A[7] - A[1] * A[8]) * id;
  r[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  r[3] = c01 * id;
  r[4] = (A[0] * A[8] - A[2] * A[6]) * id;
  r[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  r[6] = c02 * id;
  r[7] = (A[1] * A[6] - A[0] * A[7]) * id;
  r[8] = (A[0] * A[4] - A[1] * A[3]) * id;
  std::memcpy(inv, r, sizeof(r));
  return det != 0.0 && std::isfinite(id);
}

// Symmetric 3x3 eigen-decomposition by cyclic Jacobi.  evals ascending, evecs
// columns (row-major V[i*3+j] = component i of eigenvector j).  Stands in for
// Eigen::JacobiSVD on a symmetric PSD matrix (fast_gicp calculate_covariances,
// App. A.1) and Eigen::SelfAdjointEigenSolver (ndt_omp VoxelGridCovariance, A.4).
inline void sym3_eigen(const double* Ain, double* evals, double* V) {
  double A[9];
  std::memcpy(A, Ain, sizeof(A));
  double Q[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
    double diag = A[0] * A[0] + A[4] * A[4] + A[8] * A[8];
    if (off <= 1e-34 * diag || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double apq = A[p * 3 + q];
        if (apq == 0.0) continue;
        double app = A[p * 3 + p], aqq = A[q * 3 + q];
        double theta = (aqq - app) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        // A <- J^T A J
        for (int k = 0; k < 3; ++k) {
          double akp = A[k * 3 + p], akq = A[k * 3 + q];
          A[k * 3 + p] = c * akp - s * akq;
          A[k * 3 + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          double apk = A[p * 3 + k], aqk = A[q * 3 + k];
          A[p * 3 + k] = c * apk - s * aqk;
          A[q * 3 + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          double qkp = Q[k * 3 + p], qkq = Q[k * 3 + q];
          Q[k * 3 + p] = c * qkp - s * qkq;
          Q[k * 3 + q] = s * qkp + c * qkq;
        }
      }
  }
  int order[3] = {0, 1, 2};
  double d[3] = {A[0], A[4], A[8]};
  std::sort(order, order + 3, [&](int a, int b) { return d[a] < d[b]; });
  for (int j = 0; j < 3; ++j) {
    evals[j] = d[order[j]];
    for (int i = 0; i < 3; ++i) V[i * 3 + j] = Q[i * 3 + order[j]];
  }
}

// ---- 4x4 general inverse by cofactors (Eigen's fixed-size 4x4 inverse is also
// an analytic cofactor form; differences are O(1e-16) relative, App. A.2 last line).
inline bool m4_inverse(const double* m, double* out) {
  double inv[16];
  inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
  double id = 1.0 / det;
  for (int i = 0; i < 16; ++i) out[i] = inv[i] * id;
  return det != 0.0;
}

// ---- 6x6 ----------------------------------------------------------------------
// Solve (A) x = rhs for symmetric A by LDL^T with symmetric (diagonal) pivoting,
// as Eigen::LDLT does (fast_gicp LsqRegistration::step_lm, App. A.1).
inline void ldlt6_solve(const double* Ain, const double* rhs, double* x) {
  const int n = 6;
  double A[36];
  std::memcpy(A, Ain, sizeof(A));
  int perm[6] = {0, 1, 2, 3, 4, 5};
  for (int k = 0; k < n; ++k) {
    int piv = k;
    double best = std::fabs(A[k * n + k]);
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(A[i * n + i]) > best) { best = std::fabs(A[i * n + i]); piv = i; }
    if (piv != k) {
      for (int j = 0; j < n; ++j) std::swap(A[k * n + j], A[piv * n + j]);
      for (int i = 0; i < n; ++i) std::swap(A[i * n + k], A[i * n + piv]);
      std::swap(perm[k], perm[piv]);
    }
    double d = A[k * n + k];
    if (d == 0.0) continue;
    for (int i = k + 1; i < n; ++i) A[i * n + k] /= d;  // L(i,k)
    for (int i = k + 1; i < n; ++i)
      for (int j = k + 1; j <= i; ++j) {
        A[i * n + j] -= A[i * n + k] * d * A[j * n + k];
        A[j * n + i] = A[i * n + j];
      }
  }
  double y[6];
  for (int i = 0; i < n; ++i) y[i] = rhs[perm[i]];
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < i; ++k) y[i] -= A[i * n + k] * y[k];
  for (int i = 0; i < n; ++i) y[i] = (A[i * n + i] != 0.0) ? y[i] / A[i * n + i] : 0.0;
  for (int i = n - 1; i >= 0; --i)
    for (int k = i + 1; k < n; ++k) y[i] -= A[k * n + i] * y[k];
  for (int i = 0; i < n; ++i) x[perm[i]] = y[i];
}

// x = pinv(A) rhs through a one-sided (Hestenes) Jacobi SVD of a general 6x6;
// stands in for Eigen::JacobiSVD<Matrix6d>(H, FullU|FullV).solve(-g) in
// ndt_omp computeTransformation (App. A.3).  Singular values below
// eps*6*sigma_max are dropped, as Eigen's default threshold does.
inline void svd6_solve(const double* Ain, const double* rhs, double* x) {
  const int n = 6;
  double U[36], V[36];
  std::memcpy(U, Ain, sizeof(U));  // columns get orthogonalised: A V = U S
  for (int i = 0; i < 36; ++i) V[i] = 0.0;
  for (int i = 0; i < n; ++i) V[i * n + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < n; ++i) {
          alpha += U[i * n + p] * U[i * n + p];
          beta += U[i * n + q] * U[i * n + q];
          gamma += U[i * n + p] * U[i * n + q];
        }
        if (gamma == 0.0 || std::fabs(gamma) <= 1e-15 * std::sqrt(alpha * beta)) continue;
        rotated = true;
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < n; ++i) {
          double up = U[i * n + p], uq = U[i * n + q];
          U[i * n + p] = c * up - s * uq;
          U[i * n + q] = s * up + c * uq;
          double vp = V[i * n + p], vq = V[i * n + q];
          V[i * n + p] = c * vp - s * vq;
          V[i * n + q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  double sig[6], smax = 0.0;
  for (int j = 0; j < n; ++j) {
    double s = 0;
    for (int i = 0; i < n; ++i) s += U[i * n + j] * U[i * n + j];
    sig[j] = std::sqrt(s);
    smax = std::max(smax, sig[j]);
  }
  double thr = 2.220446049250313e-16 * n * smax;
  for (int i = 0; i < n; ++i) x[i] = 0.0;
  for (int j = 0; j < n; ++j) {
    if (!(sig[j] > thr)) continue;
    double proj = 0;  // (u_j . rhs) / sigma_j^2 with u_j = U_col / sigma
    for (int i = 0; i < n; ++i) proj += U[i * n + j] * rhs[i];
    proj /= (sig[j] * sig[j]);
    for (int i = 0; i < n; ++i) x[i] += V[i * n + j] * proj;
  }
}

}  // namespace orc
