// ORACLE — TEST INFRASTRUCTURE ONLY.  Parity unpinned (see oracle/README.md).
//
// pcl::GeneralizedIterativeClosestPoint<PointXYZI, PointXYZI> — what registration_method == "GICP" selects
// (/root/reference/src/mrg_slam/registrations.cpp:93-103; "GICP_OMP", :104-116, is pclomp's OpenMP copy of the same
// algorithm) — SURVEY.md §8a row G / Appendix A.5.  Neither PCL nor pclomp is in /root/reference, so this restates the
// public upstream sources from memory of PCL 1.12: pcl/registration/impl/gicp.hpp (computeCovariances,
// computeTransformation, estimateRigidTransformationBFGS, OptimizationFunctorWithIndices, computeRDerivative, applyState)
// and pcl/registration/bfgs.h (BFGS<Functor>, a port of GSL's vector_bfgs2 + linear_minimize.c).  The product has no
// engine for this method yet (select_registration_method raises / returns nullptr for "GICP"); this file is the checker a
// device implementation will be held against.  What tests/test_gicp_pcl.py can pin here without the upstream binaries:
// the analytic gradient against finite differences of the cost, the Euler-angle derivative matrices against finite
// differences of applyState, the line search's Wolfe conditions, and recovery of known transforms.
//
// Arithmetic types follow upstream: float Matrix4f state and float point transforms, float products widened to double
// in the covariance moments, double Mahalanobis matrices, double BFGS.
//
// Details restated from memory that a reader with the upstream sources should check first (none can be checked here):
//   * gradient_tol = 1e-2 (local constant of estimateRigidTransformationBFGS) and the loop
//     `do { inner++; result = minimizeOneStep(x); if (result) break; result = testGradient(tol); } while (Running && inner < max)`;
//   * BFGS parameters set by gicp.hpp: sigma = rho = 0.01, tau1 = 9, tau2 = 0.05, tau3 = 0.5, order = 3; bfgs.h defaults
//     step_size = 1, bracket_iters = sect_iters = 100;
//   * the initial state from the matrix: x3 = atan2(T(2,1), T(2,2)), x4 = asin(-T(2,0)), x5 = atan2(T(1,0), T(0,0)), evaluated
//     here with the float overloads (the arguments are Matrix4f entries);
//   * applyState ADDS the translation (t.col(3) += T) and multiplies the rotation on the left, all in float;
//   * the correspondence test `nn_dists[0] < corr_dist_threshold_^2`, mahalanobis_ reset to identity per align;
//   * covariances: `cov(k,l) += pt[k] * pt[l]` are float products; JacobiSVD orders singular values (= |eigenvalues|), the
//     smallest gets gicp_epsilon_;
//   * use_reciprocal_correspondences (false in config/mrg_slam.yaml:106) is not restated.
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "kdtree.hpp"
#include "linalg.hpp"
#include "oracle.h"

using namespace orc;

namespace {

struct Vec6 {
  double v[6];
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
};
inline double dot6(const Vec6& a, const Vec6& b) {
  double s = 0;
  for (int i = 0; i < 6; ++i) s += a[i] * b[i];
  return s;
}
inline double norm6(const Vec6& a) { return std::sqrt(dot6(a, a)); }

// ---- float 4x4, column-major (Eigen::Matrix4f)
struct M4f {
  float m[16];
  float& operator()(int r, int c) { return m[c * 4 + r]; }
  float operator()(int r, int c) const { return m[c * 4 + r]; }
};
inline M4f m4f_identity() {
  M4f I;
  for (int i = 0; i < 16; ++i) I.m[i] = (i % 5 == 0) ? 1.f : 0.f;
  return I;
}
inline M4f m4f_mul(const M4f& A, const M4f& B) {
  M4f C;
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      float s = 0.f;
      for (int k = 0; k < 4; ++k) s += A(r, k) * B(k, c);
      C(r, c) = s;
    }
  return C;
}
// Matrix4f * Vector4f(x, y, z, 1): Eigen's fixed-size product, ((c0 x + c1 y) + c2 z) + c3
inline void m4f_point(const M4f& T, const float* p, float* q) {
  for (int r = 0; r < 3; ++r) {
    float s = T(r, 0) * p[0];
    s = s + T(r, 1) * p[1];
    s = s + T(r, 2) * p[2];
    q[r] = s + T(r, 3);
  }
}

// GICP::applyState: t <- [Rz(x5) Ry(x4) Rx(x3) | (x0, x1, x2)] applied on the left of t, all in float
void apply_state(M4f& t, const Vec6& x) {
  const float a = (float)x[3], b = (float)x[4], c = (float)x[5];
  const float ca = std::cos(a), sa = std::sin(a), cb = std::cos(b), sb = std::sin(b), cc = std::cos(c), sc = std::sin(c);
  const float Rx[9] = {1, 0, 0, 0, ca, -sa, 0, sa, ca};
  const float Ry[9] = {cb, 0, sb, 0, 1, 0, -sb, 0, cb};
  const float Rz[9] = {cc, -sc, 0, sc, cc, 0, 0, 0, 1};
  float ZY[9], R[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float s = 0.f;
      for (int k = 0; k < 3; ++k) s += Rz[i * 3 + k] * Ry[k * 3 + j];
      ZY[i * 3 + j] = s;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float s = 0.f;
      for (int k = 0; k < 3; ++k) s += ZY[i * 3 + k] * Rx[k * 3 + j];
      R[i * 3 + j] = s;
    }
  float top[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float s = 0.f;
      for (int k = 0; k < 3; ++k) s += R[i * 3 + k] * t(k, j);
      top[i * 3 + j] = s;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t(i, j) = top[i * 3 + j];
  t(0, 3) += (float)x[0];
  t(1, 3) += (float)x[1];
  t(2, 3) += (float)x[2];
}

// GICP::computeRDerivative: g[3..5] = trace(dR/dangle * dCost_dR_T), R = Rz(psi) Ry(theta) Rx(phi)
void compute_r_derivative(const Vec6& x, const double* D /*dCost_dR_T, row-major 3x3*/, Vec6& g) {
  const double phi = x[3], theta = x[4], psi = x[5];
  const double cphi = std::cos(phi), sphi = std::sin(phi), ctheta = std::cos(theta), stheta = std::sin(theta), cpsi = std::cos(psi),
               spsi = std::sin(psi);
  double dPhi[9], dTheta[9], dPsi[9];
  dPhi[0] = 0.; dPhi[3] = 0.; dPhi[6] = 0.;
  dPhi[1] = sphi * spsi + cphi * cpsi * stheta;
  dPhi[4] = -cpsi * sphi + cphi * spsi * stheta;
  dPhi[7] = cphi * ctheta;
  dPhi[2] = cphi * spsi - cpsi * sphi * stheta;
  dPhi[5] = -cphi * cpsi - sphi * spsi * stheta;
  dPhi[8] = -ctheta * sphi;

  dTheta[0] = -cpsi * stheta; dTheta[3] = -spsi * stheta; dTheta[6] = -ctheta;
  dTheta[1] = cpsi * ctheta * sphi; dTheta[4] = ctheta * sphi * spsi; dTheta[7] = -sphi * stheta;
  dTheta[2] = cphi * cpsi * ctheta; dTheta[5] = cphi * ctheta * spsi; dTheta[8] = -cphi * stheta;

  dPsi[0] = -ctheta * spsi; dPsi[3] = cpsi * ctheta; dPsi[6] = 0.;
  dPsi[1] = -cphi * cpsi - sphi * spsi * stheta; dPsi[4] = -cphi * spsi + cpsi * sphi * stheta; dPsi[7] = 0.;
  dPsi[2] = cpsi * sphi - cphi * spsi * stheta; dPsi[5] = sphi * spsi + cphi * cpsi * stheta; dPsi[8] = 0.;
  auto inner = [&](const double* A) {  // matricesInnerProd: trace(A * D) = sum_ij A(j,i) D(i,j)
    double r = 0.;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) r += A[j * 3 + i] * D[i * 3 + j];
    return r;
  };
  g[3] = inner(dPhi);
  g[4] = inner(dTheta);
  g[5] = inner(dPsi);
}

// ---- the optimisation problem of one outer iteration (OptimizationFunctorWithIndices)
struct Problem {
  const float* src = nullptr;  // the guess-transformed source cloud (`output`), 4 floats per point
  const float* tgt = nullptr;
  const std::vector<int>* idx_src = nullptr;
  const std::vector<int>* idx_tgt = nullptr;
  const std::vector<double>* mahalanobis = nullptr;  // 9 per source point, row-major
  int evals = 0;

  // fdf: cost f = (1/m) sum d^T M d, d = T(x) p_src - p_tgt (float differences widened), gradient g
  void fdf(const Vec6& x, double* f, Vec6* g) {
    ++evals;
    M4f T = m4f_identity();  // base_transformation_
    apply_state(T, x);
    const int m = (int)idx_src->size();
    double fs = 0.;
    double gt[3] = {0, 0, 0};
    double D[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // dCost_dR_T
    for (int i = 0; i < m; ++i) {
      const float* ps = src + 4 * (size_t)(*idx_src)[i];
      const float* pt = tgt + 4 * (size_t)(*idx_tgt)[i];
      float q[3];
      m4f_point(T, ps, q);
      const double d[3] = {(double)(q[0] - pt[0]), (double)(q[1] - pt[1]), (double)(q[2] - pt[2])};
      const double* M = &(*mahalanobis)[(size_t)(*idx_src)[i] * 9];
      double Md[3];
      m3_vec(M, d, Md);
      fs += d[0] * Md[0] + d[1] * Md[1] + d[2] * Md[2];
      if (g) {
        for (int a = 0; a < 3; ++a) gt[a] += Md[a];
        for (int a = 0; a < 3; ++a)  // p_base_src * Md^T with base_transformation_ = identity
          for (int b = 0; b < 3; ++b) D[a * 3 + b] += (double)ps[a] * Md[b];
      }
    }
    if (f) *f = fs / m;
    if (g) {
      for (int a = 0; a < 3; ++a) (*g)[a] = gt[a] * (2.0 / m);
      for (int a = 0; a < 9; ++a) D[a] *= 2.0 / m;
      compute_r_derivative(x, D, *g);
    }
  }
};

// ---- pcl/registration/bfgs.h (GSL vector_bfgs2 + linear_minimize.c)
enum { BFGS_SUCCESS = 0, BFGS_RUNNING = 1, BFGS_NOPROGRESS = 2, BFGS_NEG_EPS = -3 };

int solve_quadratic(double a, double b, double c, double* x0, double* x1) {  // gsl_poly_solve_quadratic
  if (a == 0) {
    if (b == 0) return 0;
    *x0 = -c / b;
    return 1;
  }
  const double disc = b * b - 4 * a * c;
  if (disc > 0) {
    if (b == 0) {
      const double r = std::sqrt(-c / a);
      *x0 = -r;
      *x1 = r;
    } else {
      const double sgnb = b > 0 ? 1 : -1;
      const double temp = -0.5 * (b + sgnb * std::sqrt(disc));
      const double r1 = temp / a, r2 = c / temp;
      if (r1 < r2) { *x0 = r1; *x1 = r2; } else { *x0 = r2; *x1 = r1; }
    }
    return 2;
  }
  if (disc == 0) {
    *x0 = -0.5 * b / a;
    *x1 = -0.5 * b / a;
    return 2;
  }
  return 0;
}
double interp_quad(double f0, double fp0, double f1, double zl, double zh) {
  const double fl = f0 + zl * (fp0 + zl * (f1 - f0 - fp0));
  const double fh = f0 + zh * (fp0 + zh * (f1 - f0 - fp0));
  const double c = 2 * (f1 - f0 - fp0);  // curvature
  double zmin = zl, fmin = fl;
  if (fh < fmin) { zmin = zh; fmin = fh; }
  if (c > 0) {  // positive curvature required for a minimum
    const double z = -fp0 / c;
    if (z > zl && z < zh) {
      const double f = f0 + z * (fp0 + z * (f1 - f0 - fp0));
      if (f < fmin) { zmin = z; fmin = f; }
    }
  }
  return zmin;
}
inline double cubic(double c0, double c1, double c2, double c3, double z) { return c0 + z * (c1 + z * (c2 + z * c3)); }
inline void check_extremum(double c0, double c1, double c2, double c3, double z, double* zmin, double* fmin) {
  const double y = cubic(c0, c1, c2, c3, z);
  if (y < *fmin) { *zmin = z; *fmin = y; }
}
double interp_cubic(double f0, double fp0, double f1, double fp1, double zl, double zh) {
  const double eta = 3 * (f1 - f0) - 2 * fp0 - fp1;
  const double xi = fp0 + fp1 - 2 * (f1 - f0);
  const double c0 = f0, c1 = fp0, c2 = eta, c3 = xi;
  double zmin = zl, fmin = cubic(c0, c1, c2, c3, zl);
  check_extremum(c0, c1, c2, c3, zh, &zmin, &fmin);
  double z0 = 0, z1 = 0;
  const int n = solve_quadratic(3 * c3, 2 * c2, c1, &z0, &z1);
  if (n == 2) {
    if (z0 > zl && z0 < zh) check_extremum(c0, c1, c2, c3, z0, &zmin, &fmin);
    if (z1 > zl && z1 < zh) check_extremum(c0, c1, c2, c3, z1, &zmin, &fmin);
  } else if (n == 1) {
    if (z0 > zl && z0 < zh) check_extremum(c0, c1, c2, c3, z0, &zmin, &fmin);
  }
  return zmin;
}
double interpolate(double a, double fa, double fpa, double b, double fb, double fpb, double xmin, double xmax, int order) {
  double zmin = (xmin - a) / (b - a), zmax = (xmax - a) / (b - a);
  if (zmin > zmax) std::swap(zmin, zmax);
  double z;
  if (order > 2 && !std::isnan(fpb)) z = interp_cubic(fa, fpa * (b - a), fb, fpb * (b - a), zmin, zmax);
  else z = interp_quad(fa, fpa * (b - a), fb, zmin, zmax);
  return a + z * (b - a);
}

struct Bfgs {
  Problem* fn;
  // parameters (gicp.hpp sets sigma, rho, tau1..3, order; the rest are bfgs.h's defaults)
  double rho = 0.01, sigma = 0.01, tau1 = 9, tau2 = 0.05, tau3 = 0.5, step_size = 1.0;
  int order = 3, bracket_iters = 100, sect_iters = 100;
  // state
  int iter = 0;
  double f = 0, delta_f = 0, fp0 = 0, pnorm = 0, g0norm = 0;
  Vec6 x0, g0, p, gradient, dx;
  // the values of the function / slope along the current direction at the line search's trial points
  double f_along(double alpha) {
    Vec6 xa;
    for (int i = 0; i < 6; ++i) xa[i] = x0[i] + alpha * p[i];
    double fa;
    fn->fdf(xa, &fa, nullptr);
    return fa;
  }
  double df_along(double alpha, Vec6* g_out = nullptr, double* f_out = nullptr) {
    Vec6 xa, ga;
    for (int i = 0; i < 6; ++i) xa[i] = x0[i] + alpha * p[i];
    double fa;
    fn->fdf(xa, &fa, &ga);
    if (g_out) *g_out = ga;
    if (f_out) *f_out = fa;
    return dot6(ga, p);
  }

  int minimize_init(const Vec6& x) {
    iter = 0;
    delta_f = 0;
    for (int i = 0; i < 6; ++i) dx[i] = 0;
    fn->fdf(x, &f, &gradient);
    x0 = x;
    g0 = gradient;
    g0norm = norm6(g0);
    for (int i = 0; i < 6; ++i) p[i] = gradient[i] * -1 / g0norm;
    pnorm = norm6(p);
    fp0 = -g0norm;
    return BFGS_SUCCESS;
  }

  // linear_minimize.c: minimize() — Fletcher's line search (bracketing, then sectioning)
  int line_search(double alpha1, double* alpha_new) {
    const double f0 = f, fp0_ = fp0;  // f(0) and slope at 0 along p
    double falpha, falpha_prev = f0, fpalpha = 0, fpalpha_prev = fp0_;
    double alpha = alpha1, alpha_prev = 0, alpha_next;
    double a = 0, b = alpha, fa = f0, fb = 0, fpa = fp0_, fpb = 0;
    const double kNaN = std::numeric_limits<double>::quiet_NaN();
    int i = 0;
    bool bracketed = false;
    // Begin bracketing
    while (i++ < bracket_iters) {
      falpha = f_along(alpha);
      // Fletcher's rho test
      if (falpha > f0 + alpha * rho * fp0_ || falpha >= falpha_prev) {
        a = alpha_prev; fa = falpha_prev; fpa = fpalpha_prev;
        b = alpha; fb = falpha; fpb = kNaN;
        bracketed = true;
        break;  // goto sectioning
      }
      fpalpha = df_along(alpha);
      // Fletcher's sigma test
      if (std::fabs(fpalpha) <= -sigma * fp0_) {
        *alpha_new = alpha;
        return BFGS_SUCCESS;
      }
      if (fpalpha >= 0) {
        a = alpha; fa = falpha; fpa = fpalpha;
        b = alpha_prev; fb = falpha_prev; fpb = fpalpha_prev;
        bracketed = true;
        break;  // goto sectioning
      }
      const double delta = alpha - alpha_prev;
      {
        const double lower = alpha + delta, upper = alpha + tau1 * delta;
        alpha_next = interpolate(alpha_prev, falpha_prev, fpalpha_prev, alpha, falpha, fpalpha, lower, upper, order);
      }
      alpha_prev = alpha; falpha_prev = falpha; fpalpha_prev = fpalpha;
      alpha = alpha_next;
    }
    (void)bracketed;  // as in GSL, running out of bracketing iterations falls through to sectioning of [a, b]
    // Sectioning of bracket [a, b]
    while (i++ < sect_iters) {
      const double delta = b - a;
      {
        const double lower = a + tau2 * delta, upper = b - tau3 * delta;
        alpha = interpolate(a, fa, fpa, b, fb, fpb, lower, upper, order);
      }
      falpha = f_along(alpha);
      if ((a - alpha) * fpa <= std::numeric_limits<double>::epsilon()) return BFGS_NOPROGRESS;  // roundoff prevents progress
      if (falpha > f0 + rho * alpha * fp0_ || falpha >= fa) {
        b = alpha; fb = falpha; fpb = kNaN;  // a_next = a
      } else {
        fpalpha = df_along(alpha);
        if (std::fabs(fpalpha) <= -sigma * fp0_) {
          *alpha_new = alpha;
          return BFGS_SUCCESS;  // terminate
        }
        if (((b - a) >= 0 && fpalpha >= 0) || ((b - a) <= 0 && fpalpha <= 0)) {
          b = a; fb = fa; fpb = fpa;
          a = alpha; fa = falpha; fpa = fpalpha;
        } else {
          a = alpha; fa = falpha; fpa = fpalpha;
        }
      }
    }
    return BFGS_SUCCESS;
  }

  int minimize_one_step(Vec6& x) {
    double alpha = 0.0, alpha1;
    const double f0 = f;
    if (pnorm == 0.0 || g0norm == 0.0 || fp0 == 0) {
      for (int i = 0; i < 6; ++i) dx[i] = 0;
      return BFGS_NOPROGRESS;
    }
    if (delta_f < 0) {
      const double del = std::max(-delta_f, 10 * std::numeric_limits<double>::epsilon() * std::fabs(f0));
      alpha1 = std::min(1.0, 2.0 * del / (-fp0));
    } else {
      alpha1 = std::fabs(step_size);
    }
    const int status = line_search(alpha1, &alpha);
    if (status != BFGS_SUCCESS) return status;
    // updatePosition(alpha, x, f, gradient)
    Vec6 gnew;
    double fnew;
    df_along(alpha, &gnew, &fnew);
    for (int i = 0; i < 6; ++i) x[i] = x0[i] + alpha * p[i];
    f = fnew;
    gradient = gnew;
    delta_f = f - f0;
    // Choose a new direction for the next step: the memoryless BFGS update p' = g1 - A dx - B dg
    {
      Vec6 dx0, dg0;
      for (int i = 0; i < 6; ++i) { dx0[i] = x[i] - x0[i]; dg0[i] = gradient[i] - g0[i]; }
      dx = dx0;
      const double dxg = dot6(dx0, gradient), dgg = dot6(dg0, gradient), dxdg = dot6(dx0, dg0), dgnorm = norm6(dg0);
      double A, B;
      if (dxdg != 0) {
        B = dxg / dxdg;
        A = -(1.0 + dgnorm * dgnorm / dxdg) * B + dgg / dxdg;
      } else {
        B = 0;
        A = 0;
      }
      for (int i = 0; i < 6; ++i) p[i] = -A * dx0[i] + gradient[i] + -B * dg0[i];
    }
    g0 = gradient;
    x0 = x;
    g0norm = norm6(g0);
    pnorm = norm6(p);
    // update direction and fp0
    const double dir = (dot6(p, gradient) > 0) ? -1.0 : 1.0;
    for (int i = 0; i < 6; ++i) p[i] *= dir / pnorm;
    pnorm = norm6(p);
    fp0 = dot6(p, g0);
    ++iter;
    return BFGS_SUCCESS;
  }

  int test_gradient(double epsilon) const {
    if (epsilon < 0) return BFGS_NEG_EPS;
    return norm6(gradient) < epsilon ? BFGS_SUCCESS : BFGS_RUNNING;
  }
};

// GICP::computeCovariances: kNN moments with FLOAT products widened to double, E[xx^T] - mean mean^T, SVD, singular values
// replaced by (1, 1, gicp_epsilon)
void pcl_covariances(const float* pts, int n, const KdTree& tree, int k, double eps, std::vector<double>& covs9) {
  covs9.assign((size_t)n * 9, 0.0);
#pragma omp parallel
  {
    std::vector<int> idx(k);
    std::vector<float> d2(k);
#pragma omp for schedule(dynamic, 64)
    for (int i = 0; i < n; ++i) {
      const int found = tree.knn(pts + 4 * (size_t)i, k, idx.data(), d2.data());
      double mean[3] = {0, 0, 0}, cov[9] = {0};
      for (int j = 0; j < found; ++j) {
        const float* q = pts + 4 * (size_t)idx[j];
        mean[0] += q[0]; mean[1] += q[1]; mean[2] += q[2];
        cov[0] += q[0] * q[0];
        cov[3] += q[1] * q[0]; cov[4] += q[1] * q[1];
        cov[6] += q[2] * q[0]; cov[7] += q[2] * q[1]; cov[8] += q[2] * q[2];
      }
      for (int a = 0; a < 3; ++a) mean[a] /= (double)k;
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b <= a; ++b) {
          cov[a * 3 + b] /= (double)k;
          cov[a * 3 + b] -= mean[a] * mean[b];
          cov[b * 3 + a] = cov[a * 3 + b];
        }
      double ev[3], V[9];
      sym3_eigen(cov, ev, V);
      // JacobiSVD orders by singular value = |eigenvalue|: the float products can leave the matrix slightly indefinite, and the
      // regularised direction is the one of the smallest MAGNITUDE
      int small = 0;
      if (std::fabs(ev[1]) < std::fabs(ev[small])) small = 1;
      if (std::fabs(ev[2]) < std::fabs(ev[small])) small = 2;
      double* out = &covs9[(size_t)i * 9];
      for (int c = 0; c < 3; ++c) {
        const double v = c == small ? eps : 1.0;
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) out[a * 3 + b] += v * V[a * 3 + c] * V[b * 3 + c];
      }
    }
  }
}

struct GicpPcl {
  orc_gicp_pcl_params prm;
  const float* target;
  int nt;
  const float* source;
  int ns;
  KdTree tree, tree_src;
  std::vector<double> cov_t, cov_s, mahalanobis;
  int inner_total = 0, evals_total = 0;

  // estimateRigidTransformationBFGS.  Returns false where upstream throws ("BFGS solver failed to converge" / too few
  // correspondences).
  bool estimate(const float* output, const std::vector<int>& is, const std::vector<int>& it, M4f& transformation) {
    if (is.size() < 4) return false;
    Vec6 x;
    x[0] = transformation(0, 3); x[1] = transformation(1, 3); x[2] = transformation(2, 3);
    x[3] = std::atan2(transformation(2, 1), transformation(2, 2));
    x[4] = std::asin(-transformation(2, 0));
    x[5] = std::atan2(transformation(1, 0), transformation(0, 0));
    Problem pb;
    pb.src = output; pb.tgt = target; pb.idx_src = &is; pb.idx_tgt = &it; pb.mahalanobis = &mahalanobis;
    Bfgs bfgs;
    bfgs.fn = &pb;
    const double gradient_tol = 1e-2;
    int inner = 0;
    int result = bfgs.minimize_init(x);
    result = BFGS_RUNNING;
    do {
      inner++;
      result = bfgs.minimize_one_step(x);
      if (result) break;
      result = bfgs.test_gradient(gradient_tol);
    } while (result == BFGS_RUNNING && inner < prm.max_optimizer_iterations);
    inner_total += inner;
    evals_total += pb.evals;
    if (result == BFGS_NOPROGRESS || result == BFGS_SUCCESS || inner == prm.max_optimizer_iterations) {
      transformation = m4f_identity();
      apply_state(transformation, x);
      return true;
    }
    return false;
  }

  M4f guess;
  std::vector<float> output;   // guess * input (pcl::transformPointCloud, float)
  std::vector<int> is, it;     // correspondences of the current outer iteration

  int prepare(const float* guess_colmajor) {
    std::memcpy(guess.m, guess_colmajor, sizeof(guess.m));
    tree.build(target, nt);
    tree_src.build(source, ns);
    const int k = prm.correspondence_randomness;
    if (k > nt || k > ns) return -1;
    pcl_covariances(target, nt, tree, k, prm.gicp_epsilon, cov_t);
    pcl_covariances(source, ns, tree_src, k, prm.gicp_epsilon, cov_s);
    mahalanobis.assign((size_t)ns * 9, 0.0);
    for (int i = 0; i < ns; ++i) mahalanobis[(size_t)i * 9] = mahalanobis[(size_t)i * 9 + 4] = mahalanobis[(size_t)i * 9 + 8] = 1.0;
    output.resize((size_t)ns * 4);
    for (int i = 0; i < ns; ++i) {
      m4f_point(guess, source + 4 * (size_t)i, &output[4 * (size_t)i]);
      output[4 * (size_t)i + 3] = source[4 * (size_t)i + 3];
    }
    return 0;
  }

  // the correspondence pass of one outer iteration (gicp.hpp: the for loop over the source points)
  int correspond(const M4f& transformation) {
    is.clear(); it.clear();
    is.reserve(ns); it.reserve(ns);
    const double dist_threshold = prm.max_correspondence_distance * prm.max_correspondence_distance;
    // transform_R = transformation_ * guess in double
    double R[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int kk = 0; kk < 4; ++kk) s += (double)transformation(i, kk) * (double)guess(kk, j);
        R[i * 3 + j] = s;
      }
    for (int i = 0; i < ns; ++i) {
      float q[4] = {0, 0, 0, 0};
      m4f_point(transformation, &output[4 * (size_t)i], q);
      int idx;
      float d2;
      if (tree.knn(q, 1, &idx, &d2) != 1) return -1;
      if ((double)d2 < dist_threshold) {
        double C1[9], C2[9], M[9], tmp[9];
        std::memcpy(C1, &cov_s[(size_t)i * 9], sizeof(C1));
        std::memcpy(C2, &cov_t[(size_t)idx * 9], sizeof(C2));
        m3_mul(R, C1, M);          // M = R*C1
        m3_mul_bt(M, R, tmp);      // temp = M*R'
        for (int a = 0; a < 9; ++a) tmp[a] += C2[a];
        m3_inverse(tmp, &mahalanobis[(size_t)i * 9]);
        is.push_back(i);
        it.push_back(idx);
      }
    }
    return (int)is.size();
  }

  int align(const float* guess_colmajor, orc_result* out) {
    if (prepare(guess_colmajor) != 0) return -1;
    M4f transformation = m4f_identity(), previous = m4f_identity();
    int nr_iterations = 0;
    bool converged = false;
    while (!converged) {
      if (correspond(transformation) < 0) return -1;
      previous = transformation;
      double delta = 0.;
      if (!estimate(output.data(), is, it, transformation)) break;  // upstream: exception caught, loop left, converged_ stays false
      for (int kk = 0; kk < 4; ++kk)
        for (int l = 0; l < 4; ++l) {
          const double ratio = (kk < 3 && l < 3) ? 1. / prm.rotation_epsilon : 1. / prm.transformation_epsilon;
          const double c_delta = ratio * std::fabs((double)previous(kk, l) - (double)transformation(kk, l));
          if (c_delta > delta) delta = c_delta;
        }
      nr_iterations++;
      if (nr_iterations >= prm.maximum_iterations || delta < 1) {
        converged = true;
        previous = transformation;
      }
    }
    const M4f final_T = m4f_mul(previous, guess);
    if (out) {
      std::memcpy(out->T, final_T.m, sizeof(final_T.m));
      out->converged = converged ? 1 : 0;
      out->iterations = nr_iterations;
      out->error = 0;
      out->lm_evals = evals_total;
    }
    return 0;
  }
};

}  // namespace

extern "C" {

void orc_gicp_pcl_default_params(orc_gicp_pcl_params* p) {
  p->transformation_epsilon = 0.1;         // reg_transformation_epsilon (registrations.cpp:97)
  p->maximum_iterations = 64;              // reg_maximum_iterations (:98)
  p->use_reciprocal_correspondences = 0;   // reg_use_reciprocal_correspondences (:99); false in config/mrg_slam.yaml:106
  p->max_correspondence_distance = 2.0;    // reg_max_correspondence_distance (:100)
  p->correspondence_randomness = 20;       // reg_correspondence_randomness (:101)
  p->max_optimizer_iterations = 20;        // reg_max_optimizer_iterations (:102)
  p->rotation_epsilon = 2e-3;              // PCL default
  p->gicp_epsilon = 1e-3;                  // PCL default
}

int orc_gicp_pcl_align(const float* target, int nt, const float* source, int ns, const orc_gicp_pcl_params* p, const float* guess_colmajor,
                       orc_result* out) {
  GicpPcl g;
  g.prm = *p;
  g.target = target; g.nt = nt; g.source = source; g.ns = ns;
  return g.align(guess_colmajor, out);
}

// ---- test hooks
void orc_gicp_pcl_apply_state(const double* x6, float* T_colmajor /* in: base, out: result */) {
  M4f T;
  std::memcpy(T.m, T_colmajor, sizeof(T.m));
  Vec6 x;
  for (int i = 0; i < 6; ++i) x[i] = x6[i];
  apply_state(T, x);
  std::memcpy(T_colmajor, T.m, sizeof(T.m));
}
// cost and gradient of the functor for explicit correspondences and Mahalanobis matrices (9 per SOURCE point, row-major)
double orc_gicp_pcl_fdf(const float* src, const float* tgt, const int* idx_src, const int* idx_tgt, int m, const double* mahalanobis, int ns,
                        const double* x6, double* grad6) {
  std::vector<int> is(idx_src, idx_src + m), it(idx_tgt, idx_tgt + m);
  std::vector<double> M(mahalanobis, mahalanobis + (size_t)ns * 9);
  Problem pb;
  pb.src = src; pb.tgt = tgt; pb.idx_src = &is; pb.idx_tgt = &it; pb.mahalanobis = &M;
  Vec6 x, g;
  for (int i = 0; i < 6; ++i) x[i] = x6[i];
  double f;
  pb.fdf(x, &f, grad6 ? &g : nullptr);
  if (grad6)
    for (int i = 0; i < 6; ++i) grad6[i] = g[i];
  return f;
}
// BFGS on the same functor from x6 (in/out); returns the number of minimizeOneStep calls, status in *status
int orc_gicp_pcl_bfgs(const float* src, const float* tgt, const int* idx_src, const int* idx_tgt, int m, const double* mahalanobis, int ns,
                      double* x6, int max_inner, double gradient_tol, int* status, double* f_out, int* evals) {
  std::vector<int> is(idx_src, idx_src + m), it(idx_tgt, idx_tgt + m);
  std::vector<double> M(mahalanobis, mahalanobis + (size_t)ns * 9);
  Problem pb;
  pb.src = src; pb.tgt = tgt; pb.idx_src = &is; pb.idx_tgt = &it; pb.mahalanobis = &M;
  Vec6 x;
  for (int i = 0; i < 6; ++i) x[i] = x6[i];
  Bfgs bfgs;
  bfgs.fn = &pb;
  bfgs.minimize_init(x);
  int inner = 0, result = BFGS_RUNNING;
  do {
    inner++;
    result = bfgs.minimize_one_step(x);
    if (result) break;
    result = bfgs.test_gradient(gradient_tol);
  } while (result == BFGS_RUNNING && inner < max_inner);
  for (int i = 0; i < 6; ++i) x6[i] = x[i];
  if (status) *status = result;
  if (f_out) *f_out = bfgs.f;
  if (evals) *evals = pb.evals;
  return inner;
}
// ---- a persistent problem for the host test of the device state machine (mrg_slam_b200/csrc/gicp_pcl_sm.hpp)
struct orc_gicp_pcl { GicpPcl g; };
orc_gicp_pcl* orc_gicp_pcl_create(const float* target, int nt, const float* source, int ns, const orc_gicp_pcl_params* p,
                                  const float* guess_colmajor) {
  orc_gicp_pcl* o = new orc_gicp_pcl();
  o->g.prm = *p;
  o->g.target = target; o->g.nt = nt; o->g.source = source; o->g.ns = ns;
  if (o->g.prepare(guess_colmajor) != 0) { delete o; return nullptr; }
  return o;
}
void orc_gicp_pcl_destroy(orc_gicp_pcl* o) { delete o; }
int orc_gicp_pcl_correspond(orc_gicp_pcl* o, const float* transformation_colmajor) {
  M4f T;
  std::memcpy(T.m, transformation_colmajor, sizeof(T.m));
  return o->g.correspond(T);
}
double orc_gicp_pcl_eval(orc_gicp_pcl* o, const double* x6, double* grad6) {
  Problem pb;
  pb.src = o->g.output.data(); pb.tgt = o->g.target; pb.idx_src = &o->g.is; pb.idx_tgt = &o->g.it; pb.mahalanobis = &o->g.mahalanobis;
  Vec6 x, g;
  for (int i = 0; i < 6; ++i) x[i] = x6[i];
  double f;
  pb.fdf(x, &f, &g);
  for (int i = 0; i < 6; ++i) grad6[i] = g[i];
  return f;
}

void orc_gicp_pcl_covariances(const float* xyzi, int n, int k, double gicp_epsilon, double* cov9_out) {
  KdTree t;
  t.build(xyzi, n);
  std::vector<double> c;
  pcl_covariances(xyzi, n, t, k, gicp_epsilon, c);
  std::memcpy(cov9_out, c.data(), sizeof(double) * c.size());
}

}  // extern "C"
