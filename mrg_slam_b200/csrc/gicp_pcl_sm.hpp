// pcl::GeneralizedIterativeClosestPoint ("GICP" / "GICP_OMP", /root/reference/src/mrg_slam/registrations.cpp:93-116; SURVEY 8a
// row G) as a resumable state machine: the outer correspondence loop of computeTransformation, estimateRigidTransformationBFGS
// and pcl/registration/bfgs.h (GSL vector_bfgs2 + Fletcher's line search) cut at every point where upstream evaluates the
// cost functor or searches correspondences.  One state per pair lives in device memory; an evaluation kernel serves the pending
// request of every active pair (correspondences, or cost + gradient sums at a 6-vector x), a step kernel feeds the answer to
// gp_advance(), which runs until the next request.  So the optimiser never leaves the device, as for the other methods.
//
// The header is plain C++ (host and device): tests/test_gicp_pcl_sm.py compiles it for the host and drives it with the
// ORACLE's own correspondence search and functor (oracle/gicp_pcl.cpp), where it must reproduce orc_gicp_pcl_align bit for bit
// — the state machine is checked without a GPU; only the two kernels that answer the requests are CUDA-specific.
#pragma once
#include <cmath>
#include <cfloat>

#if defined(__CUDACC__)
#define GP_HD __host__ __device__
#else
#define GP_HD
#endif

namespace gp {

enum Request { REQ_NONE = 0, REQ_CORRESPOND = 1, REQ_EVAL = 2, REQ_DONE = 3 };
enum { ST_SUCCESS = 0, ST_RUNNING = 1, ST_NOPROGRESS = 2 };
// where the machine waits
enum Phase {
  PH_START = 0,      // before the first correspondence search
  PH_CORR,           // waiting for correspondences of an outer iteration
  PH_INIT,           // waiting for fdf(x) of BFGS::minimizeInit
  PH_BRACKET_F,      // line search, bracketing: waiting for f(alpha)
  PH_BRACKET_DF,     // ... for f'(alpha)
  PH_SECTION_F,      // line search, sectioning: waiting for f(alpha)
  PH_SECTION_DF,     // ... for f'(alpha)
  PH_UPDATE,         // waiting for fdf at the accepted alpha (updatePosition)
  PH_FINISHED
};

struct Params {
  double transformation_epsilon, rotation_epsilon;
  int maximum_iterations, max_optimizer_iterations;
  double gradient_tol;  // 1e-2 in gicp.hpp
  double rho, sigma, tau1, tau2, tau3, step_size;
  int order, bracket_iters, sect_iters;
};
GP_HD inline void default_params(Params& p) {
  p.transformation_epsilon = 0.1; p.rotation_epsilon = 2e-3;
  p.maximum_iterations = 64; p.max_optimizer_iterations = 20;
  p.gradient_tol = 1e-2;
  p.rho = 0.01; p.sigma = 0.01; p.tau1 = 9; p.tau2 = 0.05; p.tau3 = 0.5; p.step_size = 1.0;
  p.order = 3; p.bracket_iters = 100; p.sect_iters = 100;
}

struct State {
  // ---- request to the evaluation kernel
  int request;
  double xreq[6];            // REQ_EVAL: the state vector to evaluate at
  float transformation[16];  // REQ_CORRESPOND: transformation_ (column-major float); queries are transformation_ * (guess * p)
  float guess[16];
  // ---- outer loop (computeTransformation)
  float previous[16];
  int nr_iterations, converged, phase, evals, inner_total;
  // ---- BFGS (bfgs.h)
  int inner;
  double f, delta_f, fp0, pnorm, g0norm;
  double x[6], x0[6], g0[6], p[6], gradient[6];
  // ---- line search (linear_minimize.c: minimize)
  double ls_f0, ls_fp0, alpha, alpha_prev, falpha, falpha_prev, fpalpha_prev, a, b, fa, fb, fpa, fpb;
  int ls_i;
  double step_f0;  // f before the step (minimizeOneStep's f0)
};

GP_HD inline double dot6(const double* a, const double* b) {
  double s = 0;
  for (int i = 0; i < 6; ++i) s += a[i] * b[i];
  return s;
}
GP_HD inline double norm6(const double* a) { return sqrt(dot6(a, a)); }

// ---- float 4x4 column-major helpers (Eigen::Matrix4f)
GP_HD inline void m4f_identity(float* T) {
  for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.f : 0.f;
}
GP_HD inline void m4f_mul(const float* A, const float* B, float* C) {
  float R[16];
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      float s = 0.f;
      for (int k = 0; k < 4; ++k) s += A[k * 4 + r] * B[c * 4 + k];
      R[c * 4 + r] = s;
    }
  for (int i = 0; i < 16; ++i) C[i] = R[i];
}
// GICP::applyState on an identity base: T = [Rz(x5) Ry(x4) Rx(x3) | x0..2], float
GP_HD inline void apply_state_identity(const double* x, float* T) {
  const float a = (float)x[3], b = (float)x[4], c = (float)x[5];
  const float ca = cosf(a), sa = sinf(a), cb = cosf(b), sb = sinf(b), cc = cosf(c), sc = sinf(c);
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
  // R * I: the product with the identity block is evaluated like upstream (sum of three terms, two of them exact zeros)
  m4f_identity(T);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float s = 0.f;
      for (int k = 0; k < 3; ++k) s += R[i * 3 + k] * (k == j ? 1.f : 0.f);
      T[j * 4 + i] = s;
    }
  T[12] += (float)x[0];
  T[13] += (float)x[1];
  T[14] += (float)x[2];
}

// GICP::computeRDerivative: g[3..5] from D = dCost_dR_T (row-major 3x3), R = Rz(psi) Ry(theta) Rx(phi)
GP_HD inline void compute_r_derivative(const double* x, const double* D, double* g) {
  const double phi = x[3], theta = x[4], psi = x[5];
  const double cphi = cos(phi), sphi = sin(phi), ctheta = cos(theta), stheta = sin(theta), cpsi = cos(psi), spsi = sin(psi);
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
  const double* A[3] = {dPhi, dTheta, dPsi};
  for (int t = 0; t < 3; ++t) {
    double r = 0.;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) r += A[t][j * 3 + i] * D[i * 3 + j];
    g[3 + t] = r;
  }
}

// The sums an evaluation returns: fs = sum d^T M d, gt = sum M d, D = sum p_src (M d)^T (row-major), m = correspondences.
// Turns them into the functor's f and g (OptimizationFunctorWithIndices::fdf).
GP_HD inline void finish_fdf(const double* x, double fs, const double* gt, const double* Dsum, int m, double* f, double* g) {
  *f = fs / m;
  for (int a = 0; a < 3; ++a) g[a] = gt[a] * (2.0 / m);
  double D[9];
  for (int a = 0; a < 9; ++a) D[a] = Dsum[a] * (2.0 / m);
  compute_r_derivative(x, D, g);
}

// ---- line-search interpolation (linear_minimize.c)
GP_HD inline int solve_quadratic(double a, double b, double c, double* x0, double* x1) {
  if (a == 0) {
    if (b == 0) return 0;
    *x0 = -c / b;
    return 1;
  }
  const double disc = b * b - 4 * a * c;
  if (disc > 0) {
    if (b == 0) {
      const double r = sqrt(-c / a);
      *x0 = -r;
      *x1 = r;
    } else {
      const double sgnb = b > 0 ? 1 : -1;
      const double temp = -0.5 * (b + sgnb * sqrt(disc));
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
GP_HD inline double interp_quad(double f0, double fp0, double f1, double zl, double zh) {
  const double fl = f0 + zl * (fp0 + zl * (f1 - f0 - fp0));
  const double fh = f0 + zh * (fp0 + zh * (f1 - f0 - fp0));
  const double c = 2 * (f1 - f0 - fp0);
  double zmin = zl, fmin = fl;
  if (fh < fmin) { zmin = zh; fmin = fh; }
  if (c > 0) {
    const double z = -fp0 / c;
    if (z > zl && z < zh) {
      const double f = f0 + z * (fp0 + z * (f1 - f0 - fp0));
      if (f < fmin) { zmin = z; fmin = f; }
    }
  }
  return zmin;
}
GP_HD inline double cubic(double c0, double c1, double c2, double c3, double z) { return c0 + z * (c1 + z * (c2 + z * c3)); }
GP_HD inline void check_extremum(double c0, double c1, double c2, double c3, double z, double* zmin, double* fmin) {
  const double y = cubic(c0, c1, c2, c3, z);
  if (y < *fmin) { *zmin = z; *fmin = y; }
}
GP_HD inline double interp_cubic(double f0, double fp0, double f1, double fp1, double zl, double zh) {
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
GP_HD inline double interpolate(double a, double fa, double fpa, double b, double fb, double fpb, double xmin, double xmax, int order) {
  double zmin = (xmin - a) / (b - a), zmax = (xmax - a) / (b - a);
  if (zmin > zmax) { const double t = zmin; zmin = zmax; zmax = t; }
  double z;
  if (order > 2 && !(fpb != fpb)) z = interp_cubic(fa, fpa * (b - a), fb, fpb * (b - a), zmin, zmax);
  else z = interp_quad(fa, fpa * (b - a), fb, zmin, zmax);
  return a + z * (b - a);
}

// ---- the machine ------------------------------------------------------------------------------------------------------
GP_HD inline void request_eval_along(State& s, double alpha) {
  for (int i = 0; i < 6; ++i) s.xreq[i] = s.x0[i] + alpha * s.p[i];
  s.request = REQ_EVAL;
}
GP_HD inline void request_correspondences(State& s) {
  s.request = REQ_CORRESPOND;
  s.phase = PH_CORR;
}
GP_HD inline void finish(State& s) {
  s.request = REQ_DONE;
  s.phase = PH_FINISHED;
}

GP_HD inline void init(State& s, const float* guess_colmajor) {
  for (int i = 0; i < 16; ++i) s.guess[i] = guess_colmajor[i];
  m4f_identity(s.transformation);
  m4f_identity(s.previous);
  s.nr_iterations = 0; s.converged = 0; s.evals = 0; s.inner_total = 0; s.inner = 0;
  s.f = 0; s.delta_f = 0; s.fp0 = 0; s.pnorm = 0; s.g0norm = 0;
  for (int i = 0; i < 6; ++i) { s.x[i] = s.x0[i] = s.g0[i] = s.p[i] = s.gradient[i] = s.xreq[i] = 0; }
  s.ls_f0 = s.ls_fp0 = s.alpha = s.alpha_prev = s.falpha = s.falpha_prev = s.fpalpha_prev = 0;
  s.a = s.b = s.fa = s.fb = s.fpa = s.fpb = 0;
  s.ls_i = 0;
  s.step_f0 = 0;
  s.phase = PH_START;
  s.request = REQ_NONE;
}

// final_transformation_ = previous_transformation_ * guess
GP_HD inline void final_transformation(const State& s, float* T) { m4f_mul(s.previous, s.guess, T); }

// estimateRigidTransformationBFGS returned (ok) or threw (!ok): the rest of one pass of computeTransformation's while loop
GP_HD inline void outer_iteration_done(State& s, const Params& prm, bool ok) {
  if (!ok) {  // exception caught: break, converged_ stays false; previous_transformation_ = transformation_ before the call
    finish(s);
    return;
  }
  double delta = 0.;
  for (int k = 0; k < 4; ++k)
    for (int l = 0; l < 4; ++l) {
      const double ratio = (k < 3 && l < 3) ? 1. / prm.rotation_epsilon : 1. / prm.transformation_epsilon;
      const double c_delta = ratio * fabs((double)s.previous[l * 4 + k] - (double)s.transformation[l * 4 + k]);
      if (c_delta > delta) delta = c_delta;
    }
  s.nr_iterations++;
  if (s.nr_iterations >= prm.maximum_iterations || delta < 1) {
    s.converged = 1;
    for (int i = 0; i < 16; ++i) s.previous[i] = s.transformation[i];
    finish(s);
    return;
  }
  request_correspondences(s);
}

// the inner do-while of estimateRigidTransformationBFGS after minimizeOneStep returned `result`
GP_HD inline void inner_step_done(State& s, const Params& prm, int result) {
  for (;;) {
    bool leave = result != ST_SUCCESS;
    if (!leave) {
      result = norm6(s.gradient) < prm.gradient_tol ? ST_SUCCESS : ST_RUNNING;  // testGradient
      leave = !(result == ST_RUNNING && s.inner < prm.max_optimizer_iterations);
    }
    if (leave) break;
    // next minimizeOneStep
    s.inner++;
    s.step_f0 = s.f;
    if (s.pnorm == 0.0 || s.g0norm == 0.0 || s.fp0 == 0) {
      result = ST_NOPROGRESS;
      continue;
    }
    double alpha1;
    if (s.delta_f < 0) {
      double del = -s.delta_f;
      const double lim = 10 * DBL_EPSILON * fabs(s.step_f0);
      if (lim > del) del = lim;
      alpha1 = 2.0 * del / (-s.fp0);
      if (alpha1 > 1.0) alpha1 = 1.0;
    } else {
      alpha1 = fabs(prm.step_size);
    }
    // lineSearch: set up and ask for f(alpha1)
    s.ls_f0 = s.f; s.ls_fp0 = s.fp0;
    s.falpha_prev = s.ls_f0; s.fpalpha_prev = s.ls_fp0;
    s.alpha = alpha1; s.alpha_prev = 0;
    s.a = 0; s.b = s.alpha; s.fa = s.ls_f0; s.fb = 0; s.fpa = s.ls_fp0; s.fpb = 0;
    s.ls_i = 0;
    s.ls_i++;  // while (i++ < bracket_iters): upstream's 100 is never exhausted on entry
    s.phase = PH_BRACKET_F;
    request_eval_along(s, s.alpha);
    return;
  }
  s.inner_total += s.inner;
  if (result == ST_NOPROGRESS || result == ST_SUCCESS || s.inner == prm.max_optimizer_iterations) {
    apply_state_identity(s.x, s.transformation);
    outer_iteration_done(s, prm, true);
  } else {
    outer_iteration_done(s, prm, false);
  }
}

// sectioning step: compute the next trial alpha in [a, b] and ask for f there, or give up the line search
GP_HD inline void section_next(State& s, const Params& prm) {
  if (s.ls_i++ < prm.sect_iters) {
    const double delta = s.b - s.a;
    const double lower = s.a + prm.tau2 * delta, upper = s.b - prm.tau3 * delta;
    s.alpha = interpolate(s.a, s.fa, s.fpa, s.b, s.fb, s.fpb, lower, upper, prm.order);
    s.phase = PH_SECTION_F;
    request_eval_along(s, s.alpha);
    return;
  }
  // out of iterations: minimize() returns SUCCESS with alpha_new untouched (= 0 in minimizeOneStep)
  s.alpha = 0.0;
  s.phase = PH_UPDATE;
  request_eval_along(s, 0.0);
}

// line search accepted alpha: updatePosition needs f and g there
GP_HD inline void accept_alpha(State& s, double alpha) {
  s.alpha = alpha;
  s.phase = PH_UPDATE;
  request_eval_along(s, alpha);
}

// Feed the answer to the pending request.  REQ_CORRESPOND: m = number of correspondences (the rest ignored).
// REQ_EVAL: f, g = functor value and gradient at xreq.
GP_HD inline void advance(State& s, const Params& prm, int m, double f, const double* g) {
  const double kNaN = nan("");
  switch (s.phase) {
    case PH_START:
      request_correspondences(s);
      return;
    case PH_CORR: {
      for (int i = 0; i < 16; ++i) s.previous[i] = s.transformation[i];  // previous_transformation_ = transformation_
      if (m < 4) { outer_iteration_done(s, prm, false); return; }
      // estimateRigidTransformationBFGS: initial x from transformation_
      const float* T = s.transformation;
      s.x[0] = T[12]; s.x[1] = T[13]; s.x[2] = T[14];
      s.x[3] = (double)atan2f(T[1 * 4 + 2], T[2 * 4 + 2]);  // std::atan2(T(2,1), T(2,2)) on floats
      s.x[4] = (double)asinf(-T[0 * 4 + 2]);                 // asin(-T(2,0))
      s.x[5] = (double)atan2f(T[0 * 4 + 1], T[0 * 4 + 0]);  // std::atan2(T(1,0), T(0,0))
      for (int i = 0; i < 6; ++i) s.xreq[i] = s.x[i];
      s.request = REQ_EVAL;
      s.phase = PH_INIT;
      return;
    }
    case PH_INIT: {  // BFGS::minimizeInit
      s.evals++;
      s.delta_f = 0;
      s.f = f;
      for (int i = 0; i < 6; ++i) { s.gradient[i] = g[i]; s.x0[i] = s.x[i]; s.g0[i] = g[i]; }
      s.g0norm = norm6(s.g0);
      for (int i = 0; i < 6; ++i) s.p[i] = s.gradient[i] * -1 / s.g0norm;
      s.pnorm = norm6(s.p);
      s.fp0 = -s.g0norm;
      s.inner = 0;
      // do { inner++; result = minimizeOneStep(x); ... } — the first pass through the loop body is unconditional
      s.inner = 1;
      s.step_f0 = s.f;
      if (s.pnorm == 0.0 || s.g0norm == 0.0 || s.fp0 == 0) { inner_step_done(s, prm, ST_NOPROGRESS); return; }
      {
        const double alpha1 = fabs(prm.step_size);  // delta_f == 0 on the first step
        s.ls_f0 = s.f; s.ls_fp0 = s.fp0;
        s.falpha_prev = s.ls_f0; s.fpalpha_prev = s.ls_fp0;
        s.alpha = alpha1; s.alpha_prev = 0;
        s.a = 0; s.b = s.alpha; s.fa = s.ls_f0; s.fb = 0; s.fpa = s.ls_fp0; s.fpb = 0;
        s.ls_i = 0;
        s.ls_i++;
        s.phase = PH_BRACKET_F;
        request_eval_along(s, s.alpha);
      }
      return;
    }
    case PH_BRACKET_F: {
      s.evals++;
      s.falpha = f;
      // Fletcher's rho test
      if (s.falpha > s.ls_f0 + s.alpha * prm.rho * s.ls_fp0 || s.falpha >= s.falpha_prev) {
        s.a = s.alpha_prev; s.fa = s.falpha_prev; s.fpa = s.fpalpha_prev;
        s.b = s.alpha; s.fb = s.falpha; s.fpb = kNaN;
        section_next(s, prm);
        return;
      }
      s.phase = PH_BRACKET_DF;
      request_eval_along(s, s.alpha);  // upstream's df(alpha): the gradient at the same point
      return;
    }
    case PH_BRACKET_DF: {
      s.evals++;
      const double fpalpha = dot6(g, s.p);
      // Fletcher's sigma test
      if (fabs(fpalpha) <= -prm.sigma * s.ls_fp0) { accept_alpha(s, s.alpha); return; }
      if (fpalpha >= 0) {
        s.a = s.alpha; s.fa = s.falpha; s.fpa = fpalpha;
        s.b = s.alpha_prev; s.fb = s.falpha_prev; s.fpb = s.fpalpha_prev;
        section_next(s, prm);
        return;
      }
      const double delta = s.alpha - s.alpha_prev;
      const double lower = s.alpha + delta, upper = s.alpha + prm.tau1 * delta;
      const double alpha_next = interpolate(s.alpha_prev, s.falpha_prev, s.fpalpha_prev, s.alpha, s.falpha, fpalpha, lower, upper, prm.order);
      s.alpha_prev = s.alpha; s.falpha_prev = s.falpha; s.fpalpha_prev = fpalpha;
      s.alpha = alpha_next;
      if (s.ls_i++ < prm.bracket_iters) {
        s.phase = PH_BRACKET_F;
        request_eval_along(s, s.alpha);
        return;
      }
      section_next(s, prm);  // bracketing ran out of iterations: sectioning of [a, b] as it stands
      return;
    }
    case PH_SECTION_F: {
      s.evals++;
      s.falpha = f;
      if ((s.a - s.alpha) * s.fpa <= DBL_EPSILON) {  // roundoff prevents progress
        inner_step_done(s, prm, ST_NOPROGRESS);
        return;
      }
      if (s.falpha > s.ls_f0 + prm.rho * s.alpha * s.ls_fp0 || s.falpha >= s.fa) {
        s.b = s.alpha; s.fb = s.falpha; s.fpb = kNaN;
        section_next(s, prm);
        return;
      }
      s.phase = PH_SECTION_DF;
      request_eval_along(s, s.alpha);
      return;
    }
    case PH_SECTION_DF: {
      s.evals++;
      const double fpalpha = dot6(g, s.p);
      if (fabs(fpalpha) <= -prm.sigma * s.ls_fp0) { accept_alpha(s, s.alpha); return; }
      if (((s.b - s.a) >= 0 && fpalpha >= 0) || ((s.b - s.a) <= 0 && fpalpha <= 0)) {
        s.b = s.a; s.fb = s.fa; s.fpb = s.fpa;
        s.a = s.alpha; s.fa = s.falpha; s.fpa = fpalpha;
      } else {
        s.a = s.alpha; s.fa = s.falpha; s.fpa = fpalpha;
      }
      section_next(s, prm);
      return;
    }
    case PH_UPDATE: {  // updatePosition + the BFGS direction update of minimizeOneStep
      s.evals++;
      double xn[6];
      for (int i = 0; i < 6; ++i) xn[i] = s.x0[i] + s.alpha * s.p[i];
      for (int i = 0; i < 6; ++i) { s.x[i] = xn[i]; s.gradient[i] = g[i]; }
      s.f = f;
      s.delta_f = s.f - s.step_f0;
      double dx0[6], dg0[6];
      for (int i = 0; i < 6; ++i) { dx0[i] = s.x[i] - s.x0[i]; dg0[i] = s.gradient[i] - s.g0[i]; }
      const double dxg = dot6(dx0, s.gradient), dgg = dot6(dg0, s.gradient), dxdg = dot6(dx0, dg0), dgnorm = norm6(dg0);
      double A, B;
      if (dxdg != 0) {
        B = dxg / dxdg;
        A = -(1.0 + dgnorm * dgnorm / dxdg) * B + dgg / dxdg;
      } else {
        B = 0;
        A = 0;
      }
      for (int i = 0; i < 6; ++i) s.p[i] = -A * dx0[i] + s.gradient[i] + -B * dg0[i];
      for (int i = 0; i < 6; ++i) { s.g0[i] = s.gradient[i]; s.x0[i] = s.x[i]; }
      s.g0norm = norm6(s.g0);
      s.pnorm = norm6(s.p);
      const double dir = (dot6(s.p, s.gradient) > 0) ? -1.0 : 1.0;
      for (int i = 0; i < 6; ++i) s.p[i] *= dir / s.pnorm;
      s.pnorm = norm6(s.p);
      s.fp0 = dot6(s.p, s.g0);
      inner_step_done(s, prm, ST_SUCCESS);
      return;
    }
    default:
      return;
  }
}

}  // namespace gp
