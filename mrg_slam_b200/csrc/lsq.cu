// FAST_GICP / FAST_VGICP: correspondence + Mahalanobis + 6x6 accumulation kernels and the on-device
// Levenberg-Marquardt state machine; fitness score and cloud transform.
//
// Replaces: A7 FastVGICP::update_correspondences/linearize/compute_error, A9 FastGICP::*,
// A8 LsqRegistration::computeTransformation/step_lm/is_converged (fast_gicp, selected at
// src/mrg_slam/registrations.cpp:55-63,76-84), A13 pcl::Registration::getFitnessScore
// (in-tree copy: src/mrg_slam/information_matrix_calculator.cpp:46-81), A.10 pcl::transformPointCloud.
#include <cfloat>
#include <cmath>
#include <algorithm>

#include "internal.hpp"
#include "knn.cuh"

namespace b2r {

#ifndef B2R_NN1_MODE
#define B2R_NN1_MODE VISIT_PAIRS
#endif
enum { PH_LINEARIZE = 0, PH_TRIAL = 1, PH_DONE = 2 };
constexpr int kAcc = 28;   // Hrr(6) Hrt(9) Htt(6) b(6) err(1)
constexpr int kPart = 29;  // per-block partials: the kAcc sums + the number of correspondences (measurement only)

struct LsqState {
  double x0[12];  // R row-major (9), t (3)
  double xi[12];  // trial pose
  double dR[9], dt[3];
  double H[36], b[6], d[6];
  double y0, yi, lambda, nu;
  int phase, outer, inner, converged, nr_iterations, evals, failed, pad;
  double corr_last;  // correspondences of the last linearisation
  double work_pts, work_corr;  // source points / correspondences processed over all evaluations (algorithmic-bytes accounting)
};

struct LsqParams {
  int method;  // B2R_FAST_GICP / B2R_FAST_VGICP / B2R_SMALL_GICP
  int neighbor_search;
  double corr_thr2;  // max_correspondence_distance^2 (double compare as upstream)
  float corr_max_d2; // search cut-off
  double rot_eps, trans_eps;
  int max_iterations, lm_max_iterations;
  double lm_init_lambda_factor;
};

__device__ __forceinline__ void apply_pose(const double* x, double px, double py, double pz, double& ax, double& ay, double& az) {
  ax = x[0] * px + x[1] * py + x[2] * pz + x[9];
  ay = x[3] * px + x[4] * py + x[5] * pz + x[10];
  az = x[6] * px + x[7] * py + x[8] * pz + x[11];
}

// accumulate one correspondence.  M: symmetric (6) Mahalanobis matrix; a: transformed source point; e: residual.
__device__ __forceinline__ void accumulate(double* acc, const double* M, double ax, double ay, double az, double e0, double e1, double e2,
                                           double w, bool lin) {
  const double Me0 = M[0] * e0 + M[1] * e1 + M[2] * e2;
  const double Me1 = M[1] * e0 + M[3] * e1 + M[4] * e2;
  const double Me2 = M[2] * e0 + M[4] * e1 + M[5] * e2;
  acc[27] += w * (e0 * Me0 + e1 * Me1 + e2 * Me2);
  if (!lin) return;
  // J = [S | -I], S = skew(a).  Hrt = -S^T M = S M, Htt = M, and Hrr = S^T M S = -S (M S) = S (S M)^T because M S = -(S M)^T
  // (M symmetric, S skew): the product S M serves both blocks.
  const double Mf[9] = {M[0], M[1], M[2], M[1], M[3], M[4], M[2], M[4], M[5]};
  // S*X rows: [ -az X1 + ay X2 ; az X0 - ax X2 ; -ay X0 + ax X1 ]
  double SM[9];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    SM[0 * 3 + j] = -az * Mf[1 * 3 + j] + ay * Mf[2 * 3 + j];
    SM[1 * 3 + j] = az * Mf[0 * 3 + j] - ax * Mf[2 * 3 + j];
    SM[2 * 3 + j] = -ay * Mf[0 * 3 + j] + ax * Mf[1 * 3 + j];
  }
  acc[0] += w * (-az * SM[0 * 3 + 1] + ay * SM[0 * 3 + 2]);
  acc[1] += w * (-az * SM[1 * 3 + 1] + ay * SM[1 * 3 + 2]);
  acc[2] += w * (-az * SM[2 * 3 + 1] + ay * SM[2 * 3 + 2]);
  acc[3] += w * (az * SM[1 * 3 + 0] - ax * SM[1 * 3 + 2]);
  acc[4] += w * (az * SM[2 * 3 + 0] - ax * SM[2 * 3 + 2]);
  acc[5] += w * (-ay * SM[2 * 3 + 0] + ax * SM[2 * 3 + 1]);
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[6 + t] += w * SM[t];
#pragma unroll
  for (int t = 0; t < 6; ++t) acc[15 + t] += w * M[t];
  // b = J^T M e = [S^T Me ; -Me] = [-S*Me ; -Me]
  acc[21] -= w * (-az * Me1 + ay * Me2);
  acc[22] -= w * (az * Me0 - ax * Me2);
  acc[23] -= w * (-ay * Me0 + ax * Me1);
  acc[24] -= w * Me0;
  acc[25] -= w * Me1;
  acc[26] -= w * Me2;
}

// FAST_VGICP Mahalanobis matrix M = (C_B + R C_A R^T)^-1 of one source point against one voxel.  fast_gicp's covariances are
// PLANE-regularised (calculate_covariances: singular values replaced by (1, 1, 1e-3)), so C_A = V V^T - k n n^T with n the
// direction that got 1e-3, k = 1 - 1e-3 and V V^T = I to the last bit or two, hence R C_A R^T = G - k m m^T with m = R n and
// G = R R^T — NOT the identity: the pose starts from a float guess whose rotation is orthonormal to 1e-7 only, upstream uses it
// as it is, and M amplifies that by its condition number (~1e3), so G is formed once per block from the pose actually used.
// 9 + 6 + 6 multiply-adds instead of the 45 of the triple product, and 24 instead of 48 bytes of source covariance to read; the
// 3x3 inverse stays explicit.  (Sherman-Morrison on a per-voxel (C_B + I)^-1 would also remove the inverse, but it has to
// assume G = I: measured 1e-5 off in H and b.)
__device__ __forceinline__ void rrt(const double* R /*row-major*/, double* G /*xx,xy,xz,yy,yz,zz*/) {
  G[0] = R[0] * R[0] + R[1] * R[1] + R[2] * R[2];
  G[1] = R[0] * R[3] + R[1] * R[4] + R[2] * R[5];
  G[2] = R[0] * R[6] + R[1] * R[7] + R[2] * R[8];
  G[3] = R[3] * R[3] + R[4] * R[4] + R[5] * R[5];
  G[4] = R[3] * R[6] + R[4] * R[7] + R[5] * R[8];
  G[5] = R[6] * R[6] + R[7] * R[7] + R[8] * R[8];
}
template <bool COV_IN_MEMORY>
__device__ __forceinline__ void vgicp_mahalanobis(const double* R /*row-major*/, const double* G, const double* __restrict__ nrm,
                                                  const double* __restrict__ cov_b, double* M) {
  const double2 n01 = __ldg(reinterpret_cast<const double2*>(nrm));
  const double n2 = __ldg(nrm + 2);
  const double m0 = R[0] * n01.x + R[1] * n01.y + R[2] * n2;
  const double m1 = R[3] * n01.x + R[4] * n01.y + R[5] * n2;
  const double m2 = R[6] * n01.x + R[7] * n01.y + R[8] * n2;
  const double k = 1.0 - 1e-3;
  const double k0 = k * m0, k1 = k * m1, k2 = k * m2;
  double S[6];
  S[0] = ((COV_IN_MEMORY ? __ldg(&cov_b[0]) : cov_b[0]) + G[0]) - k0 * m0;
  S[1] = ((COV_IN_MEMORY ? __ldg(&cov_b[1]) : cov_b[1]) + G[1]) - k0 * m1;
  S[2] = ((COV_IN_MEMORY ? __ldg(&cov_b[2]) : cov_b[2]) + G[2]) - k0 * m2;
  S[3] = ((COV_IN_MEMORY ? __ldg(&cov_b[3]) : cov_b[3]) + G[3]) - k1 * m1;
  S[4] = ((COV_IN_MEMORY ? __ldg(&cov_b[4]) : cov_b[4]) + G[4]) - k1 * m2;
  S[5] = ((COV_IN_MEMORY ? __ldg(&cov_b[5]) : cov_b[5]) + G[5]) - k2 * m2;
  sym3_inverse(S, M);
}

// grid = (chunks, pairs).  Phase LINEARIZE: correspondences + M at x0, accumulate H, b, err.
// Phase TRIAL: err at xi with the correspondences and M of x0 (fast_gicp compute_error semantics).
template <int METHOD>
__global__ void __launch_bounds__(256, 2) lsq_eval_kernel(const CloudView* __restrict__ views, const PairDesc* __restrict__ pairs,
                                                        const LsqState* __restrict__ states, LsqParams prm, double* __restrict__ partials,
                                                        int32_t* __restrict__ corr_cache, const long long* __restrict__ corr_off,
                                                        int32_t* __restrict__ corr_out, uint8_t* __restrict__ corr_valid) {
  const int pair = blockIdx.y;
  const LsqState& st = states[pair];
  const int phase = st.phase;
  if (phase == PH_DONE) return;
  const CloudView& src = views[pairs[pair].src];
  const CloudView& tgt = views[pairs[pair].tgt];
  __shared__ double sx0[12], sxi[12], sx0t[9], sG[6];
  __shared__ double red[kAcc * 8];
  if (threadIdx.x < 12) { sx0[threadIdx.x] = st.x0[threadIdx.x]; sxi[threadIdx.x] = st.xi[threadIdx.x]; }
  if (threadIdx.x < 9) sx0t[threadIdx.x] = st.x0[(threadIdx.x % 3) * 3 + threadIdx.x / 3];  // R^T of the linearisation pose
  if (threadIdx.x == 32) rrt(st.x0, sG);
  __syncthreads();
  const bool lin = phase == PH_LINEARIZE;
  double acc[kAcc];
#pragma unroll
  for (int t = 0; t < kAcc; ++t) acc[t] = 0.0;
  int ncorr = 0;
  // FastGICP keeps correspondences_ from update_correspondences (linearize) for compute_error: position in the
  // target's cell-sorted copy or -1, one int per source point of the pair
  int32_t* cc = corr_cache ? corr_cache + corr_off[pair] : nullptr;
  float Tf[12];
  if constexpr (METHOD == B2R_FAST_GICP) {
#pragma unroll
    for (int t = 0; t < 12; ++t) Tf[t] = (float)sx0[t];
  }
  // FAST_GICP / SMALL_GICP walk the source in ITS cell order (spts) when it has one: a warp's 32 nearest-neighbour queries are
  // then neighbours in space (same target rows, similar sweep lengths), whatever order the cloud came in — the same measure that
  // took the fitness kernel from 10.1 to 7.7 ms.  `i` stays the original index (covariance, debug export); the correspondence
  // cache is indexed by the walk position, consistently in both phases.
  const bool cell_walk = METHOD != B2R_FAST_VGICP && src.spts != nullptr;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < src.n; j += gridDim.x * blockDim.x) {
    const float4 p = cell_walk ? __ldg(&src.spts[j]) : __ldg(&src.pts[j]);
    const int i = cell_walk ? __float_as_int(p.w) : j;
    const double px = (double)p.x, py = (double)p.y, pz = (double)p.z;
    double ax, ay, az;
    apply_pose(sx0, px, py, pz, ax, ay, az);
    double RCR[6];
    if constexpr (METHOD != B2R_FAST_VGICP) {
      double CA[6];
      const double* pc = src.cov + (size_t)i * 6;
#pragma unroll
      for (int t = 0; t < 6; ++t) CA[t] = __ldg(&pc[t]);
      rsrt(sx0, CA, RCR);
    }
    if constexpr (METHOD == B2R_FAST_VGICP) {
      const int vx = vgicp_coord_d(ax, tgt.vres), vy = vgicp_coord_d(ay, tgt.vres), vz = vgicp_coord_d(az, tgt.vres);
      bool first = true;
      const int noff = prm.neighbor_search == B2R_DIRECT1 ? 1 : (prm.neighbor_search == B2R_DIRECT7 ? 7 : 27);
      for (int o = 0; o < noff; ++o) {
        int ox, oy, oz;
        neighbor_offset(prm.neighbor_search, o, ox, oy, oz);
        const int cx = vx + ox - tgt.vmin[0], cy = vy + oy - tgt.vmin[1], cz = vz + oz - tgt.vmin[2];
        if (cx < 0 || cy < 0 || cz < 0 || cx >= tgt.vd[0] || cy >= tgt.vd[1] || cz >= tgt.vd[2]) continue;
        const int rec = __ldg(&tgt.v_table[(cz * tgt.vd[1] + cy) * tgt.vd[0] + cx]);
        if (rec < 0) continue;
        const VoxRec& v = tgt.vrec[rec];
        double M[6];
        vgicp_mahalanobis<true>(sx0, sG, src.nrm + (size_t)i * 4, v.cov, M);
        const double m0 = __ldg(&v.mean[0]), m1 = __ldg(&v.mean[1]), m2 = __ldg(&v.mean[2]);
        const double w = __ldg(&v.w);
        ++ncorr;
        if (lin) {
          accumulate(acc, M, ax, ay, az, m0 - ax, m1 - ay, m2 - az, w, true);
        } else {
          double bx, by, bz;
          apply_pose(sxi, px, py, pz, bx, by, bz);
          accumulate(acc, M, bx, by, bz, m0 - bx, m1 - by, m2 - bz, w, false);
        }
        if (corr_out && first) {
          corr_out[(size_t)i * 3 + 0] = vx + ox; corr_out[(size_t)i * 3 + 1] = vy + oy; corr_out[(size_t)i * 3 + 2] = vz + oz;
          corr_valid[i] = 1;
          first = false;
        }
      }
    } else if constexpr (METHOD == B2R_SMALL_GICP) {
      // small_gicp GICPFactor (registrations.cpp:46-54): nearest neighbour of T*p in DOUBLE, rejected when its squared
      // distance exceeds max_dist_sq; M = (C_B + R C_A R^T)^-1; residual r = q - T*p; J = [R skew(p) | -R];
      // H = J^T M J, b = J^T M r, e = r^T M r / 2 (the 1/2 is applied by the step kernel).  With M' = R^T M R and
      // r' = R^T r this is the same accumulation as above on [skew(p) | -I].
      int pos;
      if (!lin && cc) {
        pos = cc[j];
      } else {
        double d2;
        pos = nn1_search_d(tgt, ax, ay, az, prm.corr_thr2, d2);
        if (cc) cc[j] = pos;
      }
      if (corr_out) corr_out[i] = pos >= 0 ? __float_as_int(tgt.spts[pos].w) : -1;
      if (pos >= 0) {
        ++ncorr;
        const float4 tq = __ldg(&tgt.spts[pos]);
        const double* pcb = tgt.cov + (size_t)__float_as_int(tq.w) * 6;
        double S[6], M[6];
#pragma unroll
        for (int t = 0; t < 6; ++t) S[t] = __ldg(&pcb[t]) + RCR[t];
        sym3_inverse(S, M);
        const double m0 = (double)tq.x, m1 = (double)tq.y, m2 = (double)tq.z;
        if (lin) {
          double Mr[6];
          rsrt(sx0t, M, Mr);
          const double r0 = m0 - ax, r1 = m1 - ay, r2 = m2 - az;
          const double e0 = sx0t[0] * r0 + sx0t[1] * r1 + sx0t[2] * r2;
          const double e1 = sx0t[3] * r0 + sx0t[4] * r1 + sx0t[5] * r2;
          const double e2 = sx0t[6] * r0 + sx0t[7] * r1 + sx0t[8] * r2;
          accumulate(acc, Mr, px, py, pz, e0, e1, e2, 1.0, true);
        } else {
          double bx, by, bz;
          apply_pose(sxi, px, py, pz, bx, by, bz);
          accumulate(acc, M, bx, by, bz, m0 - bx, m1 - by, m2 - bz, 1.0, false);
        }
      }
    } else {
      // FastGICP::update_correspondences: float transform ((c0 x + c1 y) + c2 z) + c3, exact 1-NN, d2 < thr^2
      float q[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        float s = __fmul_rn(Tf[r * 3 + 0], p.x);
        s = __fadd_rn(s, __fmul_rn(Tf[r * 3 + 1], p.y));
        s = __fadd_rn(s, __fmul_rn(Tf[r * 3 + 2], p.z));
        q[r] = __fadd_rn(s, Tf[9 + r]);
      }
      int pos;
      bool ok;
      if (!lin && cc) {
        pos = cc[j];
        ok = pos >= 0;
      } else {
        float d2;
        pos = nn1_search<B2R_NN1_MODE>(tgt, q[0], q[1], q[2], prm.corr_max_d2, d2);
        ok = pos >= 0 && (double)d2 < prm.corr_thr2;
        if (cc) cc[j] = ok ? pos : -1;
      }
      if (corr_out) corr_out[i] = ok ? __float_as_int(tgt.spts[pos].w) : -1;
      if (ok) {
        ++ncorr;
        const float4 tq = __ldg(&tgt.spts[pos]);
        const double* pcb = tgt.cov + (size_t)__float_as_int(tq.w) * 6;
        double S[6], M[6];
#pragma unroll
        for (int t = 0; t < 6; ++t) S[t] = __ldg(&pcb[t]) + RCR[t];
        sym3_inverse(S, M);
        const double m0 = (double)tq.x, m1 = (double)tq.y, m2 = (double)tq.z;
        if (lin) {
          accumulate(acc, M, ax, ay, az, m0 - ax, m1 - ay, m2 - az, 1.0, true);
        } else {
          double bx, by, bz;
          apply_pose(sxi, px, py, pz, bx, by, bz);
          accumulate(acc, M, bx, by, bz, m0 - bx, m1 - by, m2 - bz, 1.0, false);
        }
      }
    }
  }
  double* out = partials + ((size_t)pair * gridDim.x + blockIdx.x) * kPart;
  if (lin) {
    block_reduce_butterfly<kAcc>(acc, red, out);
    const int bc = block_sum_int(ncorr, (int*)red);
    if (threadIdx.x == 0) out[28] = (double)bc;
  } else {
    double e[1] = {acc[27]};
    block_reduce_to<1>(e, red, out + 27);
  }
}

// ---- FAST_VGICP with DIRECT1 lookup (the only mode mrg_slam reaches: registrations.cpp:76-84 calls no search-method setter):
// same arithmetic as lsq_eval_kernel<B2R_FAST_VGICP>, software-pipelined and compacted per warp.  A point costs three
// DEPENDENT memory round trips (point -> table cell -> voxel record) and the 28 f64 accumulators leave room for only 16 warps
// per SM, so the plain loop is bound by that latency chain; and with a far initial guess (loop-closure candidates) half of the
// source points fall into empty voxels, whose lanes would idle through the Mahalanobis / Hessian arithmetic.  Every thread
// keeps three points in flight (s = points per round of the grid):
//   i + 2s : point load issued
//   i + 1s : point has arrived -> transform, voxel coordinate, table probe issued
//   i      : probe has arrived -> a hit is appended to the WARP's ring in shared memory (ballot + popc, point order, no block
//            barrier); its voxel record and the point's covariance are prefetched into L1 (CCTL.PF1, no registers)
//   ring   : as soon as it holds 32 hits: Mahalanobis + accumulation, one hit per lane, every operand an L1 hit
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ int vgicp_probe(const CloudView& tgt, const double* x, const float4& p) {
  double ax, ay, az;
  apply_pose(x, (double)p.x, (double)p.y, (double)p.z, ax, ay, az);
  const double res = tgt.vres;
  int cx, cy, cz;
  if (res == 1.0) {  // the reference's reg_resolution (config/mrg_slam.yaml:108): x / 1.0 == x exactly, no division needed
    cx = (int)floor(ax - 0.5); cy = (int)floor(ay - 0.5); cz = (int)floor(az - 0.5);
  } else {
    cx = vgicp_coord_d(ax, res); cy = vgicp_coord_d(ay, res); cz = vgicp_coord_d(az, res);
  }
  cx -= tgt.vmin[0]; cy -= tgt.vmin[1]; cz -= tgt.vmin[2];
  if (cx < 0 || cy < 0 || cz < 0 || cx >= tgt.vd[0] || cy >= tgt.vd[1] || cz >= tgt.vd[2]) return -1;
  return __ldg(&tgt.v_table[(cz * tgt.vd[1] + cy) * tgt.vd[0] + cx]);
}

// One correspondence of the VGICP cost: source point i of the pair against voxel record rec.
// SEL 0: phase decided at run time (28 accumulators); 1: linearisation only; 2: trial only (acc[0] is the error sum).
template <int SEL>
__device__ __forceinline__ void vgicp_point(double* acc, const CloudView& src, const CloudView& tgt, const double* sx0, const double* sxi,
                                            const double* sG, int i, int rec, bool lin) {
  const float4 p = __ldg(&src.pts[i]);
  const double px = (double)p.x, py = (double)p.y, pz = (double)p.z;
  const VoxRec& v = tgt.vrec[rec];
  double M[6];
  double m0, m1, m2, w;
  {
    // the record's 80 hot bytes as five 16-byte loads (half the requests of ten 8-byte ones in the L1 data pipe), in two stages so
    // that the linearisation kernel, at its 128-register limit, does not hold all twenty registers at once
    const double2* q = reinterpret_cast<const double2*>(&v);
    const double2 r0 = __ldg(q), r1 = __ldg(q + 1), r2 = __ldg(q + 2);
    const double cb[6] = {r0.x, r0.y, r1.x, r1.y, r2.x, r2.y};
    vgicp_mahalanobis<false>(sx0, sG, src.nrm + (size_t)i * 4, cb, M);
    if (SEL != 2) asm volatile("" ::: "memory");
    const double2 r3 = __ldg(q + 3), r4 = __ldg(q + 4);
    m0 = r3.x; m1 = r3.y; m2 = r4.x; w = r4.y;
  }
  if constexpr (SEL == 2) {
    double bx, by, bz;
    apply_pose(sxi, px, py, pz, bx, by, bz);
    const double e0 = m0 - bx, e1 = m1 - by, e2 = m2 - bz;
    const double Me0 = M[0] * e0 + M[1] * e1 + M[2] * e2;
    const double Me1 = M[1] * e0 + M[3] * e1 + M[4] * e2;
    const double Me2 = M[2] * e0 + M[4] * e1 + M[5] * e2;
    acc[0] += w * (e0 * Me0 + e1 * Me1 + e2 * Me2);
  } else if (SEL == 1 || lin) {
    double ax, ay, az;
    apply_pose(sx0, px, py, pz, ax, ay, az);
    accumulate(acc, M, ax, ay, az, m0 - ax, m1 - ay, m2 - az, w, true);
  } else {
    double bx, by, bz;
    apply_pose(sxi, px, py, pz, bx, by, bz);
    accumulate(acc, M, bx, by, bz, m0 - bx, m1 - by, m2 - bz, w, false);
  }
}

constexpr int kVgRing = 64;  // per warp: <= 31 left over + 32 new
// SEL 0: one kernel for both phases (a pair's blocks follow its phase); 1 / 2: only the pairs in the linearisation / trial phase
// (the other pairs' blocks exit at once).  The trial pass needs one accumulator instead of 28 and runs with four blocks per SM
// (64 registers) instead of two, so a round launches the two specialisations back to back instead of the common kernel
// (B2R_VGICP_SPLIT=0 restores that): 22.9 -> 21.8 ms per 4096-pair step.  Three blocks for the linearisation (80 registers, 220 B
// of spills) lost: 24.3 ms.
#ifndef B2R_VG_TRIAL_BLOCKS
#define B2R_VG_TRIAL_BLOCKS 4
#endif
#ifndef B2R_VG_LIN_BLOCKS
#define B2R_VG_LIN_BLOCKS 2
#endif
// THREADS: 256 for large batches; 128 (the same registers per thread, twice the blocks per SM) when a launch has few blocks — smaller
// blocks leave shorter tails: 3.28 -> 3.20 ms per 512-pair step, but 19.8 -> 20.0 ms at 4096 pairs.
template <int SEL, int THREADS = 256>
__global__ void __launch_bounds__(THREADS, (SEL == 2 ? B2R_VG_TRIAL_BLOCKS : B2R_VG_LIN_BLOCKS) * (256 / THREADS)) vgicp_eval_kernel(const CloudView* __restrict__ views, const PairDesc* __restrict__ pairs,
                                                                          const LsqState* __restrict__ states, double* __restrict__ partials) {
  const int pair = blockIdx.y;
  const LsqState& st = states[pair];
  const int phase = st.phase;
  if (phase == PH_DONE) return;
  if (SEL == 1 && phase != PH_LINEARIZE) return;
  if (SEL == 2 && phase != PH_TRIAL) return;
  const CloudView& src = views[pairs[pair].src];
  const CloudView& tgt = views[pairs[pair].tgt];
  constexpr int NA = SEL == 2 ? 1 : kAcc;
  __shared__ double sx0[12], sxi[12], sG[6];
  __shared__ double red[NA * 8];
  __shared__ int2 s_ring[8][kVgRing];  // (source point, voxel record) hits of each warp
  if (threadIdx.x < 12) { sx0[threadIdx.x] = st.x0[threadIdx.x]; sxi[threadIdx.x] = st.xi[threadIdx.x]; }
  if (threadIdx.x == 32) rrt(st.x0, sG);
  __syncthreads();
  const bool lin = SEL == 1 || (SEL == 0 && phase == PH_LINEARIZE);
  double acc[NA];
#pragma unroll
  for (int t = 0; t < NA; ++t) acc[t] = 0.0;
  int ncorr = 0;
  const int n = src.n, s = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  int2* ring = s_ring[threadIdx.x >> 5];
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  int head = 0, tail = 0;  // ring positions (warp-uniform)
  // prologue: fill the pipeline
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
  float4 p2 = i0 + s < n ? __ldg(&src.pts[i0 + s]) : zero4;
  int rec1 = i0 < n ? vgicp_probe(tgt, sx0, __ldg(&src.pts[i0])) : -1;
  const int wbase0 = i0 - lane;  // first point of the warp: the trip count below is the same for all its lanes
  for (int wbase = wbase0; wbase < n; wbase += s) {
    const int i = wbase + lane;
    const float4 p3 = i + 2 * s < n ? __ldg(&src.pts[i + 2 * s]) : zero4;
    const int rec2 = i + s < n ? vgicp_probe(tgt, sx0, p2) : -1;
    // ---- append this round's hits of the warp
    const bool hit = rec1 >= 0;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      ring[(tail + __popc(m & ((1u << lane) - 1u))) & (kVgRing - 1)] = make_int2(i, rec1);
      const char* vr = (const char*)&tgt.vrec[rec1];
#ifndef B2R_VG_PREFETCH_LIN
#define B2R_VG_PREFETCH_LIN 0
#endif
      // Prefetching the hit's voxel record and normal into L1 (CCTL.PF1) paid when one kernel served both phases at two blocks
      // per SM; with the split kernels it costs more L1 requests than it hides latency (measured per 4096-pair step: 20.83 ms
      // with the prefetches, 20.54 without).
      if (SEL != 2 && B2R_VG_PREFETCH_LIN) {
        prefetch_l1(vr); prefetch_l1(vr + 32); prefetch_l1(vr + 64); prefetch_l1(vr + 79);  // mean, cov, w: the first 80 bytes
        prefetch_l1(src.nrm + (size_t)i * 4);
      }
    }
    tail += __popc(m);
    __syncwarp();
    // ---- drain: 32 hits at a time (and the remainder after the warp's last points)
    const bool last = wbase + s >= n;
    while (tail - head >= 32 || (last && tail - head > 0)) {
      const int pos = head + lane;
      if (pos < tail) {
        const int2 it = ring[pos & (kVgRing - 1)];
        vgicp_point<SEL>(acc, src, tgt, sx0, sxi, sG, it.x, it.y, lin);
        ++ncorr;
      }
      head = min(head + 32, tail);
      __syncwarp();
    }
    p2 = p3;
    rec1 = rec2;
  }
  double* out = partials + ((size_t)pair * gridDim.x + blockIdx.x) * kPart;
  if constexpr (SEL == 2) {
    block_reduce_to<1>(acc, red, out + 27);
  } else {
    if (lin) {
      block_reduce_butterfly<kAcc>(acc, red, out);
      const int bc = block_sum_int(ncorr, (int*)red);
      if (threadIdx.x == 0) out[28] = (double)bc;
    } else {
      double e[1] = {acc[27]};
      block_reduce_to<1>(e, red, out + 27);
    }
  }
}

// ---- small dense math for the step kernel (one thread per pair) ----
__device__ void ldlt6_solve_dev(const double* Ain, const double* rhs, double* x) {
  double A[36];
  for (int i = 0; i < 36; ++i) A[i] = Ain[i];
  int perm[6] = {0, 1, 2, 3, 4, 5};
  for (int k = 0; k < 6; ++k) {
    int piv = k;
    double best = fabs(A[k * 6 + k]);
    for (int i = k + 1; i < 6; ++i)
      if (fabs(A[i * 6 + i]) > best) { best = fabs(A[i * 6 + i]); piv = i; }
    if (piv != k) {
      for (int j = 0; j < 6; ++j) { double t = A[k * 6 + j]; A[k * 6 + j] = A[piv * 6 + j]; A[piv * 6 + j] = t; }
      for (int i = 0; i < 6; ++i) { double t = A[i * 6 + k]; A[i * 6 + k] = A[i * 6 + piv]; A[i * 6 + piv] = t; }
      int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
    }
    const double dk = A[k * 6 + k];
    if (dk == 0.0) continue;
    for (int i = k + 1; i < 6; ++i) A[i * 6 + k] /= dk;
    for (int i = k + 1; i < 6; ++i)
      for (int j = k + 1; j <= i; ++j) {
        A[i * 6 + j] -= A[i * 6 + k] * dk * A[j * 6 + k];
        A[j * 6 + i] = A[i * 6 + j];
      }
  }
  double y[6];
  for (int i = 0; i < 6; ++i) y[i] = rhs[perm[i]];
  for (int i = 0; i < 6; ++i)
    for (int k = 0; k < i; ++k) y[i] -= A[i * 6 + k] * y[k];
  for (int i = 0; i < 6; ++i) y[i] = (A[i * 6 + i] != 0.0) ? y[i] / A[i * 6 + i] : 0.0;
  for (int i = 5; i >= 0; --i)
    for (int k = i + 1; k < 6; ++k) y[i] -= A[k * 6 + i] * y[k];
  for (int i = 0; i < 6; ++i) x[perm[i]] = y[i];
}

__device__ void so3_exp_dev(const double* om, double* R) {
  const double theta_sq = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  double imag, real;
  if (theta_sq < 1e-10) {
    const double theta_quad = theta_sq * theta_sq;
    imag = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    const double theta = sqrt(theta_sq), half = 0.5 * theta;
    imag = sin(half) / theta;
    real = cos(half);
  }
  const double w = real, x = imag * om[0], y = imag * om[1], z = imag * om[2];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

__device__ bool lsq_is_converged(const LsqState& s, const LsqParams& prm) {
  double m = 0.0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) m = fmax(m, fabs(s.dR[r * 3 + c] - (r == c ? 1.0 : 0.0)) / prm.rot_eps);
  for (int r = 0; r < 3; ++r) m = fmax(m, fabs(s.dt[r]) / prm.trans_eps);
  return m < 1.0;
}

// d = solve(H + lambda I, -b); delta = [so3_exp(d0..2) | d3..5]; xi = delta * x0
__device__ void lsq_propose(LsqState& s) {
  double A[36], nb[6];
  for (int i = 0; i < 36; ++i) A[i] = s.H[i];
  for (int j = 0; j < 6; ++j) { A[j * 6 + j] += s.lambda; nb[j] = -s.b[j]; }
  ldlt6_solve_dev(A, nb, s.d);
  so3_exp_dev(s.d, s.dR);
  s.dt[0] = s.d[3]; s.dt[1] = s.d[4]; s.dt[2] = s.d[5];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) s.xi[r * 3 + c] = s.dR[r * 3 + 0] * s.x0[0 * 3 + c] + s.dR[r * 3 + 1] * s.x0[1 * 3 + c] + s.dR[r * 3 + 2] * s.x0[2 * 3 + c];
    s.xi[9 + r] = s.dR[r * 3 + 0] * s.x0[9] + s.dR[r * 3 + 1] * s.x0[10] + s.dR[r * 3 + 2] * s.x0[11] + s.dt[r];
  }
}

// small_gicp se3_exp (Sophus form): rotation so3_exp(omega), translation V(omega) * v
__device__ void se3_exp_dev(const double* a, double* R, double* t) {
  const double th2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
  const double th = sqrt(th2);
  so3_exp_dev(a, R);
  if (th < 1e-10) {
    for (int r = 0; r < 3; ++r) t[r] = R[r * 3 + 0] * a[3] + R[r * 3 + 1] * a[4] + R[r * 3 + 2] * a[5];
    return;
  }
  const double Om[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
  double Om2[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Om2[r * 3 + c] = Om[r * 3 + 0] * Om[0 * 3 + c] + Om[r * 3 + 1] * Om[1 * 3 + c] + Om[r * 3 + 2] * Om[2 * 3 + c];
  const double c1 = (1.0 - cos(th)) / th2, c2 = (th - sin(th)) / (th2 * th);
  for (int r = 0; r < 3; ++r) {
    double acc = 0.0;
    for (int c = 0; c < 3; ++c) acc += (((r == c) ? 1.0 : 0.0) + c1 * Om[r * 3 + c] + c2 * Om2[r * 3 + c]) * a[3 + c];
    t[r] = acc;
  }
}

// small_gicp LevenbergMarquardtOptimizer: d = solve(H + lambda I, -b); xi = x0 * se3_exp(d)  (right-multiplied update)
__device__ void sg_propose(LsqState& s) {
  double A[36], nb[6];
  for (int i = 0; i < 36; ++i) A[i] = s.H[i];
  for (int j = 0; j < 6; ++j) { A[j * 6 + j] += s.lambda; nb[j] = -s.b[j]; }
  ldlt6_solve_dev(A, nb, s.d);
  se3_exp_dev(s.d, s.dR, s.dt);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) s.xi[r * 3 + c] = s.x0[r * 3 + 0] * s.dR[0 * 3 + c] + s.x0[r * 3 + 1] * s.dR[1 * 3 + c] + s.x0[r * 3 + 2] * s.dR[2 * 3 + c];
    s.xi[9 + r] = s.x0[r * 3 + 0] * s.dt[0] + s.x0[r * 3 + 1] * s.dt[1] + s.x0[r * 3 + 2] * s.dt[2] + s.x0[9 + r];
  }
}

__device__ void lsq_finish(LsqState& s, int* done_count) {
  s.phase = PH_DONE;
  atomicAdd(done_count, 1);
}

// getFinalTransformation / hasConverged / nr_iterations_ of a finished pair, as the row the C ABI hands out (written on the
// device: for a sharded batch the row array IS the all-gather send buffer)
__device__ void lsq_write_row(const LsqState& s, b2r_result& r) {
  for (int rr = 0; rr < 3; ++rr) {
    for (int c = 0; c < 3; ++c) r.T[c * 4 + rr] = (float)s.x0[rr * 3 + c];
    r.T[12 + rr] = (float)s.x0[9 + rr];
    r.T[rr * 4 + 3] = 0.f;
  }
  r.T[15] = 1.f;
  r.converged = s.converged;
  r.iterations = s.nr_iterations;
  r.error = s.y0;
  r.evals = s.evals;
  clear_row_padding(r);
  r.fitness = 0.0;
}

// One warp per pair: fixed-order sum of the chunk partials, then the LM state machine of
// LsqRegistration::computeTransformation / step_lm advanced by one evaluation.
__device__ void lsq_step_body(LsqState* __restrict__ states, int npairs, const LsqParams& prm, const double* __restrict__ partials, int chunks,
                              const int* __restrict__ src_n, int* __restrict__ done_count) {
  const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pair >= npairs) return;
  LsqState& s = states[pair];
  const int phase = s.phase;
  if (phase == PH_DONE) return;
  double v = 0.0;
  if (lane < kPart && (phase == PH_LINEARIZE || lane == 27)) {
    const double* p = partials + (size_t)pair * chunks * kPart + lane;
    for (int c = 0; c < chunks; ++c) v += p[(size_t)c * kPart];
  }
  const double ncorr = __shfl_sync(0xffffffffu, v, 28);
  double a[kAcc];
#pragma unroll
  for (int t = 0; t < kAcc; ++t) a[t] = __shfl_sync(0xffffffffu, v, t);
  if (lane != 0) return;
  s.evals++;
  if (phase == PH_LINEARIZE) s.corr_last = ncorr;
  s.work_pts += (double)src_n[pair];
  s.work_corr += s.corr_last;
  if (prm.method == B2R_SMALL_GICP) {
    // LevenbergMarquardtOptimizer::optimize (small_gicp; init_lambda 1e-3, lambda_factor 10, max_inner_iterations =
    // lm_max_iterations) advanced by one evaluation.  s.outer = i, s.inner = j, s.y0 = e.
    constexpr double kLambdaFactor = 10.0;
    if (phase == PH_LINEARIZE) {
      const int rr[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
      for (int t = 0; t < 6; ++t) {
        s.H[rr[t][0] * 6 + rr[t][1]] = a[t];
        s.H[rr[t][1] * 6 + rr[t][0]] = a[t];
        s.H[(3 + rr[t][0]) * 6 + 3 + rr[t][1]] = a[15 + t];
        s.H[(3 + rr[t][1]) * 6 + 3 + rr[t][0]] = a[15 + t];
      }
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          s.H[r * 6 + 3 + c] = a[6 + r * 3 + c];
          s.H[(3 + c) * 6 + r] = a[6 + r * 3 + c];
        }
      for (int t = 0; t < 6; ++t) s.b[t] = a[21 + t];
      s.y0 = 0.5 * a[27];
      s.nr_iterations = s.outer;
      s.inner = 0;
      if (prm.lm_max_iterations <= 0) { lsq_finish(s, done_count); return; }  // no inner iteration: success stays false
      sg_propose(s);
      s.phase = PH_TRIAL;
      return;
    }
    s.yi = 0.5 * a[27];
    if (s.yi <= s.y0) {
      const double rn = sqrt(s.d[0] * s.d[0] + s.d[1] * s.d[1] + s.d[2] * s.d[2]);
      const double tn = sqrt(s.d[3] * s.d[3] + s.d[4] * s.d[4] + s.d[5] * s.d[5]);
      s.converged = (rn <= prm.rot_eps && tn <= prm.trans_eps) ? 1 : 0;
      for (int t = 0; t < 12; ++t) s.x0[t] = s.xi[t];
      s.lambda /= kLambdaFactor;
      s.y0 = s.yi;
      s.nr_iterations = s.outer;
      s.outer++;
      if (s.converged || s.outer >= prm.max_iterations) { lsq_finish(s, done_count); return; }
      s.phase = PH_LINEARIZE;
      return;
    }
    s.lambda *= kLambdaFactor;
    s.inner++;
    if (s.inner >= prm.lm_max_iterations) { lsq_finish(s, done_count); return; }  // !success: break, converged stays false
    sg_propose(s);
    return;
  }
  if (phase == PH_LINEARIZE) {
    // unpack the symmetric system
    const int rr[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
    for (int t = 0; t < 6; ++t) {
      s.H[rr[t][0] * 6 + rr[t][1]] = a[t];
      s.H[rr[t][1] * 6 + rr[t][0]] = a[t];
      s.H[(3 + rr[t][0]) * 6 + 3 + rr[t][1]] = a[15 + t];
      s.H[(3 + rr[t][1]) * 6 + 3 + rr[t][0]] = a[15 + t];
    }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        s.H[r * 6 + 3 + c] = a[6 + r * 3 + c];
        s.H[(3 + c) * 6 + r] = a[6 + r * 3 + c];
      }
    for (int t = 0; t < 6; ++t) s.b[t] = a[21 + t];
    s.y0 = a[27];
    s.nr_iterations = s.outer;
    if (s.lambda < 0.0) {
      double mx = 0.0;
      for (int i = 0; i < 6; ++i) mx = fmax(mx, fabs(s.H[i * 6 + i]));
      s.lambda = prm.lm_init_lambda_factor * mx;
    }
    s.nu = 2.0;
    s.inner = 0;
    if (prm.lm_max_iterations <= 0) {  // step_lm's loop body never runs: "lm not converged!!"
      s.failed = 1;
      lsq_finish(s, done_count);
      return;
    }
    lsq_propose(s);
    s.phase = PH_TRIAL;
    return;
  }
  // ---- PH_TRIAL
  s.yi = a[27];
  double denom = 0.0;
  for (int j = 0; j < 6; ++j) denom += s.d[j] * (s.lambda * s.d[j] - s.b[j]);
  const double rho = (s.y0 - s.yi) / denom;
  bool step_ok;
  if (rho < 0) {
    if (lsq_is_converged(s, prm)) {
      step_ok = true;  // x0 unchanged
    } else {
      s.lambda = s.nu * s.lambda;
      s.nu = 2 * s.nu;
      s.inner++;
      if (s.inner >= prm.lm_max_iterations) {
        s.failed = 1;  // "lm not converged!!" -> break out of the outer loop, converged_ stays false
        lsq_finish(s, done_count);
        return;
      }
      lsq_propose(s);
      return;  // another trial with the same H, b
    }
  } else {
    for (int t = 0; t < 12; ++t) s.x0[t] = s.xi[t];
    const double f = 1 - pow(2 * rho - 1, 3);
    const double third = 1.0 / 3.0;
    s.lambda = s.lambda * (third < f ? f : third);
    step_ok = true;
  }
  (void)step_ok;
  s.converged = lsq_is_converged(s, prm) ? 1 : 0;
  s.nr_iterations = s.outer;
  s.outer++;
  if (s.converged || s.outer >= prm.max_iterations) {
    lsq_finish(s, done_count);
    return;
  }
  s.phase = PH_LINEARIZE;
}

__global__ void lsq_step_kernel(LsqState* __restrict__ states, LsqParams prm, const double* __restrict__ partials, int chunks,
                                const int* __restrict__ src_n, LoopArgs la) {
  lsq_step_body(states, la.npairs, prm, partials, chunks, src_n, &la.ctl->done);
  loop_tail(la);
}
__global__ void lsq_rows_kernel(const LsqState* __restrict__ states, int npairs, b2r_result* __restrict__ rows) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair < npairs) lsq_write_row(states[pair], rows[pair]);
}

__global__ void lsq_init_kernel(LsqState* __restrict__ states, int npairs, const float* __restrict__ guesses, int max_iterations,
                                double init_lambda, int* __restrict__ done_count) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= npairs) return;
  LsqState& s = states[pair];
  const float* g = guesses + (size_t)pair * 16;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) s.x0[r * 3 + c] = (double)g[c * 4 + r];
    s.x0[9 + r] = (double)g[12 + r];
  }
  for (int t = 0; t < 12; ++t) s.xi[t] = s.x0[t];
  for (int t = 0; t < 9; ++t) s.dR[t] = (t % 4 == 0) ? 1.0 : 0.0;
  s.dt[0] = s.dt[1] = s.dt[2] = 0.0;
  s.y0 = s.yi = 0.0;
  s.lambda = init_lambda;  // < 0: fast_gicp derives it from the first Hessian; small_gicp starts at 1e-3
  s.nu = 2.0;
  s.phase = PH_LINEARIZE;
  s.outer = 0; s.inner = 0; s.converged = 0; s.nr_iterations = 0; s.evals = 0; s.failed = 0;
  s.corr_last = 0.0; s.work_pts = 0.0; s.work_corr = 0.0;
  if (max_iterations <= 0) { s.phase = PH_DONE; atomicAdd(done_count, 1); }
}

static void launch_lsq_eval(Ctx& ctx, int method, dim3 grid, const CloudView* views, const PairDesc* pairs, const LsqState* states,
                            const LsqParams& prm, double* partials, int32_t* corr_cache, const long long* corr_off, int32_t* corr_out,
                            uint8_t* corr_valid) {
  static const bool pipelined = [] { const char* e = getenv("B2R_VGICP_PIPE"); return !e || atoi(e) != 0; }();
  if (method == B2R_FAST_VGICP && prm.neighbor_search == B2R_DIRECT1 && !corr_out && pipelined)
    B2R_LAUNCH(ctx, vgicp_eval_kernel<0>, grid, 256, 0, views, pairs, states, partials);
  else if (method == B2R_FAST_VGICP)
    B2R_LAUNCH(ctx, lsq_eval_kernel<B2R_FAST_VGICP>, grid, 256, 0, views, pairs, states, prm, partials, corr_cache, corr_off, corr_out, corr_valid);
  else if (method == B2R_FAST_GICP)
    B2R_LAUNCH(ctx, lsq_eval_kernel<B2R_FAST_GICP>, grid, 256, 0, views, pairs, states, prm, partials, corr_cache, corr_off, corr_out, corr_valid);
  else
    B2R_LAUNCH(ctx, lsq_eval_kernel<B2R_SMALL_GICP>, grid, 256, 0, views, pairs, states, prm, partials, corr_cache, corr_off, corr_out, corr_valid);
}

static LsqParams make_params(const b2r_config& cfg) {
  LsqParams p;
  p.method = cfg.method;
  p.neighbor_search = cfg.method == B2R_FAST_VGICP ? cfg.neighbor_search : B2R_DIRECT1;
  p.corr_thr2 = cfg.max_correspondence_distance * cfg.max_correspondence_distance;
  p.corr_max_d2 = p.corr_thr2 >= (double)FLT_MAX ? INFINITY : (float)(p.corr_thr2 * 1.0001);
  if (cfg.method == B2R_SMALL_GICP && !(p.corr_thr2 < 1e30)) p.corr_thr2 = INFINITY;
  p.rot_eps = cfg.rotation_epsilon;
  p.trans_eps = cfg.transformation_epsilon;
  p.max_iterations = cfg.maximum_iterations;
  p.lm_max_iterations = cfg.lm_max_iterations;
  p.lm_init_lambda_factor = cfg.lm_init_lambda_factor;
  return p;
}

static int pick_chunks(const Ctx& ctx, int npairs, int maxn, bool uniform_cost = false) {
  int by_size = std::max(1, (maxn + 1023) / 1024);
  static const int fill = [] { const char* e = getenv("B2R_FILL_PER_SM"); return e ? atoi(e) : 32; }();  // blocks per SM a launch should offer: short tails when few pairs are active
  int by_fill = std::max(1, (fill * ctx.num_sms + npairs - 1) / npairs);
  // large batches: still split every pair into blocks of <= ~8k points, so that a pair whose correspondence search is slow
  // (far guess) cannot leave the last wave of the launch to a few long-running blocks
  static const int tail_pts = [] { const char* e = getenv("B2R_CHUNK_POINTS"); return e ? atoi(e) : 8192; }();
  // (a VGICP pass costs the same for every point — one table probe — so its blocks may be 4x larger: fewer block reductions)
  const int tp = uniform_cost ? 4 * tail_pts : tail_pts;
  by_fill = std::max(by_fill, (maxn + tp - 1) / tp);
  return std::max(1, std::min(by_size, by_fill));
}

void lsq_align_batch(Ctx& ctx, const b2r_config& cfg, const BatchArgs& b) {
  const int np = b.np;
  if (np == 0) return;
  const int chunks = pick_chunks(ctx, np, b.maxn, cfg.method == B2R_FAST_VGICP);
  LsqParams prm = make_params(cfg);
  DBuf<LsqState> ds; ds.alloc(np, ctx.stream);
  DBuf<double> part; part.alloc((size_t)np * chunks * kPart, ctx.stream);
  DBuf<LoopCtl> ctl; ctl.alloc(1, ctx.stream);
  ctl.zero(ctx.stream);
  // FAST_GICP / SMALL_GICP: correspondence cache, one int per source point of every pair
  DBuf<int32_t> corr;
  DBuf<long long> coff;
  std::vector<long long> hoff;
  if (cfg.method == B2R_FAST_GICP || cfg.method == B2R_SMALL_GICP) {
    hoff.resize(np);
    long long tot = 0;
    for (int i = 0; i < np; ++i) { hoff[i] = tot; tot += b.src_sizes[i]; }
    corr.alloc((size_t)std::max(1ll, tot), ctx.stream);
    coff.alloc(np, ctx.stream);
    B2R_CUDA(cudaMemcpyAsync(coff.p, hoff.data(), sizeof(long long) * np, cudaMemcpyHostToDevice, ctx.stream));
  }
  B2R_CUDA(cudaMemsetAsync(ds.p, 0, sizeof(LsqState) * np, ctx.stream));
  B2R_LAUNCH(ctx, lsq_init_kernel, (np + 127) / 128, 128, 0, ds.p, np, b.d_guesses, cfg.maximum_iterations,
             cfg.method == B2R_SMALL_GICP ? 1e-3 : -1.0, &ctl.p->done);
  const long max_rounds = (long)std::max(1, cfg.maximum_iterations) * (1 + std::max(1, cfg.lm_max_iterations)) + 2;
  // ---- the whole LM iteration as one device-side loop
  const CloudView* a_views = b.d_views;
  const PairDesc* a_pairs = b.d_pairs;
  const LsqState* a_states_c = ds.p;
  LsqState* a_states = ds.p;
  double* a_part = part.p;
  const double* a_part_c = part.p;
  int32_t* a_corr = corr.p;
  const long long* a_coff = coff.p;
  int32_t* a_null_i = nullptr;
  uint8_t* a_null_b = nullptr;
  int a_chunks = chunks;
  const int* a_src_n = b.d_src_n;
  LoopArgs la;
  memset(&la, 0, sizeof(la));
  static const bool pipelined = [] { const char* e = getenv("B2R_VGICP_PIPE"); return !e || atoi(e) != 0; }();
  const bool pipe = cfg.method == B2R_FAST_VGICP && prm.neighbor_search == B2R_DIRECT1 && pipelined;
  void* eval_args_pipe[] = {&a_views, &a_pairs, &a_states_c, &a_part};
  void* eval_args_gen[] = {&a_views, &a_pairs, &a_states_c, &prm, &a_part, &a_corr, &a_coff, &a_null_i, &a_null_b};
  void* step_args[] = {&a_states, &prm, &a_part_c, &a_chunks, &a_src_n, &la};
  static const bool split = [] { const char* e = getenv("B2R_VGICP_SPLIT"); return !e || atoi(e) != 0; }();
  static const int small_batch = [] { const char* e = getenv("B2R_VG_SMALL_BATCH"); return e ? atoi(e) : 640; }();  // pairs per launch up to which 128-thread blocks are used (measured: better at 512, worse from 1024 on)
  const bool small = pipe && split && np <= small_batch;
  const void* eval2_fn = (pipe && split) ? (small ? (const void*)vgicp_eval_kernel<2, 128> : (const void*)vgicp_eval_kernel<2>) : nullptr;
  const void* eval_fn = pipe ? (split ? (small ? (const void*)vgicp_eval_kernel<1, 128> : (const void*)vgicp_eval_kernel<1>) : (const void*)vgicp_eval_kernel<0>)
                             : (cfg.method == B2R_FAST_VGICP ? (const void*)lsq_eval_kernel<B2R_FAST_VGICP>
                                : (cfg.method == B2R_FAST_GICP ? (const void*)lsq_eval_kernel<B2R_FAST_GICP> : (const void*)lsq_eval_kernel<B2R_SMALL_GICP>));
  run_device_loop(ctx, eval_fn, dim3(chunks, np), dim3(small ? 128 : 256), pipe ? eval_args_pipe : eval_args_gen, (const void*)lsq_step_kernel,
                  dim3((np + 3) / 4), dim3(128), step_args, la, ctl.p, np, max_rounds, PROF_LSQ_EVAL, eval2_fn);
  B2R_LAUNCH(ctx, lsq_rows_kernel, (np + 127) / 128, 128, 0, ds.p, np, b.d_rows);
  if (ctx.profile) {
    // per evaluation pass 40 B per source point + one 64 B voxel record (VGICP) or 40 B target
    // point + covariance (GICP) per correspondence, summed over the passes each pair actually ran
    std::vector<LsqState> hs(np);
    B2R_CUDA(cudaMemcpyAsync(hs.data(), ds.p, sizeof(LsqState) * np, cudaMemcpyDeviceToHost, ctx.stream));
    B2R_CUDA(cudaStreamSynchronize(ctx.stream));
    double wp = 0.0, wc = 0.0;
    for (int i = 0; i < np; ++i) { wp += hs[i].work_pts; wc += hs[i].work_corr; }
    ctx.prof_bytes[PROF_LSQ_EVAL] += 40.0 * wp + (cfg.method == B2R_FAST_VGICP ? 64.0 : 40.0) * wc;
  }
}

void lsq_debug_linearize(Ctx& ctx, const b2r_config& cfg, const CloudView* d_views, int n_src, const double* T_lin, const double* T_trial,
                         bool trial, double* H, double* b, double* err, int32_t* corr_out, uint8_t* corr_valid) {
  const LsqParams prm = make_params(cfg);
  LsqState s;
  memset(&s, 0, sizeof(s));
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) { s.x0[r * 3 + c] = T_lin[r * 4 + c]; s.xi[r * 3 + c] = T_trial[r * 4 + c]; }
    s.x0[9 + r] = T_lin[r * 4 + 3];
    s.xi[9 + r] = T_trial[r * 4 + 3];
  }
  s.phase = trial ? PH_TRIAL : PH_LINEARIZE;
  const int chunks = pick_chunks(ctx, 1, n_src);
  PairDesc pd{0, 1};
  DBuf<PairDesc> dp; dp.alloc(1, ctx.stream);
  DBuf<LsqState> ds; ds.alloc(1, ctx.stream);
  DBuf<double> part; part.alloc((size_t)chunks * kPart, ctx.stream);
  part.zero(ctx.stream);
  DBuf<int32_t> dc;
  DBuf<uint8_t> dvld;
  const bool vg = cfg.method == B2R_FAST_VGICP;
  if (corr_out) {
    dc.alloc((size_t)n_src * (vg ? 3 : 1), ctx.stream);
    dc.zero(ctx.stream);
    dvld.alloc((size_t)n_src, ctx.stream);
    dvld.zero(ctx.stream);
  }
  B2R_CUDA(cudaMemcpyAsync(dp.p, &pd, sizeof(pd), cudaMemcpyHostToDevice, ctx.stream));
  B2R_CUDA(cudaMemcpyAsync(ds.p, &s, sizeof(s), cudaMemcpyHostToDevice, ctx.stream));
  launch_lsq_eval(ctx, cfg.method, dim3(chunks, 1), d_views, dp.p, ds.p, prm, part.p, nullptr, nullptr, corr_out ? dc.p : nullptr,
                  corr_out ? dvld.p : nullptr);
  std::vector<double> hp((size_t)chunks * kPart);
  B2R_CUDA(cudaMemcpyAsync(hp.data(), part.p, sizeof(double) * hp.size(), cudaMemcpyDeviceToHost, ctx.stream));
  if (corr_out) {
    B2R_CUDA(cudaMemcpyAsync(corr_out, dc.p, sizeof(int32_t) * n_src * (vg ? 3 : 1), cudaMemcpyDeviceToHost, ctx.stream));
    if (vg && corr_valid) B2R_CUDA(cudaMemcpyAsync(corr_valid, dvld.p, n_src, cudaMemcpyDeviceToHost, ctx.stream));
  }
  B2R_CUDA(cudaStreamSynchronize(ctx.stream));
  double a[kAcc] = {0};
  for (int c = 0; c < chunks; ++c)
    for (int t = 0; t < kAcc; ++t) a[t] += hp[(size_t)c * kPart + t];
  *err = a[27];
  if (H && !trial) {
    const int rr[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
    for (int t = 0; t < 6; ++t) {
      H[rr[t][0] * 6 + rr[t][1]] = H[rr[t][1] * 6 + rr[t][0]] = a[t];
      H[(3 + rr[t][0]) * 6 + 3 + rr[t][1]] = H[(3 + rr[t][1]) * 6 + 3 + rr[t][0]] = a[15 + t];
    }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) H[r * 6 + 3 + c] = H[(3 + c) * 6 + r] = a[6 + r * 3 + c];
    for (int t = 0; t < 6; ++t) b[t] = a[21 + t];
  }
}

// ------------------------------------------------------------------------------------------------ transform + fitness
// pcl::transformPointCloud, SSE association (x*c0 + y*c1) + (z*c2 + c3)  
__device__ __forceinline__ void pcl_transform(const float* T /*col-major 16*/, float x, float y, float z, float& ox, float& oy, float& oz) {
  ox = __fadd_rn(__fadd_rn(__fmul_rn(x, T[0]), __fmul_rn(y, T[4])), __fadd_rn(__fmul_rn(z, T[8]), T[12]));
  oy = __fadd_rn(__fadd_rn(__fmul_rn(x, T[1]), __fmul_rn(y, T[5])), __fadd_rn(__fmul_rn(z, T[9]), T[13]));
  oz = __fadd_rn(__fadd_rn(__fmul_rn(x, T[2]), __fmul_rn(y, T[6])), __fadd_rn(__fmul_rn(z, T[10]), T[14]));
}

struct Mat16 { float m[16]; };

__global__ void transform_kernel(const float4* __restrict__ in, int n, Mat16 T, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = in[i];
  float4 o;
  pcl_transform(T.m, p.x, p.y, p.z, o.x, o.y, o.z);
  o.w = p.w;
  out[i] = o;
}

// views[0] = source, views[1] = target: the nearestKSearch(point, 1) answers of getFitnessScore's loop, all at once
__global__ void nn_export_kernel(const CloudView* __restrict__ views, Mat16 T, int32_t* __restrict__ idx, float* __restrict__ d2o,
                                 float* __restrict__ xyz) {
  const CloudView& src = views[0];
  const CloudView& tgt = views[1];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= src.n) return;
  const float4 p = __ldg(&src.pts[i]);
  float qx, qy, qz;
  pcl_transform(T.m, p.x, p.y, p.z, qx, qy, qz);
  float d2;
  const int pos = nn1_search<B2R_NN1_MODE>(tgt, qx, qy, qz, INFINITY, d2);
  idx[i] = pos >= 0 ? __float_as_int(tgt.spts[pos].w) : -1;
  d2o[i] = d2;
  if (xyz) { xyz[3 * (size_t)i] = qx; xyz[3 * (size_t)i + 1] = qy; xyz[3 * (size_t)i + 2] = qz; }
}
void nearest_neighbors(Ctx& ctx, const CloudView* d_views, int n_src, const float* T_colmajor, int32_t* d_idx, float* d_d2, float* d_xyz) {
  if (n_src == 0) return;
  Mat16 T;
  memcpy(T.m, T_colmajor, sizeof(T.m));
  B2R_LAUNCH(ctx, nn_export_kernel, (n_src + 127) / 128, 128, 0, d_views, T, d_idx, d_d2, d_xyz);
}

void transform_cloud(Ctx& ctx, const float4* in, int n, const float* T_colmajor, float4* out) {
  if (n == 0) return;
  Mat16 T;
  memcpy(T.m, T_colmajor, sizeof(T.m));
  B2R_LAUNCH(ctx, transform_kernel, (n + 255) / 256, 256, 0, in, n, T, out);
}

// grid = (chunks, pairs): per source point exact 1-NN squared distance into the target; sum of those <= max_range
// (getFitnessScore, A13) and, when inlier_d2 > 0, the number of points whose nearest neighbour is closer than that (A15: the
// inlier fraction of ScanMatchingOdometryComponent::publish_scan_matching_status, scan_matching_odometry_component.cpp:403-415).
// The transform is the pair's result row (written on the device by the optimiser: no host round trip in between).
// Measured and rejected for the far queries (4 % of the source points of a loop-closure pair have their nearest target point
// beyond one grid cell; 73 % of the warps hold at least one): (a) abandoning a query after a budget of candidate tests and
// searching it warp-cooperatively (32 rows per step, strips swept with coalesced loads): 10.1 -> 14.5-38 ms per 4096-pair batch,
// the strips of a far query are many and short; (b) queueing the abandoned queries per block / per warp and draining them one per
// lane: 12.1-20 ms — ncu shows the budgeted first phase alone costs what the plain kernel costs (7.0 G of 7.4 G warp
// instructions): the time is in the ORDINARY queries' short, divergent sweeps (5-6 of 32 lanes in the distance tests), not in a
// few expensive ones (profiles/r2/ncu_fitness_batch4096_*.md).
#ifndef B2R_FIT_BLOCKS
#define B2R_FIT_BLOCKS 4
#endif
template <int MODE, bool LEAN>
__global__ void __launch_bounds__(256, B2R_FIT_BLOCKS) fitness_kernel(const CloudView* __restrict__ views, const PairDesc* __restrict__ pairs,
                                                       const b2r_result* __restrict__ rows, double max_range, float max_d2, float inlier_d2,
                                                       int cell_order, double* __restrict__ partials) {
  const int pair = blockIdx.y;
  const CloudView& src = views[pairs[pair].src];
  const CloudView& tgt = views[pairs[pair].tgt];
  // the queries of a warp should be neighbours in space (same target rows, similar sweep lengths): the source's own cell-sorted
  // copy gives that whatever the order the cloud came in; the sum is order-independent up to the last bits of a double
  const float4* __restrict__ qpts = (cell_order && src.spts) ? src.spts : src.pts;
  __shared__ float T[16];
  __shared__ double red[3 * 8];
  if (threadIdx.x < 16) T[threadIdx.x] = rows[pair].T[threadIdx.x];
  __syncthreads();
  double acc[3] = {0.0, 0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < src.n; i += gridDim.x * blockDim.x) {
    const float4 p = __ldg(&qpts[i]);
    float qx, qy, qz;
    pcl_transform(T, p.x, p.y, p.z, qx, qy, qz);
    float d2;
    if constexpr (LEAN) d2 = nn1_dist<MODE>(tgt, qx, qy, qz, max_d2);
    else if (nn1_search<MODE>(tgt, qx, qy, qz, max_d2, d2) < 0) d2 = INFINITY;
    if (d2 < INFINITY && (double)d2 <= max_range) { acc[0] += (double)d2; acc[1] += 1.0; }
    if (d2 < inlier_d2) acc[2] += 1.0;
  }
  block_reduce_to<3>(acc, red, partials + ((size_t)pair * gridDim.x + blockIdx.x) * 3);
}
__global__ void fitness_finish_kernel(const double* __restrict__ partials, int npairs, int chunks, b2r_result* __restrict__ rows,
                                      int* __restrict__ inliers) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= npairs) return;
  double s = 0.0, c = 0.0, in = 0.0;
  for (int k = 0; k < chunks; ++k) {
    const double* p = partials + ((size_t)pair * chunks + k) * 3;
    s += p[0]; c += p[1]; in += p[2];
  }
  rows[pair].fitness = c > 0.0 ? s / c : DBL_MAX;
  if (inliers) inliers[pair] = (int)in;
}

void fitness_batch(Ctx& ctx, const BatchArgs& b, double max_range, float inlier_d2, int* d_inlier_out) {
  const int np = b.np;
  if (np == 0) return;
  static const int fit_chunk_mul = [] { const char* e = getenv("B2R_FIT_CHUNK_MUL"); return e ? atoi(e) : 1; }();
  const int chunks = pick_chunks(ctx, np, b.maxn) * fit_chunk_mul;
  DBuf<double> part; part.alloc((size_t)np * chunks * 3, ctx.stream);
  float max_d2 = max_range >= (double)FLT_MAX ? INFINITY : (float)(max_range * 1.0001);
  if (inlier_d2 > 0.f) max_d2 = fmaxf(max_d2, inlier_d2 * 1.0001f);
  {
    double pts = 0.0;  // 16 B source point + one gathered 16 B neighbour
    for (int i = 0; i < np; ++i) pts += b.src_sizes[i];
    ProfScope ps(ctx, PROF_FITNESS, 32.0 * pts);
    static const int cell_order = [] { const char* e = getenv("B2R_FIT_CELL_ORDER"); return e ? atoi(e) : 1; }();
    static const int mode = [] { const char* e = getenv("B2R_FIT_VISIT"); return e ? atoi(e) : (int)VISIT_PAIRS; }();
    static const int lean = [] { const char* e = getenv("B2R_FIT_LEAN"); return e ? atoi(e) : 1; }();
    static const int fit_threads = [] { const char* e = getenv("B2R_FIT_THREADS"); return e ? atoi(e) : 256; }();
#define B2R_FIT_LAUNCH(M, L) \
    B2R_LAUNCH(ctx, (fitness_kernel<M, L>), dim3(chunks, np), fit_threads, 0, b.d_views, b.d_pairs, b.d_rows, max_range, max_d2, inlier_d2, cell_order, part.p)
    if (mode == VISIT_PAIRS) { if (lean) B2R_FIT_LAUNCH(VISIT_PAIRS, true); else B2R_FIT_LAUNCH(VISIT_PAIRS, false); }
    else if (mode == VISIT_MERGED) { if (lean) B2R_FIT_LAUNCH(VISIT_MERGED, true); else B2R_FIT_LAUNCH(VISIT_MERGED, false); }
    else if (mode == (VISIT_CELL3 | VISIT_LANE_RING)) { if (lean) B2R_FIT_LAUNCH(VISIT_CELL3 | VISIT_LANE_RING, true); else B2R_FIT_LAUNCH(VISIT_CELL3 | VISIT_LANE_RING, false); }
    else { if (lean) B2R_FIT_LAUNCH(VISIT_CELL3, true); else B2R_FIT_LAUNCH(VISIT_CELL3, false); }
#undef B2R_FIT_LAUNCH
  }
  B2R_LAUNCH(ctx, fitness_finish_kernel, (np + 127) / 128, 128, 0, part.p, np, chunks, b.d_rows, d_inlier_out);
}

}  // namespace b2r
