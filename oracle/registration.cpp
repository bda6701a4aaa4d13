// ORACLE — TEST INFRASTRUCTURE ONLY.  Parity unpinned (see oracle/README.md).
//
// CPU (C++/OpenMP, no Eigen) restatement of the three registration classes the
// reference's factory hands out for the north-star methods:
//   FAST_GICP  -> fast_gicp::FastGICP   (src/mrg_slam/registrations.cpp:55-63)
//   FAST_VGICP -> fast_gicp::FastVGICP  (src/mrg_slam/registrations.cpp:76-84)
//   NDT_OMP    -> pclomp::NormalDistributionsTransform (registrations.cpp:130-147)
// and of the pcl::Registration base behaviour their callers rely on
// (apps/scan_matching_odometry_component.cpp:203-275, src/mrg_slam/loop_detector.cpp:104-144).
// None of those libraries is vendored in /root/reference; the algorithms follow
// SURVEY.md Appendix A.0-A.4 (upstream fast_gicp / ndt_omp master, PCL 1.12).
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <array>
#include <memory>
#include <unordered_map>
#include <vector>
#include <omp.h>

#include "kdtree.hpp"
#include "linalg.hpp"
#include "oracle.h"

using namespace orc;

// ---- upstream-version variants (SURVEY Appendix A, items marked with a warning sign) ---------------------------------------
// The reference pins none of PCL / ndt_omp / fast_gicp, and their sources are not available here, so a handful of details of
// this restatement are "the surveyor's best knowledge of upstream master".  Each has a plausible alternative; orc_set_variant()
// switches the oracle to it so that tests/test_oracle_variants.py can MEASURE how far an alignment moves if upstream differs
// (the bound on the unpinned-parity risk).  0 is always the documented choice that the GPU engine reproduces.
int g_variant[ORC_VAR_COUNT] = {0};


namespace {

struct Pose {  // Eigen::Isometry3d
  double R[9];
  double t[3];
};

Pose pose_identity() { return Pose{{1, 0, 0, 0, 1, 0, 0, 0, 1}, {0, 0, 0}}; }
Pose pose_from_colmajor_f(const float* g) {
  Pose p;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) p.R[r * 3 + c] = (double)g[c * 4 + r];
    p.t[r] = (double)g[12 + r];
  }
  return p;
}
Pose pose_from_rowmajor_d(const double* T) {
  Pose p;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) p.R[r * 3 + c] = T[r * 4 + c];
    p.t[r] = T[r * 4 + 3];
  }
  return p;
}
void pose_to_colmajor_f(const Pose& p, float* g) {
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) g[c * 4 + r] = (float)p.R[r * 3 + c];
    g[12 + r] = (float)p.t[r];
    g[r * 4 + 3] = 0.f;
  }
  g[15] = 1.f;
}
Pose pose_mul(const Pose& a, const Pose& b) {
  Pose o;
  m3_mul(a.R, b.R, o.R);
  double rt[3];
  m3_vec(a.R, b.t, rt);
  for (int i = 0; i < 3; ++i) o.t[i] = rt[i] + a.t[i];
  return o;
}
inline void pose_apply(const Pose& T, const double* p, double* o) {
  double r[3];
  m3_vec(T.R, p, r);
  o[0] = r[0] + T.t[0]; o[1] = r[1] + T.t[1]; o[2] = r[2] + T.t[2];
}

// fast_gicp so3_exp (Sophus-derived) + Eigen Quaterniond::toRotationMatrix (App. A.1)
void so3_exp_matrix(const double* omega, double* R) {
  double theta_sq = omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2];
  double imag_factor, real_factor;
  if (theta_sq < 1e-10) {
    double theta_quad = theta_sq * theta_sq;
    imag_factor = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real_factor = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    double theta = std::sqrt(theta_sq);
    double half_theta = 0.5 * theta;
    imag_factor = std::sin(half_theta) / theta;
    real_factor = std::cos(half_theta);
  }
  double w = real_factor, x = imag_factor * omega[0], y = imag_factor * omega[1], z = imag_factor * omega[2];
  double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  double twx = tx * w, twy = ty * w, twz = tz * w;
  double txx = tx * x, txy = ty * x, txz = tz * x;
  double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

inline void cov6_to_m3(const double* c, double* M) {
  M[0] = c[0]; M[1] = c[1]; M[2] = c[2];
  M[3] = c[1]; M[4] = c[3]; M[5] = c[4];
  M[6] = c[2]; M[7] = c[4]; M[8] = c[5];
}

// fast_gicp calculate_covariances for one point, PLANE regularisation (App. A.1)
void point_covariance(const float* pts, const int* nbr, int k, double* cov6) {
  double mean[3] = {0, 0, 0};
  for (int j = 0; j < k; ++j) {
    const float* q = pts + 4 * (size_t)nbr[j];
    mean[0] += (double)q[0]; mean[1] += (double)q[1]; mean[2] += (double)q[2];
  }
  mean[0] /= k; mean[1] /= k; mean[2] /= k;
  double C[9] = {0};
  for (int j = 0; j < k; ++j) {
    const float* q = pts + 4 * (size_t)nbr[j];
    double d[3] = {(double)q[0] - mean[0], (double)q[1] - mean[1], (double)q[2] - mean[2]};
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) C[a * 3 + b] += d[a] * d[b];
  }
  for (int a = 0; a < 9; ++a) C[a] /= k;
  double ev[3], V[9];
  sym3_eigen(C, ev, V);  // ascending: column 0 = smallest => the plane normal
  const double values[3] = {1e-3, 1.0, 1.0};
  double out[9] = {0};
  for (int j = 0; j < 3; ++j)
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) out[a * 3 + b] += values[j] * V[a * 3 + j] * V[b * 3 + j];
  cov6[0] = out[0]; cov6[1] = out[1]; cov6[2] = out[2]; cov6[3] = out[4]; cov6[4] = out[5]; cov6[5] = out[8];
}

void calculate_covariances(const float* pts, int n, const KdTree& tree, int k, std::vector<double>& covs, int* knn_out) {
  covs.assign((size_t)n * 6, 0.0);
#pragma omp parallel
  {
    std::vector<int> idx(k);
    std::vector<float> d2(k);
#pragma omp for schedule(dynamic, 64)
    for (int i = 0; i < n; ++i) {
      int m = tree.knn(pts + 4 * (size_t)i, k, idx.data(), d2.data());
      for (int j = m; j < k; ++j) idx[j] = idx[m - 1];
      point_covariance(pts, idx.data(), k, &covs[(size_t)i * 6]);
      if (knn_out) std::memcpy(knn_out + (size_t)i * k, idx.data(), sizeof(int) * k);
    }
  }
}

// ---- fast_gicp GaussianVoxelMap (App. A.2) -----------------------------------
struct VoxKey {
  int x, y, z;
  bool operator==(const VoxKey& o) const { return x == o.x && y == o.y && z == o.z; }
  bool operator<(const VoxKey& o) const { return x != o.x ? x < o.x : (y != o.y ? y < o.y : z < o.z); }
};
struct VoxHash {
  size_t operator()(const VoxKey& k) const {
    size_t seed = 0;
    auto comb = [&](int v) { seed ^= std::hash<int>()(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2); };
    comb(k.x); comb(k.y); comb(k.z);
    return seed;
  }
};
struct GaussVoxel {
  int num_points = 0;
  double mean[3] = {0, 0, 0};
  double cov[6] = {0, 0, 0, 0, 0, 0};
};
struct GaussianVoxelMap {
  double resolution = 1.0;
  std::unordered_map<VoxKey, GaussVoxel, VoxHash> voxels;
  inline VoxKey coord(const double* x) const {
    const double half = g_variant[ORC_VAR_VGICP_COORD_NO_HALF] ? 0.0 : 0.5;  // fast_gicp: floor(x / resolution - 0.5)
    return VoxKey{(int)std::floor(x[0] / resolution - half), (int)std::floor(x[1] / resolution - half),
                  (int)std::floor(x[2] / resolution - half)};
  }
  void create(const float* pts, int n, const double* covs) {
    voxels.clear();
    voxels.reserve(n / 2 + 16);
    for (int i = 0; i < n; ++i) {  // sequential, point-index order
      double p[3] = {(double)pts[4 * (size_t)i], (double)pts[4 * (size_t)i + 1], (double)pts[4 * (size_t)i + 2]};
      GaussVoxel& v = voxels[coord(p)];
      v.num_points++;
      for (int d = 0; d < 3; ++d) v.mean[d] += p[d];
      for (int d = 0; d < 6; ++d) v.cov[d] += covs[(size_t)i * 6 + d];
    }
    for (auto& kv : voxels) {
      GaussVoxel& v = kv.second;
      for (int d = 0; d < 3; ++d) v.mean[d] /= v.num_points;
      for (int d = 0; d < 6; ++d) v.cov[d] /= v.num_points;
    }
  }
  const GaussVoxel* lookup(const VoxKey& k) const {
    auto it = voxels.find(k);
    return it == voxels.end() ? nullptr : &it->second;
  }
};

// M = (C_B + T C_A T^T)^-1 through the 4x4 path: RCR(3,3)=1, inverse, M(3,3)=0
void mahalanobis(const double* covB6, const double* covA6, const Pose& T, double* M9) {
  double CA[9], CB[9], RC[9], RCR[9];
  cov6_to_m3(covA6, CA);
  cov6_to_m3(covB6, CB);
  m3_mul(T.R, CA, RC);
  m3_mul_bt(RC, T.R, RCR);
  double A4[16] = {0}, inv4[16];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) A4[r * 4 + c] = CB[r * 3 + c] + RCR[r * 3 + c];
  A4[15] = 1.0;
  m4_inverse(A4, inv4);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) M9[r * 3 + c] = inv4[r * 4 + c];
}

// H += w J^T M J ; b += w J^T M e ; returns w e^T M e, with J = [skew(a) | -I]
inline double accumulate_hb(const double* a, const double* e, const double* M, double w, double* H, double* b) {
  double Me[3];
  m3_vec(M, e, Me);
  double err = w * (e[0] * Me[0] + e[1] * Me[1] + e[2] * Me[2]);
  if (!H) return err;
  double J[18] = {0, -a[2], a[1], -1, 0, 0, a[2], 0, -a[0], 0, -1, 0, -a[1], a[0], 0, 0, 0, -1};  // 3x6 row-major
  double MJ[18];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 6; ++c) MJ[r * 6 + c] = M[r * 3 + 0] * J[0 * 6 + c] + M[r * 3 + 1] * J[1 * 6 + c] + M[r * 3 + 2] * J[2 * 6 + c];
  for (int r = 0; r < 6; ++r) {
    for (int c = 0; c < 6; ++c) H[r * 6 + c] += w * (J[0 * 6 + r] * MJ[0 * 6 + c] + J[1 * 6 + r] * MJ[1 * 6 + c] + J[2 * 6 + r] * MJ[2 * 6 + c]);
    b[r] += w * (J[0 * 6 + r] * Me[0] + J[1 * 6 + r] * Me[1] + J[2 * 6 + r] * Me[2]);
  }
  return err;
}

// ---- NDT target cells: pclomp::VoxelGridCovariance (App. A.4) -----------------
struct NdtLeaf {
  int nr_points = 0;
  int idx = 0;                    // dense cell index (deterministic tie-break of the KDTREE neighbour order)
  float centroid[3] = {0, 0, 0};  // VoxelGridCovariance leaf.centroid: FLOAT sum in point order / float count (the kd-tree's points)
  double mean[3] = {0, 0, 0};
  double cov[9] = {0};
  double icov[9] = {0};
};
struct NdtGrid {
  float leaf = 1.f, inv_leaf = 1.f;
  int min_b[3] = {0, 0, 0}, max_b[3] = {0, 0, 0}, div_b[3] = {0, 0, 0}, mul[3] = {0, 0, 0};
  std::unordered_map<int, NdtLeaf> leaves;
  static constexpr int kMinPts = 6;
  static constexpr double kEigMult = 0.01;

  void build(const float* pts, int n, float leaf_size) {
    leaves.clear();
    leaf = leaf_size;
    inv_leaf = 1.0f / leaf_size;
    if (n == 0) return;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = 0; i < n; ++i)
      for (int d = 0; d < 3; ++d) {
        mn[d] = std::min(mn[d], pts[4 * (size_t)i + d]);
        mx[d] = std::max(mx[d], pts[4 * (size_t)i + d]);
      }
    int64_t dx = (int64_t)((mx[0] - mn[0]) * inv_leaf) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv_leaf) + 1,
            dz = (int64_t)((mx[2] - mn[2]) * inv_leaf) + 1;
    if (dx * dy * dz > (int64_t)INT32_MAX) return;  // PCL warns and leaves the grid empty
    for (int d = 0; d < 3; ++d) {
      min_b[d] = (int)std::floor(mn[d] * inv_leaf);
      max_b[d] = (int)std::floor(mx[d] * inv_leaf);
      div_b[d] = max_b[d] - min_b[d] + 1;
    }
    mul[0] = 1; mul[1] = div_b[0]; mul[2] = div_b[0] * div_b[1];
    leaves.reserve(n / 4 + 16);
    for (int i = 0; i < n; ++i) {  // pass 1, index order
      const float* p = pts + 4 * (size_t)i;
      int ijk0 = (int)(std::floor(p[0] * inv_leaf) - (float)min_b[0]);
      int ijk1 = (int)(std::floor(p[1] * inv_leaf) - (float)min_b[1]);
      int ijk2 = (int)(std::floor(p[2] * inv_leaf) - (float)min_b[2]);
      NdtLeaf& L = leaves[ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2]];
      L.idx = ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2];
      for (int a = 0; a < 3; ++a) L.centroid[a] += p[a];  // leaf.centroid.head<4>() += Vector4f(x, y, z, 0)
      double q[3] = {(double)p[0], (double)p[1], (double)p[2]};
      for (int a = 0; a < 3; ++a) {
        L.mean[a] += q[a];
        for (int b = 0; b < 3; ++b) L.cov[a * 3 + b] += q[a] * q[b];
      }
      ++L.nr_points;
    }
    for (auto& kv : leaves) {  // pass 2
      NdtLeaf& L = kv.second;
      double pt_sum[3] = {L.mean[0], L.mean[1], L.mean[2]};
      double npt = (double)L.nr_points;
      for (int a = 0; a < 3; ++a) L.mean[a] /= npt;
      for (int a = 0; a < 3; ++a) L.centroid[a] /= (float)L.nr_points;  // leaf.centroid /= static_cast<float>(nr_points)
      if (L.nr_points < kMinPts) continue;
      if (g_variant[ORC_VAR_NDT_COV_NEWER_PCL]) {  // newer PCL: (cov - pt_sum * mean^T) / (n - 1)
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) L.cov[a * 3 + b] = (L.cov[a * 3 + b] - pt_sum[a] * L.mean[b]) / (npt - 1.0);
      } else {
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b)
            L.cov[a * 3 + b] = (L.cov[a * 3 + b] - 2 * (pt_sum[a] * L.mean[b])) / npt + L.mean[a] * L.mean[b];
        for (int a = 0; a < 9; ++a) L.cov[a] *= (npt - 1.0) / npt;
      }
      // the single-pass form above is not exactly symmetric; the eigen solver reads the lower triangle
      double S[9];
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) S[a * 3 + b] = (a >= b) ? L.cov[a * 3 + b] : L.cov[b * 3 + a];
      double ev[3], V[9];
      sym3_eigen(S, ev, V);
      if (ev[0] < 0 || ev[1] < 0 || ev[2] <= 0) { L.nr_points = -1; continue; }
      double min_ev = kEigMult * ev[2];
      if (ev[0] < min_ev) {
        ev[0] = min_ev;
        if (ev[1] < min_ev) ev[1] = min_ev;
        double VL[9], Vinv[9];
        for (int a = 0; a < 3; ++a)
          for (int j = 0; j < 3; ++j) VL[a * 3 + j] = V[a * 3 + j] * ev[j];
        m3_inverse(V, Vinv);
        m3_mul(VL, Vinv, L.cov);
      }
      m3_inverse(L.cov, L.icov);
      bool bad = false;
      for (int a = 0; a < 9; ++a)
        if (std::isinf(L.icov[a])) bad = true;
      if (bad) L.nr_points = -1;
    }
  }

  // getNeighborhoodAtPoint{1,7,27}: lookup keys by float DIVISION by the leaf size
  int neighbors(const float* p, int mode, const NdtLeaf** out) const {
    static const int off7[7][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    if (leaves.empty()) return 0;
    int ijk[3] = {(int)std::floor(p[0] / leaf), (int)std::floor(p[1] / leaf), (int)std::floor(p[2] / leaf)};
    if (g_variant[ORC_VAR_NDT_LOOKUP_MUL])  // multiplication by the inverse leaf size, as the build does
      for (int d = 0; d < 3; ++d) ijk[d] = (int)std::floor(p[d] * inv_leaf);
    int cnt = 0;
    auto probe = [&](int ox, int oy, int oz) {
      int c[3] = {ijk[0] + ox, ijk[1] + oy, ijk[2] + oz};
      for (int d = 0; d < 3; ++d)
        if (c[d] < min_b[d] || c[d] > max_b[d]) return;
      int idx = (c[0] - min_b[0]) * mul[0] + (c[1] - min_b[1]) * mul[1] + (c[2] - min_b[2]) * mul[2];
      auto it = leaves.find(idx);
      if (it != leaves.end() && it->second.nr_points >= kMinPts) out[cnt++] = &it->second;
    };
    if (mode == ORC_DIRECT1) probe(0, 0, 0);
    else if (mode == ORC_DIRECT7)
      for (int i = 0; i < 7; ++i) probe(off7[i][0], off7[i][1], off7[i][2]);
    else
      for (int ox = -1; ox <= 1; ++ox)
        for (int oy = -1; oy <= 1; ++oy)
          for (int oz = -1; oz <= 1; ++oz) probe(ox, oy, oz);
    if (mode == ORC_KDTREE) {
      // pclomp KDTREE (registrations.cpp:140-141) = pcl::NormalDistributionsTransform's own search:
      //   target_cells_.radiusSearch(x_trans_pt, resolution_, neighborhood, distances)
      // i.e. a FLANN radius search (float L2_Simple distance, dist < (float)(radius * radius), results sorted by distance) over
      // the float centroids of the leaves with >= min_points_per_voxel points.  A centroid lies inside its own cell, so every
      // centroid closer than one leaf size sits in the 27 cells around the query's: the probe above followed by the distance
      // test returns the same set.  Equal distances are ordered by cell index here (FLANN leaves that order unspecified).
      const float r2 = (float)((double)leaf * (double)leaf);
      std::pair<float, const NdtLeaf*> hit[27];
      int m = 0;
      for (int i = 0; i < cnt; ++i) {
        const float* c = out[i]->centroid;
        const float dx = p[0] - c[0], dy = p[1] - c[1], dz = p[2] - c[2];
        float d = dx * dx;
        d = d + dy * dy;
        d = d + dz * dz;
        if (d < r2) hit[m++] = {d, out[i]};
      }
      std::sort(hit, hit + m, [](const std::pair<float, const NdtLeaf*>& a, const std::pair<float, const NdtLeaf*>& b) {
        return a.first != b.first ? a.first < b.first : a.second->idx < b.second->idx;
      });
      for (int i = 0; i < m; ++i) out[i] = hit[i].second;
      cnt = m;
    }
    return cnt;
  }
};

// float 4x4 (row-major here) = Translation3f(x,y,z) * AngleAxisf(rx,X) * AngleAxisf(ry,Y) * AngleAxisf(rz,Z)
// evaluated left to right in float (Transform * rotation-matrix products).  sin/cos are taken in double
// and rounded to float (models a correctly rounded sinf/cosf).
void ndt_matrix_from_p(const double* p, float* M /*row-major 3x4*/) {
  float rx = (float)p[3], ry = (float)p[4], rz = (float)p[5];
  float cx = (float)std::cos((double)rx), sx = (float)std::sin((double)rx);
  float cy = (float)std::cos((double)ry), sy = (float)std::sin((double)ry);
  float cz = (float)std::cos((double)rz), sz = (float)std::sin((double)rz);
  float Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx};
  float Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
  float Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
  float A[9], B[9];
  auto mul = [](const float* X, const float* Y, float* Z) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        float s = X[i * 3 + 0] * Y[0 * 3 + j];
        s = s + X[i * 3 + 1] * Y[1 * 3 + j];
        s = s + X[i * 3 + 2] * Y[2 * 3 + j];
        Z[i * 3 + j] = s;
      }
  };
  mul(Rx, Ry, A);
  mul(A, Rz, B);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) M[r * 4 + c] = B[r * 3 + c];
    M[r * 4 + 3] = (float)p[r];
  }
}
void rowmajor34_to_colmajor44(const float* M, float* T) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) T[c * 4 + r] = M[r * 4 + c];
  T[3] = T[7] = T[11] = 0.f;
  T[15] = 1.f;
}

// Eigen::Matrix3f::eulerAngles(0,1,2) (Eigen >= 3.3 algorithm), float
void euler_angles_012(const float* R /*row-major 3x3*/, float* res) {
  const int i = 0, j = 1, k = 2;
  auto c = [&](int r, int cc) { return R[r * 3 + cc]; };
  res[0] = std::atan2(c(j, k), c(k, k));
  float c2 = std::sqrt(c(i, i) * c(i, i) + c(i, j) * c(i, j));
  if (res[0] > 0.f && !g_variant[ORC_VAR_EULER_NO_FIXUP]) {  // (Eigen 3.2 had no such branch: first angle in [-pi, pi])
    if (res[0] > 0.f) res[0] -= (float)M_PI; else res[0] += (float)M_PI;
    res[1] = std::atan2(-c(i, k), -c2);
  } else {
    res[1] = std::atan2(-c(i, k), c2);
  }
  float s1 = std::sin(res[0]), c1 = std::cos(res[0]);
  res[2] = std::atan2(s1 * c(k, i) - c1 * c(j, i), c1 * c(j, j) - s1 * c(k, j));
  res[0] = -res[0]; res[1] = -res[1]; res[2] = -res[2];
}

}  // namespace

struct orc_reg {
  orc_params prm;
  std::vector<float> target, source;
  int nt = 0, ns = 0;
  // pcl::Registration base state
  float final_T[16];
  bool converged = false;
  int nr_iterations = 0;
  KdTree base_tree;  // tree_ (target), used by getFitnessScore
  bool base_tree_valid = false;
  // fast_gicp
  KdTree src_tree, tgt_tree;
  bool src_tree_valid = false, tgt_tree_valid = false;
  std::vector<double> source_covs, target_covs;
  std::unique_ptr<GaussianVoxelMap> voxelmap;
  std::vector<int> gicp_corr;
  std::vector<double> gicp_M;  // 9 per source point
  struct VCorr { int i; const GaussVoxel* v; VoxKey key; };
  std::vector<VCorr> vcorr;
  std::vector<double> vM;  // 9 per correspondence
  double lm_lambda = -1.0;
  int evals = 0;
  // ndt
  NdtGrid ndt;
  double gauss_d1 = 0, gauss_d2 = 0;
  float j_ang[8][3], h_ang[16][3];
  double j_ang_d[8][3], h_ang_d[16][3];  // ORC_VAR_NDT_INNER_DOUBLE: PCL / older ndt_omp keep these (and the inner products) in double
  std::vector<float> trans_cloud;

  int threads() const { return prm.num_threads > 0 ? prm.num_threads : omp_get_max_threads(); }

  // ------------------------------------------------------------------ GICP/VGICP
  void ensure_covariances() {
    int k = prm.correspondence_randomness;
    if ((int)source_covs.size() != ns * 6) {
      if (!src_tree_valid) { src_tree.build(source.data(), ns); src_tree_valid = true; }
      calculate_covariances(source.data(), ns, src_tree, k, source_covs, nullptr);
    }
    if ((int)target_covs.size() != nt * 6) {
      if (!tgt_tree_valid) { tgt_tree.build(target.data(), nt); tgt_tree_valid = true; }
      calculate_covariances(target.data(), nt, tgt_tree, k, target_covs, nullptr);
    }
  }

  void gicp_update_correspondences(const Pose& T) {
    if (!tgt_tree_valid) { tgt_tree.build(target.data(), nt); tgt_tree_valid = true; }
    gicp_corr.assign(ns, -1);
    gicp_M.assign((size_t)ns * 9, 0.0);
    float Tf[12];
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) Tf[r * 4 + c] = (float)T.R[r * 3 + c];
      Tf[r * 4 + 3] = (float)T.t[r];
    }
    const double thr2 = prm.max_correspondence_distance * prm.max_correspondence_distance;
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads())
    for (int i = 0; i < ns; ++i) {
      const float* p = &source[4 * (size_t)i];
      float q[4];
      for (int r = 0; r < 3; ++r) {  // ((c0*x + c1*y) + c2*z) + c3
        float s = Tf[r * 4 + 0] * p[0];
        s = s + Tf[r * 4 + 1] * p[1];
        s = s + Tf[r * 4 + 2] * p[2];
        q[r] = s + Tf[r * 4 + 3];
      }
      q[3] = 0.f;
      int idx; float d2;
      int m = tgt_tree.knn(q, 1, &idx, &d2);
      if (m == 1 && (double)d2 < thr2) {
        gicp_corr[i] = idx;
        mahalanobis(&target_covs[(size_t)idx * 6], &source_covs[(size_t)i * 6], T, &gicp_M[(size_t)i * 9]);
      }
    }
  }

  void vgicp_update_correspondences(const Pose& T) {
    static const int off7[7][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    vcorr.clear();
    const int nth = threads();
    std::vector<std::vector<VCorr>> per_thread(nth);  // concatenated in thread order, as upstream
#pragma omp parallel for schedule(static) num_threads(nth)
    for (int i = 0; i < ns; ++i) {
      std::vector<VCorr>& mine = per_thread[omp_get_thread_num()];
      double p[3] = {(double)source[4 * (size_t)i], (double)source[4 * (size_t)i + 1], (double)source[4 * (size_t)i + 2]}, a[3];
      pose_apply(T, p, a);
      VoxKey c = voxelmap->coord(a);
      auto probe = [&](int ox, int oy, int oz) {
        VoxKey k{c.x + ox, c.y + oy, c.z + oz};
        const GaussVoxel* v = voxelmap->lookup(k);
        if (v) mine.push_back(VCorr{i, v, k});
      };
      if (prm.neighbor_search == ORC_DIRECT1) probe(0, 0, 0);
      else if (prm.neighbor_search == ORC_DIRECT7)
        for (int o = 0; o < 7; ++o) probe(off7[o][0], off7[o][1], off7[o][2]);
      else
        for (int ox = -1; ox <= 1; ++ox)
          for (int oy = -1; oy <= 1; ++oy)
            for (int oz = -1; oz <= 1; ++oz) probe(ox, oy, oz);
    }
    for (auto& v : per_thread) vcorr.insert(vcorr.end(), v.begin(), v.end());
    vM.assign(vcorr.size() * 9, 0.0);
#pragma omp parallel for schedule(static) num_threads(threads())
    for (int c = 0; c < (int)vcorr.size(); ++c)
      mahalanobis(vcorr[c].v->cov, &source_covs[(size_t)vcorr[c].i * 6], T, &vM[(size_t)c * 9]);
  }

  double lsq_linearize(const Pose& T, double* H, double* b) {
    ++evals;
    if (prm.method == ORC_FAST_VGICP) {
      if (!voxelmap) {
        voxelmap.reset(new GaussianVoxelMap());
        voxelmap->resolution = prm.resolution;
        voxelmap->create(target.data(), nt, target_covs.data());
      }
      vgicp_update_correspondences(T);
    } else {
      gicp_update_correspondences(T);
    }
    return lsq_sum(T, H, b);
  }
  double lsq_compute_error(const Pose& T) {
    ++evals;
    return lsq_sum(T, nullptr, nullptr);
  }
  // per-thread partial sums, combined in thread order (upstream behaviour)
  double lsq_sum(const Pose& T, double* H, double* b) {
    int nth = threads();
    std::vector<double> Hs((size_t)nth * 36, 0.0), bs((size_t)nth * 6, 0.0), es(nth, 0.0);
    const bool vg = prm.method == ORC_FAST_VGICP;
    const int count = vg ? (int)vcorr.size() : ns;
#pragma omp parallel for schedule(static) num_threads(nth)
    for (int c = 0; c < count; ++c) {
      int tid = omp_get_thread_num();
      int i; const double* M; double mean_B[3]; double w = 1.0;
      if (vg) {
        i = vcorr[c].i; M = &vM[(size_t)c * 9];
        for (int d = 0; d < 3; ++d) mean_B[d] = vcorr[c].v->mean[d];
        w = std::sqrt((double)vcorr[c].v->num_points);
      } else {
        i = c;
        if (gicp_corr[i] < 0) continue;
        M = &gicp_M[(size_t)i * 9];
        const float* q = &target[4 * (size_t)gicp_corr[i]];
        mean_B[0] = q[0]; mean_B[1] = q[1]; mean_B[2] = q[2];
      }
      double p[3] = {(double)source[4 * (size_t)i], (double)source[4 * (size_t)i + 1], (double)source[4 * (size_t)i + 2]}, a[3];
      pose_apply(T, p, a);
      double e[3] = {mean_B[0] - a[0], mean_B[1] - a[1], mean_B[2] - a[2]};
      es[tid] += accumulate_hb(a, e, M, w, H ? &Hs[(size_t)tid * 36] : nullptr, H ? &bs[(size_t)tid * 6] : nullptr);
    }
    double err = 0;
    if (H) { std::fill(H, H + 36, 0.0); std::fill(b, b + 6, 0.0); }
    for (int t = 0; t < nth; ++t) {
      err += es[t];
      if (H) {
        for (int a = 0; a < 36; ++a) H[a] += Hs[(size_t)t * 36 + a];
        for (int a = 0; a < 6; ++a) b[a] += bs[(size_t)t * 6 + a];
      }
    }
    return err;
  }

  bool is_converged(const Pose& delta) const {
    double m = 0;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) m = std::max(m, std::fabs(delta.R[r * 3 + c] - (r == c ? 1.0 : 0.0)) / prm.rotation_epsilon);
    for (int r = 0; r < 3; ++r) m = std::max(m, std::fabs(delta.t[r]) / prm.transformation_epsilon);
    return m < 1.0;
  }

  bool step_lm(Pose& x0, Pose& delta, double& y0_out) {
    double H[36], b[6];
    double y0 = lsq_linearize(x0, H, b);
    y0_out = y0;
    if (lm_lambda < 0.0) {
      double mx = 0;
      for (int i = 0; i < 6; ++i) mx = std::max(mx, std::fabs(H[i * 6 + i]));
      lm_lambda = prm.lm_init_lambda_factor * mx;
    }
    double nu = 2.0;
    for (int i = 0; i < prm.lm_max_iterations; ++i) {
      double A[36], nb[6], d[6];
      std::memcpy(A, H, sizeof(A));
      for (int j = 0; j < 6; ++j) { A[j * 6 + j] += lm_lambda; nb[j] = -b[j]; }
      ldlt6_solve(A, nb, d);
      so3_exp_matrix(d, delta.R);
      delta.t[0] = d[3]; delta.t[1] = d[4]; delta.t[2] = d[5];
      Pose xi = pose_mul(delta, x0);
      double yi = lsq_compute_error(xi);
      double denom = 0;
      for (int j = 0; j < 6; ++j) denom += d[j] * (lm_lambda * d[j] - b[j]);
      double rho = (y0 - yi) / denom;
      if (rho < 0) {
        if (is_converged(delta)) return true;
        lm_lambda = nu * lm_lambda;
        nu = 2 * nu;
        continue;
      }
      x0 = xi;
      double f = 1 - std::pow(2 * rho - 1, 3);
      double third = 1.0 / 3.0;
      lm_lambda = lm_lambda * (third < f ? f : third);  // std::max(1/3, f) incl. its NaN behaviour
      return true;
    }
    return false;
  }

  int align_lsq(const float* guess, orc_result* out) {
    ensure_covariances();
    if (prm.method == ORC_FAST_VGICP) voxelmap.reset();  // upstream rebuilds the map on every align
    Pose x0 = pose_from_colmajor_f(guess);
    lm_lambda = -1.0;
    converged = false;
    evals = 0;
    double y0 = 0;
    for (int i = 0; i < prm.maximum_iterations && !converged; ++i) {
      nr_iterations = i;
      Pose delta = pose_identity();
      if (!step_lm(x0, delta, y0)) break;  // "lm not converged!!"
      converged = is_converged(delta);
    }
    pose_to_colmajor_f(x0, final_T);
    if (out) out->error = y0;
    return 0;
  }

  // ------------------------------------------------------------------ SMALL_GICP
  // small_gicp::RegistrationPCL with reg_type GICP, as src/mrg_slam/registrations.cpp:46-54 configures it (the shipped YAML
  // default, config/mrg_slam.yaml:100).  small_gicp is not vendored in /root/reference; restated from its public sources:
  //   GICPFactor::linearize / error, NearestNeighbor rejector (max_dist_sq), LevenbergMarquardtOptimizer::optimize
  //   (init_lambda 1e-3, lambda_factor 10, max_inner_iterations 10), TerminationCriteria (|rot| <= rotation_eps 2e-3,
  //   |trans| <= translation_eps), se3_exp, covariance estimation with k neighbours and (1e-3, 1, 1) regularisation.
  // Stated simplifications (oracle and GPU alike): the k-neighbour sets are selected with float32 squared distances like
  // FAST_GICP's (small_gicp ranks them in float64: sets differ only on near-ties), and the plane normal comes from the
  // Jacobi eigen-solver used everywhere here instead of Eigen's closed-form computeDirect.
  std::vector<int> sg_corr;
  std::vector<double> sg_M;  // 9 per source point, Mahalanobis matrix of the linearisation pose

  static void se3_exp(const double* a, Pose& out) {
    const double th2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
    const double th = std::sqrt(th2);
    so3_exp_matrix(a, out.R);
    if (th < 1e-10) {
      m3_vec(out.R, a + 3, out.t);
      return;
    }
    const double Om[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
    double Om2[9];
    m3_mul(Om, Om, Om2);
    const double c1 = (1.0 - std::cos(th)) / th2, c2 = (th - std::sin(th)) / (th2 * th);
    double V[9];
    for (int i = 0; i < 9; ++i) V[i] = ((i % 4 == 0) ? 1.0 : 0.0) + c1 * Om[i] + c2 * Om2[i];
    m3_vec(V, a + 3, out.t);
  }

  double sg_linearize(const Pose& T, double* H, double* b) {
    ++evals;
    if (!tgt_tree_valid) { tgt_tree.build(target.data(), nt); tgt_tree_valid = true; }
    sg_corr.assign(ns, -1);
    sg_M.assign((size_t)ns * 9, 0.0);
    const double max_d2 = prm.max_correspondence_distance * prm.max_correspondence_distance;
    const int nth = threads();
    std::vector<double> Hs((size_t)nth * 36, 0.0), bs((size_t)nth * 6, 0.0), es(nth, 0.0);
#pragma omp parallel for schedule(static) num_threads(nth)
    for (int i = 0; i < ns; ++i) {
      const int tid = omp_get_thread_num();
      const double p[3] = {(double)source[4 * (size_t)i], (double)source[4 * (size_t)i + 1], (double)source[4 * (size_t)i + 2]};
      double a[3];
      pose_apply(T, p, a);
      double d2;
      const int j = tgt_tree.nn_double(a, &d2);
      if (j < 0 || d2 > max_d2) continue;  // rejector: sq_dist > max_dist_sq
      sg_corr[i] = j;
      double CA[9], CB[9], RC[9], RCR[9], S[9], M[9];
      cov6_to_m3(&source_covs[(size_t)i * 6], CA);
      cov6_to_m3(&target_covs[(size_t)j * 6], CB);
      m3_mul(T.R, CA, RC);
      m3_mul_bt(RC, T.R, RCR);
      for (int t = 0; t < 9; ++t) S[t] = CB[t] + RCR[t];
      m3_inverse(S, M);
      std::memcpy(&sg_M[(size_t)i * 9], M, sizeof(M));
      const float* q = &target[4 * (size_t)j];
      const double r[3] = {(double)q[0] - a[0], (double)q[1] - a[1], (double)q[2] - a[2]};
      // J = [R skew(p) | -R]
      const double Sk[9] = {0, -p[2], p[1], p[2], 0, -p[0], -p[1], p[0], 0};
      double RS[9];
      m3_mul(T.R, Sk, RS);
      double J[18];
      for (int rr = 0; rr < 3; ++rr)
        for (int c = 0; c < 3; ++c) { J[rr * 6 + c] = RS[rr * 3 + c]; J[rr * 6 + 3 + c] = -T.R[rr * 3 + c]; }
      double MJ[18], Mr[3];
      for (int rr = 0; rr < 3; ++rr)
        for (int c = 0; c < 6; ++c) MJ[rr * 6 + c] = M[rr * 3 + 0] * J[0 * 6 + c] + M[rr * 3 + 1] * J[1 * 6 + c] + M[rr * 3 + 2] * J[2 * 6 + c];
      m3_vec(M, r, Mr);
      double* Ht = &Hs[(size_t)tid * 36];
      double* bt = &bs[(size_t)tid * 6];
      for (int rr = 0; rr < 6; ++rr) {
        for (int c = 0; c < 6; ++c) Ht[rr * 6 + c] += J[0 * 6 + rr] * MJ[0 * 6 + c] + J[1 * 6 + rr] * MJ[1 * 6 + c] + J[2 * 6 + rr] * MJ[2 * 6 + c];
        bt[rr] += J[0 * 6 + rr] * Mr[0] + J[1 * 6 + rr] * Mr[1] + J[2 * 6 + rr] * Mr[2];
      }
      es[tid] += 0.5 * (r[0] * Mr[0] + r[1] * Mr[1] + r[2] * Mr[2]);
    }
    double err = 0;
    std::fill(H, H + 36, 0.0);
    std::fill(b, b + 6, 0.0);
    for (int t = 0; t < nth; ++t) {
      err += es[t];
      for (int a = 0; a < 36; ++a) H[a] += Hs[(size_t)t * 36 + a];
      for (int a = 0; a < 6; ++a) b[a] += bs[(size_t)t * 6 + a];
    }
    return err;
  }

  double sg_error(const Pose& T) {  // GICPFactor::error with the correspondences and Mahalanobis matrices of the last linearize
    ++evals;
    const int nth = threads();
    std::vector<double> es(nth, 0.0);
#pragma omp parallel for schedule(static) num_threads(nth)
    for (int i = 0; i < ns; ++i) {
      if (sg_corr[i] < 0) continue;
      const double p[3] = {(double)source[4 * (size_t)i], (double)source[4 * (size_t)i + 1], (double)source[4 * (size_t)i + 2]};
      double a[3], Mr[3];
      pose_apply(T, p, a);
      const float* q = &target[4 * (size_t)sg_corr[i]];
      const double r[3] = {(double)q[0] - a[0], (double)q[1] - a[1], (double)q[2] - a[2]};
      m3_vec(&sg_M[(size_t)i * 9], r, Mr);
      es[omp_get_thread_num()] += 0.5 * (r[0] * Mr[0] + r[1] * Mr[1] + r[2] * Mr[2]);
    }
    double err = 0;
    for (int t = 0; t < nth; ++t) err += es[t];
    return err;
  }

  int align_small_gicp(const float* guess, orc_result* out) {
    ensure_covariances();
    Pose T = pose_from_colmajor_f(guess);
    const double lambda_factor = 10.0;
    double lambda = 1e-3;  // init_lambda
    converged = false;
    evals = 0;
    nr_iterations = 0;
    double e = 0;
    for (int i = 0; i < prm.maximum_iterations && !converged; ++i) {
      double H[36], b[6];
      e = sg_linearize(T, H, b);
      bool success = false;
      for (int j = 0; j < prm.lm_max_iterations; ++j) {  // max_inner_iterations = 10
        double A[36], nb[6], d[6];
        std::memcpy(A, H, sizeof(A));
        for (int k = 0; k < 6; ++k) { A[k * 6 + k] += lambda; nb[k] = -b[k]; }
        ldlt6_solve(A, nb, d);
        Pose dT;
        se3_exp(d, dT);
        const Pose new_T = pose_mul(T, dT);  // right-multiplied update
        const double new_e = sg_error(new_T);
        if (new_e <= e) {
          const double rn = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), tn = std::sqrt(d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);
          converged = rn <= prm.rotation_epsilon && tn <= prm.transformation_epsilon;
          T = new_T;
          lambda /= lambda_factor;
          success = true;
          e = new_e;
          break;
        }
        lambda *= lambda_factor;
      }
      nr_iterations = i;
      if (!success) break;
    }
    pose_to_colmajor_f(T, final_T);
    if (out) out->error = e;
    return 0;
  }

  // ------------------------------------------------------------------------ NDT
  void ndt_angle_derivatives(const double* p) {
    double cx, cy, cz, sx, sy, sz;
    const double aeps = g_variant[ORC_VAR_NDT_ANGLE_EPS_1E5] ? 1e-5 : 10e-5;  // upstream writes the literal 10e-5
    if (std::fabs(p[3]) < aeps) { cx = 1.0; sx = 0.0; } else { cx = std::cos(p[3]); sx = std::sin(p[3]); }
    if (std::fabs(p[4]) < aeps) { cy = 1.0; sy = 0.0; } else { cy = std::cos(p[4]); sy = std::sin(p[4]); }
    if (std::fabs(p[5]) < aeps) { cz = 1.0; sz = 0.0; } else { cz = std::cos(p[5]); sz = std::sin(p[5]); }
    const double J[8][3] = {{-sx * sz + cx * sy * cz, -sx * cz - cx * sy * sz, -cx * cy},
                            {cx * sz + sx * sy * cz, cx * cz - sx * sy * sz, -sx * cy},
                            {-sy * cz, sy * sz, cy},
                            {sx * cy * cz, -sx * cy * sz, sx * sy},
                            {-cx * cy * cz, cx * cy * sz, -cx * sy},
                            {-cy * sz, -cy * cz, 0},
                            {cx * cz - sx * sy * sz, -cx * sz - sx * sy * cz, 0},
                            {sx * cz + cx * sy * sz, cx * sy * cz - sx * sz, 0}};
    const double Hh[15][3] = {{-cx * sz - sx * sy * cz, -cx * cz + sx * sy * sz, sx * cy},   // a2
                              {-sx * sz + cx * sy * cz, -cx * sy * sz - sx * cz, -cx * cy},  // a3
                              {cx * cy * cz, -cx * cy * sz, cx * sy},                        // b2
                              {sx * cy * cz, -sx * cy * sz, sx * sy},                        // b3
                              {-sx * cz - cx * sy * sz, sx * sz - cx * sy * cz, 0},          // c2
                              {cx * cz - sx * sy * sz, -sx * sy * cz - cx * sz, 0},          // c3
                              {-cy * cz, cy * sz, sy},                                       // d1
                              {-sx * sy * cz, sx * sy * sz, sx * cy},                        // d2
                              {cx * sy * cz, -cx * sy * sz, -cx * cy},                       // d3
                              {sy * sz, sy * cz, 0},                                         // e1
                              {-sx * cy * sz, -sx * cy * cz, 0},                             // e2
                              {cx * cy * sz, cx * cy * cz, 0},                               // e3
                              {-cy * cz, cy * sz, 0},                                        // f1
                              {-cx * sz - sx * sy * cz, -cx * cz + sx * sy * sz, 0},         // f2
                              {-sx * sz + cx * sy * cz, -cx * sy * sz - sx * cz, 0}};        // f3
    for (int r = 0; r < 8; ++r)
      for (int c = 0; c < 3; ++c) { j_ang[r][c] = (float)J[r][c]; j_ang_d[r][c] = J[r][c]; }
    for (int r = 0; r < 15; ++r)
      for (int c = 0; c < 3; ++c) { h_ang[r][c] = (float)Hh[r][c]; h_ang_d[r][c] = Hh[r][c]; }
    for (int c = 0; c < 3; ++c) { h_ang[15][c] = 0.f; h_ang_d[15][c] = 0.0; }
  }

  // ndt_omp updateDerivatives / updateHessian, float inner math (App. A.3)
  // S = float: current ndt_omp; S = double: pcl::NormalDistributionsTransform / older ndt_omp (variant ORC_VAR_NDT_INNER_DOUBLE)
  template <typename S>
  inline double ndt_update(double* g, double* H, const S pg[3][6], const S ph[18][6], const double* x_trans, const double* c_inv,
                           bool do_grad, bool do_hess) const {
    S x4[3] = {(S)x_trans[0], (S)x_trans[1], (S)x_trans[2]};
    S C[3][3];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) C[a][b] = (S)c_inv[a * 3 + b];
    S xC[3];  // x4 * C4 (row vector)
    for (int j = 0; j < 3; ++j) {
      S s = x4[0] * C[0][j];
      s = s + x4[1] * C[1][j];
      s = s + x4[2] * C[2][j];
      xC[j] = s;
    }
    S xCx = x4[0] * xC[0];
    xCx = xCx + x4[1] * xC[1];
    xCx = xCx + x4[2] * xC[2];
    S gd2 = (S)gauss_d2;
    // exp taken in double and rounded to float: models a correctly rounded expf (glibc's is, in all but rare cases)
    S e = (S)std::exp((double)(-gd2 * xCx * (S)0.5));
    S score_inc = (S)(-gauss_d1 * (double)e);
    e = gd2 * e;
    if (e > 1 || e < 0 || e != e) return 0;
    e = (S)((double)e * gauss_d1);
    S cg[3][6];  // C4 * pg
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 6; ++c) {
        S s = C[a][0] * pg[0][c];
        s = s + C[a][1] * pg[1][c];
        s = s + C[a][2] * pg[2][c];
        cg[a][c] = s;
      }
    S xcg[6];
    for (int c = 0; c < 6; ++c) {
      S s = x4[0] * cg[0][c];
      s = s + x4[1] * cg[1][c];
      s = s + x4[2] * cg[2][c];
      xcg[c] = s;
    }
    if (do_grad)
      for (int c = 0; c < 6; ++c) g[c] += (double)(e * xcg[c]);
    if (do_hess) {
      S G[6][6];  // pg^T * cg
      for (int a = 0; a < 6; ++a)
        for (int c = 0; c < 6; ++c) {
          S s = pg[0][a] * cg[0][c];
          s = s + pg[1][a] * cg[1][c];
          s = s + pg[2][a] * cg[2][c];
          G[a][c] = s;
        }
      for (int i = 0; i < 6; ++i) {
        S xh[6];
        for (int j = 0; j < 6; ++j) {
          if (i < 3) { xh[j] = (S)0; continue; }
          const int rb = (i - 3) * 3;  // three live rows of the 4-row block i
          S s = xC[0] * ph[rb + 0][j];
          s = s + xC[1] * ph[rb + 1][j];
          s = s + xC[2] * ph[rb + 2][j];
          xh[j] = s;
        }
        for (int j = 0; j < 6; ++j) H[i * 6 + j] += (double)(e * (-gd2 * xcg[i] * xcg[j] + xh[j] + G[j][i]));
      }
    }
    return (double)score_inc;
  }

  // computeDerivatives / computeHessian over trans_cloud (App. A.3).  Per-point results summed in index order.
  template <typename S>
  void ndt_point(int i, const S (*jang)[3], const S (*hang)[3], bool do_grad, bool do_hess, int* hits_out, double* out) {
    const float* xt = &trans_cloud[4 * (size_t)i];
    const NdtLeaf* nb[27];
    int cnt = ndt.neighbors(xt, prm.neighbor_search, nb);
    if (hits_out) hits_out[i] = cnt;
    if (!cnt) return;
    const float* xo = &source[4 * (size_t)i];
    // computePointDerivatives
    S pg[3][6] = {{1, 0, 0, 0, 0, 0}, {0, 1, 0, 0, 0, 0}, {0, 0, 1, 0, 0, 0}};
    auto dot3 = [&](const S* row) {
      S s = row[0] * (S)xo[0];
      s = s + row[1] * (S)xo[1];
      s = s + row[2] * (S)xo[2];
      return s;
    };
    pg[1][3] = dot3(jang[0]); pg[2][3] = dot3(jang[1]);
    pg[0][4] = dot3(jang[2]); pg[1][4] = dot3(jang[3]); pg[2][4] = dot3(jang[4]);
    pg[0][5] = dot3(jang[5]); pg[1][5] = dot3(jang[6]); pg[2][5] = dot3(jang[7]);
    S ph[18][6];
    std::memset(ph, 0, sizeof(ph));
    if (do_hess) {
      S xh[15];
      for (int r = 0; r < 15; ++r) xh[r] = dot3(hang[r]);
      S a[3] = {0, xh[0], xh[1]}, b[3] = {0, xh[2], xh[3]}, c[3] = {0, xh[4], xh[5]};
      S d[3] = {xh[6], xh[7], xh[8]}, e[3] = {xh[9], xh[10], xh[11]}, f[3] = {xh[12], xh[13], xh[14]};
      for (int r = 0; r < 3; ++r) {
        ph[0 + r][3] = a[r]; ph[3 + r][3] = b[r]; ph[6 + r][3] = c[r];
        ph[0 + r][4] = b[r]; ph[3 + r][4] = d[r]; ph[6 + r][4] = e[r];
        ph[0 + r][5] = c[r]; ph[3 + r][5] = e[r]; ph[6 + r][5] = f[r];
      }
    }
    for (int k = 0; k < cnt; ++k) {
      double x_trans[3] = {(double)xt[0] - nb[k]->mean[0], (double)xt[1] - nb[k]->mean[1], (double)xt[2] - nb[k]->mean[2]};
      out[0] += ndt_update<S>(out + 1, out + 7, pg, ph, x_trans, nb[k]->icov, do_grad, do_hess);
    }
  }

  double ndt_derivatives(const double* p, double* grad, double* hess, bool do_grad, bool do_hess, int* hits_out) {
    ++evals;
    ndt_angle_derivatives(p);
    std::vector<double> per((size_t)ns * 43, 0.0);
    const bool inner_double = g_variant[ORC_VAR_NDT_INNER_DOUBLE] != 0;
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads())
    for (int i = 0; i < ns; ++i) {
      if (inner_double) ndt_point<double>(i, j_ang_d, h_ang_d, do_grad, do_hess, hits_out, &per[(size_t)i * 43]);
      else ndt_point<float>(i, j_ang, h_ang, do_grad, do_hess, hits_out, &per[(size_t)i * 43]);
    }
    double score = 0;
    if (do_grad) std::fill(grad, grad + 6, 0.0);
    std::fill(hess, hess + 36, 0.0);
    for (int i = 0; i < ns; ++i) {
      const double* in = &per[(size_t)i * 43];
      score += in[0];
      if (do_grad)
        for (int a = 0; a < 6; ++a) grad[a] += in[1 + a];
      if (do_hess)
        for (int a = 0; a < 36; ++a) hess[a] += in[7 + a];
    }
    return score;
  }

  void ndt_transform_source(const double* p) {
    float M[12], T[16];
    ndt_matrix_from_p(p, M);
    rowmajor34_to_colmajor44(M, T);
    std::memcpy(final_T, T, sizeof(T));
    trans_cloud.resize((size_t)ns * 4);
    orc_transform_cloud(source.data(), ns, T, trans_cloud.data());
  }

  static double psiMT(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
  static double dpsiMT(double g_a, double g_0, double mu) { return g_a - mu * g_0; }
  static bool updateIntervalMT(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u, double& g_u, double a_t, double f_t,
                               double g_t) {
    if (f_t > f_l) { a_u = a_t; f_u = f_t; g_u = g_t; return false; }
    else if (g_t * (a_l - a_t) > 0) { a_l = a_t; f_l = f_t; g_l = g_t; return false; }
    else if (g_t * (a_l - a_t) < 0) { a_u = a_l; f_u = f_l; g_u = g_l; a_l = a_t; f_l = f_t; g_l = g_t; return false; }
    return true;
  }
  static double trialValueSelectionMT(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t,
                                      double g_t) {
    if (f_t > f_l) {  // case 1
      double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
      double w = std::sqrt(z * z - g_t * g_l);
      double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
      double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
      if (std::fabs(a_c - a_l) < std::fabs(a_q - a_l)) return a_c;
      return 0.5 * (a_q + a_c);
    } else if (g_t * g_l < 0) {  // case 2
      double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
      double w = std::sqrt(z * z - g_t * g_l);
      double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
      double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
      if (std::fabs(a_c - a_t) >= std::fabs(a_s - a_t)) return a_c;
      return a_s;
    } else if (std::fabs(g_t) <= std::fabs(g_l)) {  // case 3
      double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
      double w = std::sqrt(z * z - g_t * g_l);
      double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
      double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
      double a_t_next = (std::fabs(a_c - a_t) < std::fabs(a_s - a_t)) ? a_c : a_s;
      if (a_t > a_l) return std::min(a_t + 0.66 * (a_u - a_t), a_t_next);
      return std::max(a_t + 0.66 * (a_u - a_t), a_t_next);
    } else {  // case 4
      double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u;
      double w = std::sqrt(z * z - g_t * g_u);
      return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
    }
  }

  double computeStepLengthMT(const double* x, double* step_dir, double step_init, double step_max, double step_min, double& score,
                             double* grad, double* hess) {
    double phi_0 = -score;
    double d_phi_0 = 0;
    for (int i = 0; i < 6; ++i) d_phi_0 += grad[i] * step_dir[i];
    d_phi_0 = -d_phi_0;
    if (d_phi_0 >= 0) {
      if (d_phi_0 == 0) return 0;
      d_phi_0 *= -1;
      for (int i = 0; i < 6; ++i) step_dir[i] *= -1;
    }
    const int max_step_iterations = 10;
    int step_iterations = 0;
    const double mu = 1.e-4, nu = 0.9;
    double a_l = 0, a_u = 0;
    double f_l = psiMT(a_l, phi_0, phi_0, d_phi_0, mu), g_l = dpsiMT(d_phi_0, d_phi_0, mu);
    double f_u = psiMT(a_u, phi_0, phi_0, d_phi_0, mu), g_u = dpsiMT(d_phi_0, d_phi_0, mu);
    bool interval_converged = (step_max - step_min) < 0, open_interval = true;
    double a_t = step_init;
    if (g_variant[ORC_VAR_MT_CLAMP_MAX_FIRST]) { a_t = std::max(a_t, step_min); a_t = std::min(a_t, step_max); }
    else { a_t = std::min(a_t, step_max); a_t = std::max(a_t, step_min); }
    double x_t[6];
    for (int i = 0; i < 6; ++i) x_t[i] = x[i] + step_dir[i] * a_t;
    ndt_transform_source(x_t);
    score = ndt_derivatives(x_t, grad, hess, true, true, nullptr);
    double phi_t = -score, d_phi_t = 0;
    for (int i = 0; i < 6; ++i) d_phi_t += grad[i] * step_dir[i];
    d_phi_t = -d_phi_t;
    double psi_t = psiMT(a_t, phi_t, phi_0, d_phi_0, mu), d_psi_t = dpsiMT(d_phi_t, d_phi_0, mu);
    while (!interval_converged && step_iterations < max_step_iterations && !(psi_t <= 0 && d_phi_t <= -nu * d_phi_0)) {
      if (open_interval) a_t = trialValueSelectionMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
      else a_t = trialValueSelectionMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
      if (g_variant[ORC_VAR_MT_CLAMP_MAX_FIRST]) { a_t = std::max(a_t, step_min); a_t = std::min(a_t, step_max); }
      else { a_t = std::min(a_t, step_max); a_t = std::max(a_t, step_min); }
      for (int i = 0; i < 6; ++i) x_t[i] = x[i] + step_dir[i] * a_t;
      ndt_transform_source(x_t);
      score = ndt_derivatives(x_t, grad, hess, true, false, nullptr);
      phi_t = -score;
      d_phi_t = 0;
      for (int i = 0; i < 6; ++i) d_phi_t += grad[i] * step_dir[i];
      d_phi_t = -d_phi_t;
      psi_t = psiMT(a_t, phi_t, phi_0, d_phi_0, mu);
      d_psi_t = dpsiMT(d_phi_t, d_phi_0, mu);
      if (open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
        open_interval = false;
        f_l = f_l + phi_0 - mu * d_phi_0 * a_l; g_l = g_l + mu * d_phi_0;
        f_u = f_u + phi_0 - mu * d_phi_0 * a_u; g_u = g_u + mu * d_phi_0;
      }
      if (open_interval) interval_converged = updateIntervalMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
      else interval_converged = updateIntervalMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
      step_iterations++;
    }
    if (step_iterations) {
      double dummy[6];
      ndt_derivatives(x_t, dummy, hess, false, true, nullptr);  // computeHessian
    }
    return a_t;
  }

  void ndt_gauss_constants() {
    double res = (double)(float)prm.resolution;  // resolution_ is a float member
    double c1 = 10 * (1 - prm.ndt_outlier_ratio);
    double c2 = prm.ndt_outlier_ratio / std::pow(res, 3);
    double d3 = -std::log(c2);
    gauss_d1 = -std::log(c1 + c2) - d3;
    gauss_d2 = -2 * std::log((-std::log(c1 * std::exp(-0.5) + c2) - d3) / gauss_d1);
  }

  int align_ndt(const float* guess, orc_result* out) {
    nr_iterations = 0;
    converged = false;
    evals = 0;
    ndt_gauss_constants();
    bool is_identity = true;
    for (int c = 0; c < 4; ++c)
      for (int r = 0; r < 4; ++r)
        if (guess[c * 4 + r] != (r == c ? 1.f : 0.f)) is_identity = false;
    trans_cloud = source;  // output = copy of the source
    if (!is_identity) {
      std::memcpy(final_T, guess, sizeof(final_T));
      orc_transform_cloud(source.data(), ns, guess, trans_cloud.data());
    }
    float Rf[9], eul[3];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Rf[r * 3 + c] = final_T[c * 4 + r];
    euler_angles_012(Rf, eul);
    double p[6] = {(double)final_T[12], (double)final_T[13], (double)final_T[14], (double)eul[0], (double)eul[1], (double)eul[2]};
    double grad[6], hess[36], delta_p[6];
    double score = ndt_derivatives(p, grad, hess, true, true, nullptr);
    while (!converged) {
      double ng[6];
      for (int i = 0; i < 6; ++i) ng[i] = -grad[i];
      svd6_solve(hess, ng, delta_p);
      double nrm = 0;
      for (int i = 0; i < 6; ++i) nrm += delta_p[i] * delta_p[i];
      nrm = std::sqrt(nrm);
      if (nrm == 0 || nrm != nrm) {
        converged = (nrm == nrm);
        if (out) out->error = score;
        return 0;
      }
      for (int i = 0; i < 6; ++i) delta_p[i] /= nrm;
      nrm = computeStepLengthMT(p, delta_p, nrm, prm.ndt_step_size, prm.transformation_epsilon / 2, score, grad, hess);
      for (int i = 0; i < 6; ++i) { delta_p[i] *= nrm; p[i] += delta_p[i]; }
      if (nr_iterations > prm.maximum_iterations || (nr_iterations && (std::fabs(nrm) < prm.transformation_epsilon))) converged = true;
      nr_iterations++;
    }
    if (out) out->error = score;
    return 0;
  }
};

// ------------------------------------------------------------------------ C API
extern "C" {

void orc_default_params(int method, orc_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->method = method;
  p->num_threads = 0;
  p->transformation_epsilon = 0.1;       // reg_transformation_epsilon (config/mrg_slam.yaml:102)
  p->maximum_iterations = 64;            // reg_maximum_iterations (:103)
  p->max_correspondence_distance = 2.0;  // reg_max_correspondence_distance (:104)
  p->correspondence_randomness = 20;     // reg_correspondence_randomness (:107)
  p->resolution = 1.0;                   // reg_resolution (:108)
  p->neighbor_search = (method == ORC_NDT_OMP) ? ORC_DIRECT7 : ORC_DIRECT1;
  p->rotation_epsilon = 2e-3;
  p->lm_max_iterations = 10;
  p->lm_init_lambda_factor = 1e-9;
  p->ndt_step_size = 0.1;
  p->ndt_outlier_ratio = 0.55;
}

orc_reg* orc_reg_create(const orc_params* p) {
  orc_reg* r = new orc_reg();
  r->prm = *p;
  for (int i = 0; i < 16; ++i) r->final_T[i] = (i % 5 == 0) ? 1.f : 0.f;
  return r;
}
void orc_reg_destroy(orc_reg* r) { delete r; }

void orc_reg_set_target(orc_reg* r, const float* xyzi, int n) {
  r->target.assign(xyzi, xyzi + (size_t)n * 4);
  r->nt = n;
  r->base_tree_valid = false;
  r->tgt_tree_valid = false;
  r->target_covs.clear();
  r->voxelmap.reset();
  if (r->prm.method == ORC_NDT_OMP) r->ndt.build(r->target.data(), n, (float)r->prm.resolution);
}
void orc_reg_set_source(orc_reg* r, const float* xyzi, int n) {
  r->source.assign(xyzi, xyzi + (size_t)n * 4);
  r->ns = n;
  r->src_tree_valid = false;
  r->source_covs.clear();
}

int orc_reg_align(orc_reg* r, const float* guess, orc_result* out) {
  // pcl::Registration::align: converged_=false, final_transformation_=I, then computeTransformation
  r->converged = false;
  for (int i = 0; i < 16; ++i) r->final_T[i] = (i % 5 == 0) ? 1.f : 0.f;
  if (out) out->error = 0;
  int rc = (r->prm.method == ORC_NDT_OMP)      ? r->align_ndt(guess, out)
           : (r->prm.method == ORC_SMALL_GICP) ? r->align_small_gicp(guess, out)
                                               : r->align_lsq(guess, out);
  if (out) {
    std::memcpy(out->T, r->final_T, sizeof(r->final_T));
    out->converged = r->converged ? 1 : 0;
    out->iterations = r->nr_iterations;
    out->lm_evals = r->evals;
  }
  return rc;
}

double orc_reg_fitness(orc_reg* r, double max_range) {
  return orc_fitness_score(r->target.data(), r->nt, r->source.data(), r->ns, r->final_T, max_range, nullptr);
}

void orc_knn_covariances(const float* xyzi, int n, int k, double* cov6_out, int* knn_idx_out) {
  KdTree tree;
  tree.build(xyzi, n);
  std::vector<double> covs;
  calculate_covariances(xyzi, n, tree, k, covs, knn_idx_out);
  std::memcpy(cov6_out, covs.data(), sizeof(double) * covs.size());
}

int orc_vgicp_voxelmap(const float* xyzi, int n, const double* cov6, double resolution, int* coords_out, int* npts_out, double* mean_out,
                       double* cov6_out) {
  GaussianVoxelMap vm;
  vm.resolution = resolution;
  vm.create(xyzi, n, cov6);
  std::vector<std::pair<VoxKey, const GaussVoxel*>> items;
  for (auto& kv : vm.voxels) items.emplace_back(kv.first, &kv.second);
  std::sort(items.begin(), items.end(), [](auto& a, auto& b) { return a.first < b.first; });
  int V = (int)items.size();
  for (int v = 0; v < V; ++v) {
    coords_out[v * 3] = items[v].first.x; coords_out[v * 3 + 1] = items[v].first.y; coords_out[v * 3 + 2] = items[v].first.z;
    npts_out[v] = items[v].second->num_points;
    for (int d = 0; d < 3; ++d) mean_out[v * 3 + d] = items[v].second->mean[d];
    for (int d = 0; d < 6; ++d) cov6_out[v * 6 + d] = items[v].second->cov[d];
  }
  return V;
}

double orc_reg_linearize(orc_reg* r, const double* T_rowmajor, double* H, double* b, int* corr_out, uint8_t* corr_valid) {
  r->ensure_covariances();
  Pose T = pose_from_rowmajor_d(T_rowmajor);
  if (r->prm.method == ORC_SMALL_GICP) {
    const double err = r->sg_linearize(T, H, b);
    if (corr_out)
      for (int i = 0; i < r->ns; ++i) corr_out[i] = r->sg_corr[i];
    return err;
  }
  double err = r->lsq_linearize(T, H, b);
  if (corr_out) {
    if (r->prm.method == ORC_FAST_VGICP) {
      // DIRECT1: at most one correspondence per point
      for (int i = 0; i < r->ns; ++i) { corr_valid[i] = 0; corr_out[i * 3] = corr_out[i * 3 + 1] = corr_out[i * 3 + 2] = 0; }
      for (auto& c : r->vcorr) {
        if (corr_valid[c.i]) continue;
        corr_valid[c.i] = 1;
        corr_out[c.i * 3] = c.key.x; corr_out[c.i * 3 + 1] = c.key.y; corr_out[c.i * 3 + 2] = c.key.z;
      }
    } else {
      for (int i = 0; i < r->ns; ++i) corr_out[i] = r->gicp_corr[i];
    }
  }
  return err;
}
double orc_reg_compute_error(orc_reg* r, const double* T_rowmajor) {
  Pose T = pose_from_rowmajor_d(T_rowmajor);
  if (r->prm.method == ORC_SMALL_GICP) return r->sg_error(T);
  return r->lsq_compute_error(T);
}

int orc_ndt_grid(const float* xyzi, int n, double resolution, int* idx_out, int* npts_out, double* mean_out, double* icov_out, int* min_b_out,
                 int* div_b_out) {
  NdtGrid g;
  g.build(xyzi, n, (float)resolution);
  std::vector<std::pair<int, const NdtLeaf*>> items;
  for (auto& kv : g.leaves) items.emplace_back(kv.first, &kv.second);
  std::sort(items.begin(), items.end(), [](auto& a, auto& b) { return a.first < b.first; });
  int V = (int)items.size();
  for (int v = 0; v < V; ++v) {
    idx_out[v] = items[v].first;
    npts_out[v] = items[v].second->nr_points;
    for (int d = 0; d < 3; ++d) mean_out[v * 3 + d] = items[v].second->mean[d];
    for (int d = 0; d < 9; ++d) icov_out[v * 9 + d] = items[v].second->icov[d];
  }
  for (int d = 0; d < 3; ++d) { min_b_out[d] = g.min_b[d]; div_b_out[d] = g.div_b[d]; }
  return V;
}

double orc_reg_ndt_derivatives(orc_reg* r, const double* p6, double* grad6, double* hess36, int* hits_out) {
  r->ndt_gauss_constants();
  r->ndt_transform_source(p6);
  return r->ndt_derivatives(p6, grad6, hess36, true, true, hits_out);
}

}  // extern "C"

// ---- hooks for the oracle's own unit tests (tests/test_oracle_*.py) ----
extern "C" {
void orc_test_ldlt6_solve(const double* A, const double* rhs, double* x) { ldlt6_solve(A, rhs, x); }
void orc_test_svd6_solve(const double* A, const double* rhs, double* x) { svd6_solve(A, rhs, x); }
void orc_test_sym3_eigen(const double* A, double* evals, double* V) { sym3_eigen(A, evals, V); }
void orc_test_m4_inverse(const double* A, double* inv) { m4_inverse(A, inv); }
void orc_test_so3_exp(const double* omega, double* R) { so3_exp_matrix(omega, R); }
void orc_test_euler_angles_012(const float* R_rowmajor, float* res) { euler_angles_012(R_rowmajor, res); }
void orc_test_ndt_matrix_from_p(const double* p6, float* M_rowmajor34) { ndt_matrix_from_p(p6, M_rowmajor34); }
double orc_test_mt_trial_value(const double* v9) {
  return orc_reg::trialValueSelectionMT(v9[0], v9[1], v9[2], v9[3], v9[4], v9[5], v9[6], v9[7], v9[8]);
}
}
