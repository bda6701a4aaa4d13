// ORACLE — TEST INFRASTRUCTURE ONLY.  Parity unpinned (see oracle/README.md).
//
// CPU restatement of the PCL filters the reference calls on its hot path:
//   pcl::VoxelGrid<PointXYZI>::applyFilter    <- apps/prefiltering_component.cpp:167-171,
//                                                apps/scan_matching_odometry_component.cpp:175-179
//   pcl::RadiusOutlierRemoval                 <- apps/prefiltering_component.cpp:195-199
//   pcl::StatisticalOutlierRemoval            <- apps/prefiltering_component.cpp:190-194
//   distance filter (in-tree, literal)        <- apps/prefiltering_component.cpp:206-229
//   pcl::transformPointCloud (float) and the fitness score, which exists in-tree
//   verbatim at src/mrg_slam/information_matrix_calculator.cpp:46-81.
// PCL itself is not vendored in /root/reference; algorithms follow SURVEY.md
// Appendix A.0, A.6, A.7, A.8, A.10, A.11.
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>
#include <omp.h>

#include "kdtree.hpp"
#include "oracle.h"

using orc::KdTree;

extern int g_variant[ORC_VAR_COUNT];  // registration.cpp

extern "C" {

void orc_set_variant(int id, int value) { if (id >= 0 && id < ORC_VAR_COUNT) g_variant[id] = value; }
int orc_get_variant(int id) { return (id >= 0 && id < ORC_VAR_COUNT) ? g_variant[id] : 0; }

// A.11 — keep iff near < |p| < far; norm evaluated in float as x^2 + (y^2 + z^2).
int orc_distance_filter(const float* xyzi, int n, double near_thresh, double far_thresh, float* out) {
  int m = 0;
  for (int i = 0; i < n; ++i) {
    const float* p = xyzi + 4 * (size_t)i;
    float s = g_variant[ORC_VAR_NORM_LEFT_TO_RIGHT] ? (p[0] * p[0] + p[1] * p[1]) + p[2] * p[2] : p[0] * p[0] + (p[1] * p[1] + p[2] * p[2]);
    double d = (double)std::sqrt(s);
    if (d > near_thresh && d < far_thresh) {
      std::memcpy(out + 4 * (size_t)m, p, 16);
      ++m;
    }
  }
  return m;
}

// A.6 — returns M (points written to out, capacity n), or -1 on the INT32
// overflow branch (PCL warns and copies the input to the output unchanged).
int orc_voxelgrid(const float* xyzi, int n, float leaf, int min_points_per_voxel, float* out, int* voxel_index_out) {
  if (n == 0) return 0;
  const float inv_leaf = 1.0f / leaf;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = 0; i < n; ++i) {
    const float* p = xyzi + 4 * (size_t)i;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    for (int d = 0; d < 3; ++d) { mn[d] = std::min(mn[d], p[d]); mx[d] = std::max(mx[d], p[d]); }
  }
  int64_t dx = (int64_t)((mx[0] - mn[0]) * inv_leaf) + 1;
  int64_t dy = (int64_t)((mx[1] - mn[1]) * inv_leaf) + 1;
  int64_t dz = (int64_t)((mx[2] - mn[2]) * inv_leaf) + 1;
  if (dx * dy * dz > (int64_t)INT32_MAX) {
    std::memcpy(out, xyzi, (size_t)n * 16);
    return -1;
  }
  int min_b[3], max_b[3], div_b[3], mul[3];
  for (int d = 0; d < 3; ++d) {
    min_b[d] = (int)std::floor(mn[d] * inv_leaf);
    max_b[d] = (int)std::floor(mx[d] * inv_leaf);
    div_b[d] = max_b[d] - min_b[d] + 1;
  }
  mul[0] = 1; mul[1] = div_b[0]; mul[2] = div_b[0] * div_b[1];
  std::vector<std::pair<int, int>> iv;  // (voxel idx, point index)
  iv.reserve(n);
  for (int i = 0; i < n; ++i) {
    const float* p = xyzi + 4 * (size_t)i;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    int ijk0 = (int)(std::floor(p[0] * inv_leaf) - (float)min_b[0]);
    int ijk1 = (int)(std::floor(p[1] * inv_leaf) - (float)min_b[1]);
    int ijk2 = (int)(std::floor(p[2] * inv_leaf) - (float)min_b[2]);
    iv.emplace_back(ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2], i);
  }
  // upstream std::sort is unstable => within-voxel order unspecified; the oracle
  // fixes ascending point index (pair ordering).
  std::sort(iv.begin(), iv.end());
  if (g_variant[ORC_VAR_VOXELGRID_DESCENDING])  // "any other order an unstable sort may leave": descending index inside a voxel
    std::sort(iv.begin(), iv.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) {
      return a.first != b.first ? a.first < b.first : a.second > b.second;
    });
  int m = 0;
  size_t first = 0;
  while (first < iv.size()) {
    size_t last = first + 1;
    while (last < iv.size() && iv[last].first == iv[first].first) ++last;
    if ((int)(last - first) >= min_points_per_voxel) {
      // pcl::CentroidPoint<PointXYZI>: float sums in run order, divided by float(count)
      float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
      for (size_t j = first; j < last; ++j) {
        const float* p = xyzi + 4 * (size_t)iv[j].second;
        sx += p[0]; sy += p[1]; sz += p[2]; si += p[3];
      }
      float cnt = (float)(last - first);
      float* o = out + 4 * (size_t)m;
      o[0] = sx / cnt; o[1] = sy / cnt; o[2] = sz / cnt; o[3] = si / cnt;
      if (voxel_index_out) voxel_index_out[m] = iv[first].first;
      ++m;
    }
    first = last;
  }
  return m;
}

// A.7 — keep[i] = (#neighbours within r, incl. self) > min_neighbors.
int orc_radius_outlier(const float* xyzi, int n, double radius, int min_neighbors, uint8_t* keep) {
  KdTree tree;
  tree.build(xyzi, n);
  const float r2 = (float)(radius * radius);
  int kept = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : kept)
  for (int i = 0; i < n; ++i) {
    const float* q = xyzi + 4 * (size_t)i;
    int k;
    if (min_neighbors == 1) {  // PCL special case: nearestKSearch(2), d2[1] <= r^2
      int idx[2]; float d2[2];
      k = tree.knn(q, 2, idx, d2);
      if (k == 2 && d2[1] > r2) k = 1;
    } else {
      k = tree.radius_count(q, g_variant[ORC_VAR_RADIUS_NONSTRICT] ? std::nextafter(r2, INFINITY) : r2, min_neighbors);
    }
    keep[i] = (k > min_neighbors) ? 1 : 0;
    kept += keep[i];
  }
  return kept;
}

// A.8 — distances[i] = (float)(sum_{j=1..k} sqrt(d2_j) / k); thr = mean + mul*stddev.
int orc_statistical_outlier(const float* xyzi, int n, int mean_k, double stddev_mul, uint8_t* keep, float* distances_out,
                            double* thr_out) {
  KdTree tree;
  tree.build(xyzi, n);
  std::vector<float> distances(n, 0.f);
  std::vector<uint8_t> valid(n, 0);
#pragma omp parallel
  {
    std::vector<int> idx(mean_k + 1);
    std::vector<float> d2(mean_k + 1);
#pragma omp for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
      const float* q = xyzi + 4 * (size_t)i;
      if (!std::isfinite(q[0]) || !std::isfinite(q[1]) || !std::isfinite(q[2])) continue;
      int k = tree.knn(q, mean_k + 1, idx.data(), d2.data());
      if (k != mean_k + 1) continue;  // PCL: warns, distance 0, not counted
      double dist_sum = 0.0;
      for (int j = 1; j < mean_k + 1; ++j) dist_sum += std::sqrt(d2[j]);  // float sqrt widened
      distances[i] = (float)(dist_sum / mean_k);
      valid[i] = 1;
    }
  }
  double sum = 0, sq_sum = 0;
  int nvalid = 0;
  for (int i = 0; i < n; ++i) {  // PCL sums over all entries (invalid ones are 0.0)
    sum += distances[i];
    sq_sum += distances[i] * distances[i];  // float product widened, as `sq_sum += distance * distance` on a float&
    nvalid += valid[i];
  }
  double mean = sum / nvalid;
  double variance = (sq_sum - sum * sum / nvalid) / (nvalid - 1.0);
  double stddev = std::sqrt(variance);
  double thr = mean + stddev_mul * stddev;
  int kept = 0;
  for (int i = 0; i < n; ++i) {
    keep[i] = (distances[i] > thr) ? 0 : 1;
    kept += keep[i];
  }
  if (distances_out) std::memcpy(distances_out, distances.data(), sizeof(float) * n);
  if (thr_out) *thr_out = thr;
  return kept;
}

// A.10 — float transform with PCL's SSE association (x*c0 + y*c1) + (z*c2 + c3).
void orc_transform_cloud(const float* xyzi, int n, const float* T_colmajor, float* out) {
  const float* c0 = T_colmajor; const float* c1 = T_colmajor + 4; const float* c2 = T_colmajor + 8; const float* c3 = T_colmajor + 12;
  for (int i = 0; i < n; ++i) {
    const float* p = xyzi + 4 * (size_t)i;
    float x = p[0], y = p[1], z = p[2];
    float* o = out + 4 * (size_t)i;
    for (int r = 0; r < 3; ++r) {
      float a = x * c0[r];
      float b = y * c1[r];
      float c = z * c2[r];
      o[r] = g_variant[ORC_VAR_TRANSFORM_LEFT_TO_RIGHT] ? ((a + b) + c) + c3[r] : (a + b) + (c + c3[r]);
    }
    o[3] = p[3];
  }
}

// A.0 getFitnessScore / information_matrix_calculator.cpp:46-81 (literal):
// mean squared 1-NN distance of T*source into target, only pairs with
// d2 <= max_range (note: SQUARED distance compared to max_range); DBL_MAX if none.
double orc_fitness_score(const float* target, int nt, const float* source, int ns, const float* T_colmajor, double max_range,
                         int* nr_out) {
  KdTree tree;
  tree.build(target, nt);
  std::vector<float> tr((size_t)ns * 4);
  orc_transform_cloud(source, ns, T_colmajor, tr.data());
  std::vector<float> d2s(ns);
#pragma omp parallel for schedule(dynamic, 256)
  for (int i = 0; i < ns; ++i) {
    int idx; float d2 = INFINITY;
    tree.knn(tr.data() + 4 * (size_t)i, 1, &idx, &d2);
    d2s[i] = d2;
  }
  double fitness = 0.0;
  int nr = 0;
  for (int i = 0; i < ns; ++i)  // index-order double sum, as the reference loop does
    if (d2s[i] <= max_range) { fitness += d2s[i]; ++nr; }
  if (nr_out) *nr_out = nr;
  return nr > 0 ? fitness / nr : DBL_MAX;
}

// kNN helper exposed for the tests' cross-checks against scipy.cKDTree.
void orc_knn(const float* xyzi, int n, const float* queries, int nq, int k, int* idx_out, float* d2_out) {
  KdTree tree;
  tree.build(xyzi, n);
#pragma omp parallel for schedule(dynamic, 256)
  for (int i = 0; i < nq; ++i) tree.knn(queries + 4 * (size_t)i, k, idx_out + (size_t)i * k, d2_out + (size_t)i * k);
}

void orc_set_num_threads(int n) { omp_set_num_threads(n > 0 ? n : omp_get_num_procs()); }
int orc_get_max_threads() { return omp_get_max_threads(); }

}  // extern "C"
