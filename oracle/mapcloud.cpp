// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of two IN-TREE functions (their sources are in /root/reference, so this part of the oracle follows
// code that can be read line by line; it cannot be compiled from there because it sits on PCL / Eigen / rclcpp types):
//   mrg_slam::MapCloudGenerator::generate          <- src/mrg_slam/map_cloud_generator.cpp:14-86
//   pcl::ApproximateMeanVoxelGrid<PointXYZI>       <- include/pcl/filters/ApproximateMeanVoxelGrid.hpp:63-126,
//                                                     include/pcl/filters/ApproximateMeanVoxelGrid.h:74-76 (hash)
// Eigen pieces restated (Eigen 3.4 as shipped with ROS 2 Humble, no FMA with the reference's -msse4.2 build):
//   Matrix4f * Vector4f (fixed-size lazy product)  -> per row ((m0*x + m1*y) + m2*z) + m3*w, w = 1
//   VectorXf += VectorXf, VectorXf /= float        -> element-wise float add / true division
//   Array3f::Ones() / leaf.array()                 -> 1.0f / leaf
//   Vector3f::squaredNorm                          -> x*x + (y*y + z*z) (non-vectorised unrolled reduction splits 3 as 1 + 2)
#include <cmath>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "oracle.h"

namespace {

struct Key {
  int v[3];
  bool operator==(const Key& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; }
};
struct KeyHash {  // ApproximateMeanVoxelGrid.h:74-76: size_t of an int expression (int arithmetic wraps, then widens)
  size_t operator()(const Key& p) const {
    return (size_t)(int)(((unsigned)p.v[0] * 73856093u) ^ ((unsigned)p.v[1] * 19349663u) ^ ((unsigned)p.v[2] * 83492791u));
  }
};
struct HistoryElement {
  int count = 0;
  float centroid[4] = {0.f, 0.f, 0.f, 0.f};
};

}  // namespace

extern "C" {

// poses: keyframe->pose.matrix() (Isometry3d), column-major doubles, 16 per keyframe; first_keyframe: flag per keyframe.
// Returns the number of points written to `out` (capacity sum(n)); -1 when generate() would return nullptr.
// Output order: std::unordered_map iteration order (implementation-defined upstream too); voxel_keys_out (3 ints per output
// point, optional) lets a test bring both sides into one order.
int orc_map_cloud(const float* const* clouds, const int* n, const double* poses_colmajor, const uint8_t* first_keyframe, int count,
                  float resolution, int min_points_per_voxel, float distance_far_thresh, int skip_first_cloud, float* out,
                  int* voxel_keys_out) {
  if (count == 0) return -1;  // :20-23
  std::vector<float> cloud;   // x,y,z,intensity packed
  const bool use_distance_filter = distance_far_thresh > 0;  // :28-29
  const float distance_far_thresh_sq = distance_far_thresh * distance_far_thresh;
  for (int k = 0; k < count; ++k) {
    if (first_keyframe[k] && skip_first_cloud) continue;  // :33-35
    float pose[16];
    for (int i = 0; i < 16; ++i) pose[i] = (float)poses_colmajor[(size_t)k * 16 + i];  // :36 cast<float>()
    const float* src = clouds[k];
    for (int i = 0; i < n[k]; ++i) {
      const float* p = src + 4 * (size_t)i;
      if (use_distance_filter) {
        const float sq = p[0] * p[0] + (p[1] * p[1] + p[2] * p[2]);  // Vector3f::squaredNorm, Eigen's unrolled 3-element redux
        if (sq > distance_far_thresh_sq) continue;  // :40-42
      }
      float d[3];
      for (int r = 0; r < 3; ++r) d[r] = ((pose[0 + r] * p[0] + pose[4 + r] * p[1]) + pose[8 + r] * p[2]) + pose[12 + r] * 1.0f;  // :44
      cloud.push_back(d[0]); cloud.push_back(d[1]); cloud.push_back(d[2]); cloud.push_back(p[3]);  // :45-46
    }
  }
  const size_t N = cloud.size() / 4;
  if (N == 0 && count > 1) return -1;  // :58-61
  if (resolution <= 0.0) {             // :67-71
    std::memcpy(out, cloud.data(), N * 16);
    return (int)N;
  }
  // ---- ApproximateMeanVoxelGrid::applyFilter, centroid_size = 4 (x, y, z, intensity)
  const float inv_leaf = 1.0f / resolution;
  std::unordered_map<Key, HistoryElement, KeyHash> history;
  for (size_t cp = 0; cp < N; ++cp) {
    const float* p = &cloud[4 * cp];
    Key ixyz{{(int)std::floor(p[0] * inv_leaf), (int)std::floor(p[1] * inv_leaf), (int)std::floor(p[2] * inv_leaf)}};
    HistoryElement& hhe = history[ixyz];
    hhe.count++;
    for (int c = 0; c < 4; ++c) hhe.centroid[c] += p[c];
  }
  int op = 0;
  for (auto& kv : history) {
    HistoryElement& hhe = kv.second;
    if (hhe.count && hhe.count >= min_points_per_voxel) {
      for (int c = 0; c < 4; ++c) out[4 * (size_t)op + c] = hhe.centroid[c] / (float)hhe.count;
      if (voxel_keys_out)
        for (int c = 0; c < 3; ++c) voxel_keys_out[3 * (size_t)op + c] = kv.first.v[c];
      ++op;
    }
  }
  return op;
}

}  // extern "C"
