// Minimal stand-in for <pcl/point_types.h>: just enough to syntax-check include/b2r/pcl_adapter.hpp without PCL.
#pragma once
namespace pcl {
struct alignas(16) PointXYZI {
  float x = 0, y = 0, z = 0, data_w = 1.f;
  float intensity = 0, pad[3] = {0, 0, 0};
};
}  // namespace pcl
