#pragma once
#include <memory>
#include <vector>
namespace pcl {
template <typename PointT>
struct PointCloud {
  std::vector<PointT> points;
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  size_t size() const { return points.size(); }
  void resize(size_t n) { points.resize(n); }
  PointT& operator[](size_t i) { return points[i]; }
  const PointT& operator[](size_t i) const { return points[i]; }
};
}  // namespace pcl
