// Minimal stand-in for <pcl/search/kdtree.h>: the virtual interface pcl::Registration walks in getFitnessScore()
// (tree_->nearestKSearch(point, 1, ...)) and rebuilds in initCompute() (tree_->setInputCloud(target_)).  The "FLANN tree" here is
// a brute-force search that counts how often it was built and queried, so the tests can tell host searches from served ones.
#pragma once
#include <cfloat>
#include <memory>
#include <vector>
#include "../point_cloud.h"
namespace pcl {
using Indices = std::vector<int>;
namespace search {
template <typename PointT>
class KdTree {
 public:
  using PointCloudConstPtr = typename pcl::PointCloud<PointT>::ConstPtr;
  using IndicesConstPtr = std::shared_ptr<const Indices>;
  using Ptr = std::shared_ptr<KdTree<PointT>>;
  virtual ~KdTree() = default;
  virtual void setInputCloud(const PointCloudConstPtr& cloud, const IndicesConstPtr& = IndicesConstPtr()) {
    input_ = cloud;
    ++host_builds;
  }
  virtual int nearestKSearch(const PointT& p, int k, Indices& k_indices, std::vector<float>& k_sqr_distances) const {
    ++host_queries;
    k_indices.assign(1, -1);
    k_sqr_distances.assign(1, FLT_MAX);
    if (!input_ || k != 1) return 0;
    for (size_t i = 0; i < input_->size(); ++i) {
      const PointT& q = input_->points[i];
      const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
      float d = dx * dx;
      d = d + dy * dy;
      d = d + dz * dz;
      if (d < k_sqr_distances[0]) { k_sqr_distances[0] = d; k_indices[0] = (int)i; }
    }
    return k_indices[0] >= 0 ? 1 : 0;
  }
  PointCloudConstPtr getInputCloud() const { return input_; }
  mutable long host_builds = 0, host_queries = 0;

 protected:
  PointCloudConstPtr input_;
};
}  // namespace search
}  // namespace pcl
