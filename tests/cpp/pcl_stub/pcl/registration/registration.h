// Minimal stand-in for <pcl/registration/registration.h> reproducing the parts of pcl::Registration the adapter
// relies on: virtual setInputSource/Target, protected virtual computeTransformation, the result members, and the
// non-virtual align()/hasConverged()/getFinalTransformation().
#pragma once
#include <array>
#include <memory>
#include <string>
#include <cfloat>
#include <vector>
#include "../point_cloud.h"
#include "../search/kdtree.h"
namespace pcl {
struct StubMatrix4f {  // column-major 4x4 float, the storage of Eigen::Matrix4f
  std::array<float, 16> v{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}};
  float* data() { return v.data(); }
  const float* data() const { return v.data(); }
  static StubMatrix4f Identity() { return StubMatrix4f(); }
};
template <typename PointSource, typename PointTarget, typename Scalar = float>
class Registration {
 public:
  using Matrix4 = StubMatrix4f;
  using PointCloudSource = pcl::PointCloud<PointSource>;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = pcl::PointCloud<PointTarget>;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;
  using Ptr = std::shared_ptr<Registration>;
  using KdTree = pcl::search::KdTree<PointTarget>;
  using KdTreePtr = typename KdTree::Ptr;
  Registration() : tree_(new KdTree) {}
  virtual ~Registration() = default;
  virtual void setInputSource(const PointCloudSourceConstPtr& c) { input_ = c; }
  virtual void setInputTarget(const PointCloudTargetConstPtr& c) { target_ = c; target_cloud_updated_ = true; }
  // pcl::Registration::setSearchMethodTarget / getSearchMethodTarget (non-virtual)
  void setSearchMethodTarget(const KdTreePtr& tree, bool force_no_recompute = false) {
    tree_ = tree;
    force_no_recompute_ = force_no_recompute;
    target_cloud_updated_ = true;
  }
  KdTreePtr getSearchMethodTarget() const { return tree_; }
  // pcl::Registration::getFitnessScore(max_range) (non-virtual): transformPointCloud (SSE association), then one
  // tree_->nearestKSearch(point, 1, ...) per point on the host
  double getFitnessScore(double max_range = DBL_MAX) {
    double fitness_score = 0.0;
    const float* T = final_transformation_.data();
    Indices nn_indices(1);
    std::vector<float> nn_dists(1);
    int nr = 0;
    for (size_t i = 0; i < input_->size(); ++i) {
      const PointSource& s = input_->points[i];
      PointSource p = s;
      p.x = (s.x * T[0] + s.y * T[4]) + (s.z * T[8] + T[12]);
      p.y = (s.x * T[1] + s.y * T[5]) + (s.z * T[9] + T[13]);
      p.z = (s.x * T[2] + s.y * T[6]) + (s.z * T[10] + T[14]);
      tree_->nearestKSearch(p, 1, nn_indices, nn_dists);
      if (nn_dists[0] <= max_range) { fitness_score += nn_dists[0]; ++nr; }
    }
    return nr > 0 ? fitness_score / nr : DBL_MAX;
  }
  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  void setMaximumIterations(int n) { max_iterations_ = n; }
  void setMaxCorrespondenceDistance(double d) { corr_dist_threshold_ = d; }
  void align(PointCloudSource& output, const Matrix4& guess = Matrix4::Identity()) {
    output.points.resize(input_->size());
    for (size_t i = 0; i < input_->size(); ++i) output.points[i] = input_->points[i];
    // initCompute(): the target's kd-tree is rebuilt when the target changed, unless the caller asked it not to be
    if (target_cloud_updated_ && !force_no_recompute_) {
      tree_->setInputCloud(target_);
      target_cloud_updated_ = false;
    }
    converged_ = false;
    final_transformation_ = transformation_ = Matrix4::Identity();
    computeTransformation(output, guess);
  }
  bool hasConverged() const { return converged_; }
  Matrix4 getFinalTransformation() const { return final_transformation_; }

 protected:
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) = 0;
  std::string reg_name_;
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  Matrix4 final_transformation_, transformation_;
  bool converged_ = false;
  int nr_iterations_ = 0, max_iterations_ = 10;
  double transformation_epsilon_ = 0, corr_dist_threshold_ = 0;
  KdTreePtr tree_;
  bool target_cloud_updated_ = true, force_no_recompute_ = false;
};
}  // namespace pcl
