// Minimal stand-in for <pcl/registration/registration.h> reproducing the parts of pcl::Registration the adapter
// relies on: virtual setInputSource/Target, protected virtual computeTransformation, the result members, and the
// non-virtual align()/hasConverged()/getFinalTransformation().
#pragma once
#include <array>
#include <memory>
#include <string>
#include "../point_cloud.h"
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
  virtual ~Registration() = default;
  virtual void setInputSource(const PointCloudSourceConstPtr& c) { input_ = c; }
  virtual void setInputTarget(const PointCloudTargetConstPtr& c) { target_ = c; }
  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  void setMaximumIterations(int n) { max_iterations_ = n; }
  void setMaxCorrespondenceDistance(double d) { corr_dist_threshold_ = d; }
  void align(PointCloudSource& output, const Matrix4& guess = Matrix4::Identity()) {
    output.points.resize(input_->size());
    for (size_t i = 0; i < input_->size(); ++i) output.points[i] = input_->points[i];
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
};
}  // namespace pcl
