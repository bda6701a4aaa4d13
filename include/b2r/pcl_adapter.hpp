// pcl::Registration adapter: makes the libb2r engine a drop-in for the objects that
// mrg_slam::select_registration_method() returns (/root/reference/src/mrg_slam/registrations.cpp:28-152,
// declared at include/mrg_slam/registrations.hpp:20).
//
// Needs PCL (>= 1.11) and Eigen, i.e. the reference's own build environment; this repository's container has
// neither, so the header is syntax-checked against tests/cpp/pcl_stub/ only.  See INTEGRATION.md for the patch.
//
// What PCL makes virtual and what it does not:
//   virtual      : setInputSource, setInputTarget, computeTransformation (protected)          -> overridden here
//   non-virtual  : align, hasConverged, getFinalTransformation, getFitnessScore, getSearchMethodTarget
// align()/hasConverged()/getFinalTransformation() work unchanged because computeTransformation() fills
// final_transformation_, converged_, nr_iterations_ and the output cloud.  getFitnessScore() is non-virtual and walks
// the base class's FLANN tree_ on the host; callers that want the GPU version call fitness() (two call sites:
// src/mrg_slam/loop_detector.cpp:137, apps/scan_matching_odometry_component.cpp:403).
#pragma once
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/registration/registration.h>

#include <cstring>
#include <limits>
#include <memory>

#include "registration.hpp"

namespace b2r {

class PclRegistration : public pcl::Registration<pcl::PointXYZI, pcl::PointXYZI, float> {
 public:
  using PointT = pcl::PointXYZI;
  using Base = pcl::Registration<PointT, PointT, float>;
  using PointCloudSource = typename Base::PointCloudSource;
  using PointCloudSourceConstPtr = typename Base::PointCloudSourceConstPtr;
  using PointCloudTargetConstPtr = typename Base::PointCloudTargetConstPtr;
  using Matrix4 = typename Base::Matrix4;
  using Ptr = std::shared_ptr<PclRegistration>;

  explicit PclRegistration(int method, int device = 0) : impl_(method, device) {
    this->reg_name_ = method == B2R_NDT_OMP      ? "b2r::NDT_OMP"
                      : method == B2R_FAST_GICP  ? "b2r::FAST_GICP"
                      : method == B2R_SMALL_GICP ? "b2r::SMALL_GICP"
                      : method == B2R_GICP_PCL   ? "b2r::GICP"
                                                 : "b2r::FAST_VGICP";
    static_assert(sizeof(PointT) == sizeof(::b2r::PointXYZI), "pcl::PointXYZI layout changed");
  }

  // ---- the setters registrations.cpp calls
  void setNumThreads(int n) { impl_.setNumThreads(n); }
  void setTransformationEpsilon(double e) { Base::setTransformationEpsilon(e); impl_.setTransformationEpsilon(e); }
  void setMaximumIterations(int n) { Base::setMaximumIterations(n); impl_.setMaximumIterations(n); }
  void setMaxCorrespondenceDistance(double d) { Base::setMaxCorrespondenceDistance(d); impl_.setMaxCorrespondenceDistance(d); }
  void setCorrespondenceRandomness(int k) { impl_.setCorrespondenceRandomness(k); }
  void setResolution(double r) { impl_.setResolution(r); }
  void setUseReciprocalCorrespondences(bool on) { impl_.setUseReciprocalCorrespondences(on); }  // GICP / GICP_OMP (:99, :110)
  void setMaximumOptimizerIterations(int n) { impl_.setMaximumOptimizerIterations(n); }         // GICP / GICP_OMP (:102, :113)
  void setNeighborhoodSearchMethod(NeighborSearchMethod m) { impl_.setNeighborhoodSearchMethod(m); }

  // ---- virtuals of pcl::Registration
  void setInputSource(const PointCloudSourceConstPtr& cloud) override {
    Base::setInputSource(cloud);
    impl_.setInputSource(wrap(cloud));
  }
  void setInputTarget(const PointCloudTargetConstPtr& cloud) override {
    Base::setInputTarget(cloud);  // also marks the base kd-tree dirty; it is rebuilt lazily by initCompute()
    impl_.setInputTarget(wrap(cloud));
  }

  // GPU getFitnessScore(max_range): max_range is compared against the SQUARED distance, like PCL.
  double fitness(double max_range = std::numeric_limits<double>::max()) { return impl_.getFitnessScore(max_range); }
  ::b2r::Registration& engine() { return impl_; }

 protected:
  void computeTransformation(PointCloudSource& output, const Matrix4& guess) override {
    Matrix4f g;
    std::memcpy(g.data(), guess.data(), sizeof(float) * 16);  // Eigen::Matrix4f is column-major, like the C ABI
    ::b2r::PointCloud out;
    impl_.align(out, g);
    const Matrix4f T = impl_.getFinalTransformation();
    std::memcpy(this->final_transformation_.data(), T.data(), sizeof(float) * 16);
    this->transformation_ = this->final_transformation_;
    this->converged_ = impl_.hasConverged();
    this->nr_iterations_ = impl_.getNumberOfIterations();
    if (out.size() == output.size()) {
      for (size_t i = 0; i < out.size(); ++i) {
        output[i].x = out.points[i].x; output[i].y = out.points[i].y; output[i].z = out.points[i].z;
        output[i].intensity = out.points[i].intensity;
      }
    }
  }

 private:
  // The engine keys device clouds by the address of a b2r::PointCloud; keep one aliasing wrapper per PCL cloud so
  // that setting the same PCL cloud again (or promoting source to target) is recognised and reuses device structures.
  ::b2r::PointCloud::ConstPtr wrap(const typename pcl::PointCloud<PointT>::ConstPtr& cloud) {
    if (!cloud) return nullptr;
    // weak_ptr: a freed cloud whose address got recycled must not hit the cache
    for (int i = 0; i < 2; ++i)
      if (last_pcl_[i].lock() == cloud) return last_wrapped_[i];
    auto w = std::make_shared<::b2r::PointCloud>();
    w->points.resize(cloud->size());
    std::memcpy(static_cast<void*>(w->points.data()), static_cast<const void*>(cloud->points.data()), cloud->size() * sizeof(PointT));
    slot_ ^= 1;
    last_pcl_[slot_] = cloud;
    last_wrapped_[slot_] = w;
    return w;
  }

  ::b2r::Registration impl_;
  std::weak_ptr<const pcl::PointCloud<PointT>> last_pcl_[2];
  ::b2r::PointCloud::ConstPtr last_wrapped_[2];
  int slot_ = 0;
};

}  // namespace b2r
