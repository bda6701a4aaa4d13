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
// final_transformation_, converged_, nr_iterations_ and the output cloud.  getFitnessScore() is non-virtual and walks the base
// class's tree_ on the host, one nearestKSearch(point, 1) per source point, and initCompute() rebuilds that FLANN tree on every
// target change.  Both are taken off the host WITHOUT touching the callers (SURVEY 8b option (i)): the adapter installs a
// pcl::search::KdTree subclass (GpuServedKdTree) with setSearchMethodTarget(tree, /*force_no_recompute=*/true) whose
// setInputCloud builds nothing and whose nearestKSearch answers from a table computed on the GPU in one pass
// (b2r_nearest_neighbors) — so src/mrg_slam/loop_detector.cpp:137 and apps/scan_matching_odometry_component.cpp:403-415 get GPU
// answers unmodified.  Queries the table cannot serve (another k, a point that is not the next aligned source point) fall back to
// the real kd-tree, built lazily.  fitness() remains as the explicit, cheaper call (option (ii)): one number, no table.
#pragma once
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/registration/registration.h>
#include <pcl/search/kdtree.h>

#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <vector>

#include "registration.hpp"

namespace b2r {

// The search object handed to pcl::Registration (tree_).  pcl::Registration::getFitnessScore and the inlier loop of
// scan_matching_odometry_component.cpp:409-415 call nearestKSearch(aligned_point_i, 1, ...) for i = 0 .. n-1 in order; the table
// holds exactly those answers, so the i-th call after an align() is served from row i (the point is compared with the table's
// own transformed point, bit for bit, before the row is trusted).
class GpuServedKdTree : public pcl::search::KdTree<pcl::PointXYZI> {
 public:
  using PointT = pcl::PointXYZI;
  using Base = pcl::search::KdTree<PointT>;
  using PointCloudConstPtr = typename Base::PointCloudConstPtr;
  using IndicesConstPtr = typename Base::IndicesConstPtr;
  // fills (nearest target index, squared distance, transformed xyz) for every source point; false if unavailable
  std::function<bool(std::vector<int>&, std::vector<float>&, std::vector<float>&)> fetch;

  void setInputCloud(const PointCloudConstPtr& cloud, const IndicesConstPtr& = IndicesConstPtr()) override {
    cloud_ = cloud;  // no FLANN build: the GPU owns the target's search structures
    host_ready_ = false;
    invalidate();
  }
  void invalidate() { table_ready_ = false; cursor_ = 0; }  // a new align(): the table describes the previous result

  int nearestKSearch(const PointT& p, int k, pcl::Indices& k_indices, std::vector<float>& k_sqr_distances) const override {
    if (k == 1 && ensure_table()) {
      const size_t n = idx_.size();
      for (int attempt = 0; attempt < 2 && n; ++attempt) {  // the expected row, then row 0 (a caller starting a new pass)
        const size_t i = attempt == 0 ? (cursor_ < n ? cursor_ : 0) : 0;
        if (xyz_[3 * i] == p.x && xyz_[3 * i + 1] == p.y && xyz_[3 * i + 2] == p.z && idx_[i] >= 0) {
          k_indices.assign(1, idx_[i]);
          k_sqr_distances.assign(1, d2_[i]);
          cursor_ = i + 1;
          ++served;
          return 1;
        }
      }
    }
    if (!host_ready_ && cloud_) {  // anything else: the real kd-tree, built on first use
      const_cast<GpuServedKdTree*>(this)->Base::setInputCloud(cloud_);
      host_ready_ = true;
    }
    return Base::nearestKSearch(p, k, k_indices, k_sqr_distances);
  }
  mutable long served = 0;  // queries answered from the GPU table

 private:
  bool ensure_table() const {
    if (!table_ready_) {
      table_ready_ = true;  // one attempt per align()
      if (!fetch || !fetch(idx_, d2_, xyz_)) { idx_.clear(); d2_.clear(); xyz_.clear(); }
      cursor_ = 0;
    }
    return !idx_.empty();
  }
  PointCloudConstPtr cloud_;
  mutable bool host_ready_ = false, table_ready_ = false;
  mutable size_t cursor_ = 0;
  mutable std::vector<int> idx_;
  mutable std::vector<float> d2_, xyz_;
};

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
    // getFitnessScore() / getSearchMethodTarget()->nearestKSearch() of unmodified callers: served from the GPU; and no FLANN
    // build of the target in initCompute() (force_no_recompute)
    gpu_tree_.reset(new GpuServedKdTree);
    gpu_tree_->fetch = [this](std::vector<int>& idx, std::vector<float>& d2, std::vector<float>& xyz) {
      b2r_handle* h = impl_.handle();
      const size_t n = this->input_ ? this->input_->size() : 0;
      if (!h || n == 0 || !this->target_) return false;
      idx.resize(n); d2.resize(n); xyz.resize(3 * n);
      return b2r_nearest_neighbors(h, idx.data(), d2.data(), xyz.data()) == B2R_OK;
    };
    this->setSearchMethodTarget(gpu_tree_, /*force_no_recompute=*/true);
  }
  const GpuServedKdTree& gpuTree() const { return *gpu_tree_; }

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
    Base::setInputTarget(cloud);
    gpu_tree_->setInputCloud(cloud);  // remembers the cloud for the host fallback; builds nothing
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
    gpu_tree_->invalidate();
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
  std::shared_ptr<GpuServedKdTree> gpu_tree_;
  std::weak_ptr<const pcl::PointCloud<PointT>> last_pcl_[2];
  ::b2r::PointCloud::ConstPtr last_wrapped_[2];
  int slot_ = 0;
};

}  // namespace b2r
