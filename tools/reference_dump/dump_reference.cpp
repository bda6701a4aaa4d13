// Step 2 of tools/reference_dump/README.md — runs the REAL libraries of the reference on the exported cases.
// NOT compiled in this repository's container (no PCL / ndt_omp / fast_gicp there); written against their public APIs as the
// reference uses them: the constructor + setter sequences are those of src/mrg_slam/registrations.cpp:46-147, the filters those of
// apps/prefiltering_component.cpp:160-229.  Output: reference_dump.txt, one line per case.
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include <pcl/filters/radius_outlier_removal.h>
#include <pcl/filters/statistical_outlier_removal.h>
#include <pcl/filters/voxel_grid.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/registration/gicp.h>

#include <fast_gicp/gicp/fast_gicp.hpp>
#include <fast_gicp/gicp/fast_vgicp.hpp>
#include <pclomp/ndt_omp.h>
#ifdef WITH_SMALL_GICP
#include <small_gicp/pcl/pcl_registration.hpp>
#endif

using PointT = pcl::PointXYZI;
using Cloud = pcl::PointCloud<PointT>;

static Cloud::Ptr load(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  f.seekg(0, std::ios::end);
  const size_t n = (size_t)f.tellg() / 16;
  f.seekg(0);
  std::vector<float> raw(n * 4);
  f.read(reinterpret_cast<char*>(raw.data()), n * 16);
  Cloud::Ptr c(new Cloud);
  c->resize(n);
  for (size_t i = 0; i < n; ++i) { (*c)[i].x = raw[4 * i]; (*c)[i].y = raw[4 * i + 1]; (*c)[i].z = raw[4 * i + 2]; (*c)[i].intensity = raw[4 * i + 3]; }
  c->is_dense = false;
  return c;
}
static void save(const std::string& path, const Cloud& c) {
  std::vector<float> raw(c.size() * 4);
  for (size_t i = 0; i < c.size(); ++i) { raw[4 * i] = c[i].x; raw[4 * i + 1] = c[i].y; raw[4 * i + 2] = c[i].z; raw[4 * i + 3] = c[i].intensity; }
  std::ofstream(path, std::ios::binary).write(reinterpret_cast<const char*>(raw.data()), raw.size() * 4);
}

// the reg_* values of config/mrg_slam.yaml:100-109
static pcl::Registration<PointT, PointT>::Ptr make(const std::string& method, double resolution, const std::string& nn) {
  const double eps = 0.1, max_corr = 2.0;
  const int max_iter = 64, k = 20, threads = 8, opt_iter = 20;
  if (method == "FAST_GICP") {
    auto r = pcl::make_shared<fast_gicp::FastGICP<PointT, PointT>>();
    r->setNumThreads(threads); r->setTransformationEpsilon(eps); r->setMaximumIterations(max_iter);
    r->setMaxCorrespondenceDistance(max_corr); r->setCorrespondenceRandomness(k);
    return r;
  }
  if (method == "FAST_VGICP") {
    auto r = pcl::make_shared<fast_gicp::FastVGICP<PointT, PointT>>();
    r->setNumThreads(threads); r->setResolution(resolution); r->setTransformationEpsilon(eps); r->setMaximumIterations(max_iter);
    r->setCorrespondenceRandomness(k);
    return r;
  }
  if (method == "GICP") {
    auto r = pcl::make_shared<pcl::GeneralizedIterativeClosestPoint<PointT, PointT>>();
    r->setTransformationEpsilon(eps); r->setMaximumIterations(max_iter); r->setUseReciprocalCorrespondences(false);
    r->setMaxCorrespondenceDistance(max_corr); r->setCorrespondenceRandomness(k); r->setMaximumOptimizerIterations(opt_iter);
    return r;
  }
#ifdef WITH_SMALL_GICP
  if (method == "SMALL_GICP") {
    auto r = pcl::make_shared<small_gicp::RegistrationPCL<PointT, PointT>>();
    r->setNumThreads(threads); r->setTransformationEpsilon(eps); r->setMaximumIterations(max_iter);
    r->setMaxCorrespondenceDistance(max_corr); r->setCorrespondenceRandomness(k);
    return r;
  }
#endif
  if (method == "NDT_OMP") {
    auto r = pcl::make_shared<pclomp::NormalDistributionsTransform<PointT, PointT>>();
    r->setNumThreads(threads); r->setTransformationEpsilon(eps); r->setMaximumIterations(max_iter); r->setResolution(resolution);
    r->setNeighborhoodSearchMethod(nn == "KDTREE" ? pclomp::KDTREE : (nn == "DIRECT1" ? pclomp::DIRECT1 : pclomp::DIRECT7));
    return r;
  }
  return nullptr;
}

int main(int argc, char** argv) {
  const std::string dir = argc > 1 ? argv[1] : "/tmp/b2r_dump";
  std::ifstream cases(dir + "/cases.txt");
  std::ofstream out(dir + "/reference_dump.txt");
  out.precision(17);
  out << "# versions: PCL " << PCL_VERSION_PRETTY << " ndt_omp <git hash> fast_gicp <git hash>\n";
  std::string line;
  while (std::getline(cases, line)) {
    std::istringstream ss(line);
    std::string kind;
    ss >> kind;
    if (kind == "align") {
      int id; std::string method, tf, sf, nn; double res;
      ss >> id >> method >> tf >> sf >> res >> nn;
      Eigen::Matrix4f guess;
      for (int i = 0; i < 16; ++i) ss >> guess.data()[i];  // column-major, Eigen's storage
      auto reg = make(method, res, nn);
      if (!reg) { out << "align " << id << " skipped\n"; continue; }
      Cloud::Ptr tgt = load(dir + "/" + tf), src = load(dir + "/" + sf);
      reg->setInputTarget(tgt);
      reg->setInputSource(src);
      Cloud aligned;
      reg->align(aligned, guess);
      const Eigen::Matrix4f T = reg->getFinalTransformation();
      out << "align " << id << " " << method << " " << (reg->hasConverged() ? 1 : 0) << " " << reg->getFitnessScore();
      for (int i = 0; i < 16; ++i) out << " " << T.data()[i];
      out << "\n";
    } else if (kind == "filter") {
      int id; std::string f; double near_t, far_t, radius, sigma; float leaf; int min_nb, mean_k;
      ss >> id >> f >> near_t >> far_t >> leaf >> radius >> min_nb >> mean_k >> sigma;
      Cloud::Ptr raw = load(dir + "/" + f), dist(new Cloud), vg(new Cloud), rad(new Cloud), sor(new Cloud);
      for (const auto& p : *raw) {  // distance_filter, prefiltering_component.cpp:206-229
        const double d = p.getVector3fMap().norm();
        if (d > near_t && d < far_t) dist->push_back(p);
      }
      pcl::VoxelGrid<PointT> v; v.setLeafSize(leaf, leaf, leaf); v.setMinimumPointsNumberPerVoxel(1); v.setInputCloud(dist); v.filter(*vg);
      pcl::RadiusOutlierRemoval<PointT> r; r.setRadiusSearch(radius); r.setMinNeighborsInRadius(min_nb); r.setInputCloud(vg); r.filter(*rad);
      pcl::StatisticalOutlierRemoval<PointT> s; s.setMeanK(mean_k); s.setStddevMulThresh(sigma); s.setInputCloud(vg); s.filter(*sor);
      save(dir + "/dist_" + std::to_string(id) + ".bin", *dist);
      save(dir + "/vg_" + std::to_string(id) + ".bin", *vg);
      save(dir + "/radius_" + std::to_string(id) + ".bin", *rad);
      save(dir + "/sor_" + std::to_string(id) + ".bin", *sor);
      out << "filter " << id << " " << dist->size() << " " << vg->size() << " " << rad->size() << " " << sor->size() << "\n";
    }
  }
  std::cout << "wrote " << dir << "/reference_dump.txt" << std::endl;
  return 0;
}
