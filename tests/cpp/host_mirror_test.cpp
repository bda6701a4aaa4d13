// Exercises the C++ host mirror (include/b2r/registration.hpp) and the PCL adapter (against tests/cpp/pcl_stub)
// the way the reference's callers drive a registration object:
//   apps/scan_matching_odometry_component.cpp:203-275 (setInputTarget / setInputSource / align / hasConverged /
//   getFinalTransformation, keyframe switch) and src/mrg_slam/loop_detector.cpp:126-145.
// Usage: host_mirror_test [--expect-gpu]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include <b2r/pcl_adapter.hpp>
#include <b2r/registration.hpp>

extern "C" int b2r_synth_num_rays(int sensor);
extern "C" int b2r_synth_scan(int sensor, uint64_t seed, int scan_idx, float* out_xyzi);
extern "C" void b2r_synth_pose(uint64_t seed, int scan_idx, double* T16);

static int fails = 0;
#define CHECK(cond)                                                      \
  do {                                                                   \
    if (!(cond)) { std::printf("CHECK failed: %s (line %d)\n", #cond, __LINE__); ++fails; } \
  } while (0)

static b2r::PointCloud::Ptr make_cloud(int scan_idx) {
  const uint64_t seed = 0x5EED0000ull;
  std::vector<float> raw((size_t)b2r_synth_num_rays(0) * 4);
  int n = b2r_synth_scan(0, seed, scan_idx, raw.data());
  auto c = std::make_shared<b2r::PointCloud>();
  for (int i = 0; i < n; ++i) {
    const float* p = &raw[(size_t)i * 4];
    float d = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    if (d < 0.5f || d > 35.f || (i % 2)) continue;  // crude thinning, enough for a smoke-level check
    b2r::PointXYZI q; q.x = p[0]; q.y = p[1]; q.z = p[2]; q.intensity = p[3];
    c->points.push_back(q);
  }
  return c;
}

int main(int argc, char** argv) {
  const bool expect_gpu = argc > 1 && std::strcmp(argv[1], "--expect-gpu") == 0;
  // ---- factory dispatch (registrations.cpp:46-148)
  b2r::RegistrationParams prm;
  prm.registration_method = "FAST_VGICP";
  prm.reg_resolution = 1.0;
  auto vg = b2r::select_registration_method(prm);
  CHECK(vg && vg->config().method == B2R_FAST_VGICP && vg->config().correspondence_randomness == 20);
  prm.registration_method = "NDT_OMP"; prm.reg_nn_search_method = "DIRECT1"; prm.reg_resolution = 0.5;
  auto ndt = b2r::select_registration_method(prm);
  CHECK(ndt && ndt->config().method == B2R_NDT_OMP && ndt->config().neighbor_search == B2R_DIRECT1 && ndt->config().resolution == 0.5);
  prm.registration_method = "BOGUS";
  auto fb = b2r::select_registration_method(prm);
  CHECK(fb && fb->config().method == B2R_NDT_OMP);  // unknown -> warning + NDT
  prm.registration_method = "SMALL_GICP";  // the YAML default (registrations.cpp:46-54)
  prm.reg_max_correspondence_distance = 1.5;
  auto sg = b2r::select_registration_method(prm);
  CHECK(sg && sg->config().method == B2R_SMALL_GICP && sg->config().max_correspondence_distance == 1.5);
  prm.registration_method = "GICP_OMP";  // PCL's BFGS GICP: outside the engine
  CHECK(b2r::select_registration_method(prm) == nullptr);

  auto a = make_cloud(3), b = make_cloud(4), c = make_cloud(5);
  double Ta[16], Tb[16];
  b2r_synth_pose(0x5EED0000ull, 3, Ta);
  b2r_synth_pose(0x5EED0000ull, 4, Tb);
  const double gt_dx = Tb[3] - Ta[3];  // dominant forward motion, world frame ~ sensor frame for this trajectory

  // ---- odometry-style use of the mirror
  b2r::PointCloud aligned;
  vg->setInputTarget(a);
  vg->setInputSource(b);
  vg->align(aligned, b2r::identity4());
  if (!expect_gpu && !vg->hasConverged()) {
    // CPU-only box: the engine must refuse loudly, never fall back
    CHECK(vg->lastError().find("no CUDA device") != std::string::npos || !vg->lastError().empty());
    b2r::Matrix4f T = vg->getFinalTransformation();
    CHECK(T == b2r::identity4());  // failure leaves the guess
    std::printf("host mirror: no-device path ok (%s)\n", vg->lastError().c_str());
  } else {
    CHECK(vg->hasConverged());
    b2r::Matrix4f T = vg->getFinalTransformation();
    CHECK(std::fabs(T[12] - gt_dx) < 0.15);
    CHECK(aligned.size() == b->size());
    CHECK(std::fabs(aligned.points[0].x - (T[0] * b->points[0].x + T[4] * b->points[0].y + T[8] * b->points[0].z + T[12])) < 1e-4);
    const double f = vg->getFitnessScore();
    CHECK(f > 0 && f < 1.0);
    // keyframe switch: the current source becomes the target (scan_matching_odometry_component.cpp:332-333)
    vg->setInputTarget(b);
    vg->setInputSource(c);
    vg->align(aligned, b2r::identity4());
    CHECK(vg->hasConverged());
    // ---- the PCL adapter through pcl::Registration's own (stubbed) align()
    auto pa = std::make_shared<pcl::PointCloud<pcl::PointXYZI>>();
    auto pb = std::make_shared<pcl::PointCloud<pcl::PointXYZI>>();
    pa->points.resize(a->size()); pb->points.resize(b->size());
    std::memcpy(static_cast<void*>(pa->points.data()), a->points.data(), a->size() * 32);
    std::memcpy(static_cast<void*>(pb->points.data()), b->points.data(), b->size() * 32);
    b2r::PclRegistration::Ptr reg(new b2r::PclRegistration(B2R_FAST_VGICP));
    reg->setResolution(1.0);
    reg->setTransformationEpsilon(0.1);
    reg->setMaximumIterations(64);
    reg->setCorrespondenceRandomness(20);
    pcl::Registration<pcl::PointXYZI, pcl::PointXYZI>::Ptr base = reg;  // what select_registration_method returns
    base->setInputTarget(pa);
    base->setInputSource(pb);
    pcl::PointCloud<pcl::PointXYZI> out;
    base->align(out);
    CHECK(base->hasConverged());
    auto Tp = base->getFinalTransformation();
    for (int i = 0; i < 16; ++i) CHECK(Tp.data()[i] == T[i]);  // same engine, same bits
    CHECK(std::fabs(reg->fitness() - f) < 1e-12);
    std::printf("host mirror: gpu path ok, tx=%.4f (gt %.4f), fitness=%.5f\n", T[12], gt_dx, f);
  }
  std::printf(fails ? "FAILED (%d)\n" : "OK\n", fails);
  return fails ? 1 : 0;
}
